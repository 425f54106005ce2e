#!/bin/bash
# Round-2 late call: GPU tests (graph-replayed train step included), smoke, full bench line.
TAG=${1:-r2g}
python -m pytest tests -m gpu -q -x > gpurun_out/gpu_tests_$TAG.log 2>&1; echo "tests rc=$?"; tail -15 gpurun_out/gpu_tests_$TAG.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke_$TAG.log
python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_$TAG.err
python - $TAG <<'PY'
import json, sys
d = json.loads(open("gpurun_out/bench_%s.json" % sys.argv[1]).read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"])
PY
