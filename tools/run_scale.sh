#!/bin/bash
# usage: tools/run_scale.sh N TAG [env assignments...]  -> gpurun_out/scale_TAG_N.json
N=$1; TAG=$2; shift 2
env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 \
    bench.py --gpus $N --steps 40 --warmup 5 > gpurun_out/scale_${TAG}_${N}.json 2> gpurun_out/scale_${TAG}_${N}.err
echo "scale $TAG N=$N rc=$?"
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/scale_${TAG}_${N}.json").read().strip().splitlines()[-1])
    print("  value %.1f Mrays/s  ms/step %.4f  march %.4f  tail %.4f  %s" % (d["value"] / 1e6, d["ms_per_step"],
          d["step_roofline"]["march_ms"], d["step_roofline"]["exchange_plus_adam_ms"],
          json.dumps(d["step_roofline"].get("exchange_breakdown"))))
except Exception as e:
    print("  no result:", e)
PY
