#!/bin/bash
# Round-2 8-GPU call: correctness of every exchange mode, then the hybrid (switch + peer) exchange against pure NVLS at
# several splits, then the default bench line with the config C / D / E extras.
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port"
$T 29533 tests/dist_check.py > gpurun_out/dist_check8_r2.log 2>&1; echo "dist_check rc=$?"; tail -6 gpurun_out/dist_check8_r2.log
run() {  # TAG env...
  TAG=$1; shift
  env "$@" $T 29541 bench.py --gpus 8 --steps 40 --warmup 5 --no-fine-stage > gpurun_out/scale8_r2_${TAG}.json 2> gpurun_out/scale8_r2_${TAG}.err
  echo "scale $TAG rc=$?"
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/scale8_r2_${TAG}.json").read().strip().splitlines()[-1])
    print("  value %.1f Mrays/s  ms/step %.4f  march %.4f  tail %.4f  %s" % (d["value"] / 1e6, d["ms_per_step"],
          d["step_roofline"]["march_ms"], d["step_roofline"]["exchange_plus_adam_ms"],
          json.dumps(d["step_roofline"].get("exchange_breakdown"))))
except Exception as e:
    print("  no result:", e)
PY
}
run nvls PP_DIST_MODE=nvls
run hyb50 PP_DIST_MODE=hybrid PP_XCHG_MC_FRACTION=0.5
run hyb65 PP_DIST_MODE=hybrid PP_XCHG_MC_FRACTION=0.65
run hyb80 PP_DIST_MODE=hybrid PP_XCHG_MC_FRACTION=0.8
run p2p PP_DIST_MODE=p2p
$T 29542 bench.py --gpus 8 --steps 40 --warmup 5 > gpurun_out/scale8_r2_full.json 2> gpurun_out/scale8_r2_full.err; echo "full rc=$?"
tail -c 1500 gpurun_out/scale8_r2_full.json
