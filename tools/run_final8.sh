#!/bin/bash
# Final 8-GPU call of round 2: NCCL ceiling probe, then the bench line exactly as the driver launches it.
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port"
timeout 200 $T 29551 tools/nccl_probe.py > gpurun_out/nccl_probe8.json 2> gpurun_out/nccl_probe8.err; echo "probe rc=$?"; tail -1 gpurun_out/nccl_probe8.json
timeout 800 $T 29552 bench.py --gpus 8 --steps 50 --warmup 5 > gpurun_out/scale8_r2_final.json 2> gpurun_out/scale8_r2_final.err; echo "bench rc=$?"
python - <<PY
import json
d = json.loads(open("gpurun_out/scale8_r2_final.json").read().strip().splitlines()[-1])
print("value %.1f M  ms %.4f  e2e %.1f M  %s" % (d["value"] / 1e6, d["ms_per_step"], d["e2e"]["value"] / 1e6, json.dumps(d["step_roofline"].get("exchange_breakdown"))))
for k in ("fine_stage", "config_d_render_256", "config_e_dual_branch"):
    v = d.get(k, {})
    print(k, {kk: v[kk] for kk in v if kk not in ("workload", "e2e")})
PY
