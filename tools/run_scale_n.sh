#!/bin/bash
# usage: tools/run_scale_n.sh N  -> fine-stage exchange modes and the coarse headline at N ranks (gpurun --gpus N)
N=$1
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port"
for M in nvls nccl; do
  PP_VOXURF_DIST_MODE=$M timeout 300 $T 29535 tools/bench_voxurf_dist.py > gpurun_out/voxurf_dist${N}_r2_$M.json 2> gpurun_out/voxurf_dist${N}_r2_$M.err; echo "$M rc=$?"; tail -1 gpurun_out/voxurf_dist${N}_r2_$M.json | cut -c180-420
done
timeout 600 $T 29541 bench.py --gpus $N --steps 40 --warmup 5 > gpurun_out/scale${N}_r2_full.json 2> gpurun_out/scale${N}_r2_full.err; echo "bench rc=$?"
python - <<PY
import json
d = json.loads(open("gpurun_out/scale${N}_r2_full.json").read().strip().splitlines()[-1])
print("value %.1f M  ms %.4f  mode %s" % (d["value"] / 1e6, d["ms_per_step"], d.get("exchange", "")[:60]))
for k in ("fine_stage", "config_d_render_256", "config_e_dual_branch"):
    v = d.get(k, {})
    print(k, {kk: v[kk] for kk in v if kk not in ("workload", "e2e")})
PY
