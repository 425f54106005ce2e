#!/bin/bash
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port"
timeout 300 $T 29534 tests/dist_check_voxurf.py > gpurun_out/dist_check_voxurf8_r2.log 2>&1; echo "check rc=$?"; grep "dist_check_voxurf ok" gpurun_out/dist_check_voxurf8_r2.log
for M in nvls nccl; do
  PP_VOXURF_DIST_MODE=$M timeout 300 $T 29535 tools/bench_voxurf_dist.py > gpurun_out/voxurf_dist8_r2_$M.json 2> gpurun_out/voxurf_dist8_r2_$M.err; echo "$M rc=$?"; tail -1 gpurun_out/voxurf_dist8_r2_$M.json | cut -c1-500
done
python tools/bench_voxurf_dist.py 2>/dev/null | tail -1 | cut -c1-400
