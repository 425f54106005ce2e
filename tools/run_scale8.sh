#!/bin/bash
# 8-GPU round: correctness first, then exchange variants, then the fine stage.
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port"
$T 29533 tests/dist_check.py > gpurun_out/dist_check8b.log 2>&1; echo "dist_check rc=$?"; tail -2 gpurun_out/dist_check8b.log
tools/run_scale.sh 8 nvls_v3_fused
tools/run_scale.sh 8 nvls_v3_unfused PP_XCHG_FUSED_SYNC=0
tools/run_scale.sh 8 nvls_v3_d6 PP_XCHG_DEPTH=6
tools/run_scale.sh 8 nvls_v3_d6_rpt16 PP_XCHG_DEPTH=6 PP_XCHG_RPT=16
tools/run_scale.sh 8 nvls_v3_d2_rpt2 PP_XCHG_DEPTH=2 PP_XCHG_RPT=2
$T 29536 tools/bench_voxurf_dist.py > gpurun_out/voxurf_dist8.json 2> gpurun_out/voxurf_dist8.err; echo "voxurf dist rc=$?"; tail -1 gpurun_out/voxurf_dist8.json | cut -c1-600
