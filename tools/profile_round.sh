#!/bin/bash
# One GPU box call: tests, bench, launch list, and full ncu captures of the top kernels (run under gpurun).
TAG=${1:-r1f}
python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests_$TAG.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/gpu_tests_$TAG.log
python bench.py --steps 50 --warmup 5 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain_$TAG.log 2>&1 || { echo "plain run failed"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_list_$TAG.log 2>&1; echo "ncu list rc=$?"
# kernel-name regex : launches of that kernel to skip (the march kernel's first 8 launches are forward-only counts)
for KS in "dvgo_march_kernel:12" "voxel_adam_kernel:4" "rgbnet_backward2_kernel:4" "rgbnet_forward_kernel:4" "voxurf_march_kernel:4" "voxurf_backward_kernel:4"; do
  K=${KS%%:*}; S=${KS##*:}
  ncu --set full --clock-control none --import-source on -k regex:"$K" -s $S -c 1 -o gpurun_out/prof_${TAG}_$K $CMD > gpurun_out/ncu_full_${TAG}_$K.log 2>&1; echo "ncu $K rc=$?"
done
