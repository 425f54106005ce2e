#!/bin/bash
# Round-2 single-GPU call: tests, bench, per-operator tables, launch list, full ncu captures of the kernels this round
# changed or whose numbers the bench line cites (run under gpurun; every ncu pass only after its plain command exited 0).
TAG=${1:-r2}
python -m pytest tests -m gpu -q > gpurun_out/gpu_tests_$TAG.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/gpu_tests_$TAG.log
python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
python tools/bench_ops.py > gpurun_out/ops_$TAG.md 2> gpurun_out/ops_$TAG.err; echo "ops rc=$?"
python tools/profile_a2w.py > gpurun_out/ops_graph_$TAG.txt 2>&1; echo "a2w rc=$?"; cat gpurun_out/ops_graph_$TAG.txt
python tools/profile_scene.py > gpurun_out/scene_$TAG.md 2> gpurun_out/scene_$TAG.err; echo "scene rc=$?"; cat gpurun_out/scene_$TAG.md
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain_$TAG.log 2>&1 || { echo "plain run failed"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_list_$TAG.log 2>&1; echo "ncu list rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/ops_launches_$TAG.csv python tools/bench_ops.py > /dev/null 2>&1; echo "ncu ops list rc=$?"
for KS in "dvgo_march_kernel:12" "voxel_adam_kernel:4"; do
  K=${KS%%:*}; S=${KS##*:}
  ncu --set full --clock-control none --import-source on -k regex:"$K" -s $S -c 1 -o gpurun_out/prof_${TAG}_$K $CMD > gpurun_out/ncu_full_${TAG}_$K.log 2>&1; echo "ncu $K rc=$?"
done
for K in alpha2weight_fwd alpha2weight_bwd; do
  ncu --set full --clock-control none --import-source on -k regex:"$K" -s 3 -c 1 -o gpurun_out/prof_${TAG}_$K python tools/profile_a2w.py > gpurun_out/ncu_full_${TAG}_$K.log 2>&1; echo "ncu $K rc=$?"
done
for K in nerf_composite_fwd nerf_composite_bwd nerf_sample_pdf; do
  ncu --set full --clock-control none --import-source on -k regex:"$K" -s 8 -c 1 -o gpurun_out/prof_${TAG}_$K python tools/profile_scene.py > gpurun_out/ncu_full_${TAG}_$K.log 2>&1; echo "ncu $K rc=$?"
done
