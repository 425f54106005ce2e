#!/bin/bash
# Round-2 late call: parity of the wide-stash march + vectorised raw2alpha, A/B timing, compute-sanitizer memcheck of smoke().
TAG=${1:-r2c}
python -m pytest tests -m gpu -q -x > gpurun_out/gpu_tests_$TAG.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/gpu_tests_$TAG.log
python tools/ab_march.py 12 > gpurun_out/ab_march_$TAG.md 2> gpurun_out/ab_march_$TAG.err; echo "ab rc=$?"; cat gpurun_out/ab_march_$TAG.md; tail -3 gpurun_out/ab_march_$TAG.err
python - > gpurun_out/raw2alpha_$TAG.txt 2>&1 <<'PY'
import torch, poseprobe_b200 as pp
ru = pp.render_utils_cuda
n = 2733151
d = torch.randn(n, device="cuda") * 4
gb = torch.randn(n, device="cuda")
for name, fn in (("raw2alpha fwd", lambda: ru.raw2alpha(d, -4.595, 0.5)),
                 ("raw2alpha fwd, 4 B-offset view (one element per thread)", lambda: ru.raw2alpha(dd, -4.595, 0.5)),
                 ("raw2alpha bwd", lambda: ru.raw2alpha_backward(e, gb, 0.5))):
    dd = torch.randn(n + 1, device="cuda")[1:]
    e, _ = ru.raw2alpha(d, -4.595, 0.5)
    for _ in range(5):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(50):
        fn()
    b.record(); torch.cuda.synchronize()
    print("%s: %.2f us per call (n = %d, allocation included)" % (name, a.elapsed_time(b) * 1e3 / 50, n))
PY
echo "raw2alpha rc=$?"; cat gpurun_out/raw2alpha_$TAG.txt
timeout 170 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/memcheck_smoke_$TAG.log 2>&1; echo "memcheck rc=$?"; tail -8 gpurun_out/memcheck_smoke_$TAG.log
