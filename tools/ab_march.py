"""A/B timing of the coarse march's tuning switches at BASELINE configs[1] (160^3, 8192 rays) in ONE process.

The C library reads PP_MARCH_FLAGS / PP_MARCH_OCC (and, with profiles/r2_march_wide_stash_experiment.patch applied,
PP_MARCH_WIDE) with getenv at every launch, so the variants are alternated block by block on the same grids and rays
(same clocks, same L2 state): per variant the median of the block times of the march kernel alone (fwd + loss + bwd
in one launch) and of the whole step (march + masked Adam).  Without the patch the PP_MARCH_WIDE rows repeat the
shipped kernel.  profiles/r2_ab_march.md is this tool's output with the patch applied.
Usage: python tools/ab_march.py [blocks] ; prints one markdown table.
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402  (workload constants and batch maker only)
from poseprobe_b200 import dvgo, synthetic  # noqa: E402

VARIANTS = [("wide stash (default)", {"PP_MARCH_WIDE": "1"}),
            ("narrow stash (round-1 reverse walk)", {"PP_MARCH_WIDE": "0"}),
            ("wide, no pair merge", {"PP_MARCH_WIDE": "1", "PP_MARCH_FLAGS": "4"}),
            ("narrow, no pair merge", {"PP_MARCH_WIDE": "0", "PP_MARCH_FLAGS": "4"}),
            ("wide, 4 blocks per SM", {"PP_MARCH_WIDE": "1", "PP_MARCH_OCC": "4"})]
KEYS = ("PP_MARCH_WIDE", "PP_MARCH_FLAGS", "PP_MARCH_OCC")


def main():
    blocks = int(sys.argv[1]) if len(sys.argv) > 1 else 15
    steps = 20
    dev = torch.device("cuda", 0)
    density, k0 = synthetic.make_dvgo_grids(bench.GRID)
    eng = dvgo.FusedDVGO(density.to(dev), k0.to(dev), bench.XYZ_MIN, bench.XYZ_MAX, stepsize=bench.STEPSIZE)
    batches = [tuple(t.to(dev) for t in b) for b in bench.make_batches(bench.N_BATCHES, bench.RAYS, 0)]

    def use(env):
        for k in KEYS:
            os.environ.pop(k, None)
        os.environ.update(env)

    def block(march_only):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        for i in range(steps):
            o, d, t = batches[i % len(batches)]
            eng.loss_buf.zero_()
            eng.backward_only(o, d, t)
            if not march_only:
                eng.optimizer_step()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / steps

    times = {name: ([], []) for name, _ in VARIANTS}
    for name, env in VARIANTS:      # warm-up of every variant
        use(env)
        block(False)
    for _ in range(blocks):
        for name, env in VARIANTS:
            use(env)
            times[name][0].append(block(False))
            # the march alone accumulates into the gradient records without the Adam pass zeroing them: same traffic
            times[name][1].append(block(True))
            eng.optimizer_step()
    use({})
    print("| variant | step ms (march + Adam) | rays/s | march + loss-buffer fill only, ms |")
    print("|---|---|---|---|")
    for name, _ in VARIANTS:
        s, m = sorted(times[name][0]), sorted(times[name][1])
        sm, mm = s[len(s) // 2], m[len(m) // 2]
        print("| %s | %.4f | %.2f M | %.4f |" % (name, sm, bench.RAYS / sm / 1e3, mm))


if __name__ == "__main__":
    main()
