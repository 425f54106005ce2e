#!/usr/bin/env python
"""bench.py -- train rays/s (fwd+bwd+optimiser) of the voxel-grid volume-rendering hot path.

Workload (BASELINE.json configs[1], "config B"): DVGO coarse stage, 160^3 density + 3-channel colour
grid, 8192 rays per batch per GPU, stepsize 0.5, synthetic rays / random-init grids (seed 777,
SURVEY.md section 8d).  A step = sample -> gather -> alpha -> composite -> colour -> MSE -> backward
(grid scatter) -> masked Adam on all grids (+ grid-gradient all-reduce when N > 1).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Prints ONE JSON line (rank 0).  `value` is device-timed throughput with inputs resident in HBM;
`e2e` is the same metric through the public Python API with pinned-host inputs copied in and the loss
copied out every step.  `--impl reference` times the reference formulation on the host CPU.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

GRID = 160
RAYS = 8192
STEPSIZE = 0.5
N_BATCHES = 8          # distinct ray batches cycled through the timed region
# dram__bytes_read.sum + dram__bytes_write.sum of one train-mode march launch, from the committed ncu
# capture of this round (per launch, like `achieved`; round 1 measured the same 350.8-350.9 MB)
MARCH_DRAM_TRAFFIC_BYTES = 350.8e6
MARCH_TRAFFIC_PROFILE = "profiles/r2_dvgo_march_kernel_full.txt"
XYZ_MIN, XYZ_MAX = [-1.0, -1.0, -1.0], [1.0, 1.0, 1.0]


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.proc, self.lines = None, []
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for nme, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_batches(n_batches, n_rays, rank):
    from poseprobe_b200 import synthetic
    out = []
    for b in range(n_batches):
        o, d, _ = synthetic.make_rays(n_rays, seed=777 + 1000 * rank + b)
        tgt = torch.rand(n_rays, 3, generator=torch.Generator().manual_seed(99 + 1000 * rank + b))
        out.append((o, d, tgt))
    return out


# =================================================================================================
# reference arm: the reference's own pure-PyTorch CPU formulation (lib/dvgo_ori.py), host cores
# =================================================================================================
def _load_by_path(name, rel):
    """Import a module of this repo by file path (the reference arm must not map the product's CUDA library, which
    `import poseprobe_b200` would do)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, rel))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def make_batches_cpu(n_batches, n_rays, rank):
    synthetic = _load_by_path("_pp_synthetic", "poseprobe_b200/synthetic.py")
    out = []
    for b in range(n_batches):
        o, d, _ = synthetic.make_rays(n_rays, seed=777 + 1000 * rank + b)
        tgt = torch.rand(n_rays, 3, generator=torch.Generator().manual_seed(99 + 1000 * rank + b))
        out.append((o, d, tgt))
    return out, synthetic


def run_reference(args, rank):
    """The reference's own CPU formulation of the path (lib/dvgo_ori.py: grid_sample + cumprod compositing + Python
    Adam, restated in oracle/dvgo_cpu.py and pinned against the reference's code) on the SAME workload: every step is
    the full 8192-ray batch on the 160^3 grid, all host threads.  Rank 0 only."""
    if rank != 0:
        return
    from oracle import dvgo_cpu
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    batches, synthetic = make_batches_cpu(1, RAYS, 0)
    density, k0 = synthetic.make_dvgo_grids(GRID)
    base = dvgo_cpu.DenseBaseline(density, k0, XYZ_MIN, XYZ_MAX, STEPSIZE)
    o, d, tgt = batches[0]
    for _ in range(args.warmup):
        base.train_step(o, d, tgt)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        base.train_step(o, d, tgt)
    dt = time.perf_counter() - t0
    value = RAYS * args.steps / dt
    line = {
        "impl": "reference", "metric": "train_rays_per_s", "value": value, "unit": "rays/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.gpus),
        "cpu_baseline": {"value": value, "unit": "rays/s", "cores": cores, "kind": "port",
                         "sample": "all %d rays of the batch every step, lib/dvgo_ori.py dense formulation "
                                   "(grid_sample + cumprod + Python Adam) restated in oracle/dvgo_cpu.py" % RAYS},
        "e2e": {"value": value, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def workload_config(n_gpus):
    return {"workload": "DVGO coarse stage (BASELINE configs[1]): 160^3 density+k0(3) grid, 8192 rays/batch/GPU, "
                        "stepsize 0.5, fwd+bwd+masked Adam, TV off",
            "grid": GRID, "rays_per_gpu": RAYS, "global_rays": RAYS * n_gpus, "stepsize": STEPSIZE,
            "l2": "no flush: per-step working set (vox+grad+2 Adam moments = 262 MB) exceeds the 126 MB L2",
            "parallelism": ("ray-sharded replicas (dp%d), voxel-sharded optimiser" % n_gpus) if n_gpus > 1
            else "single GPU"}


# =================================================================================================
# our arm
# =================================================================================================
def run_ours(args, rank, world, local_rank):
    import torch.distributed as dist
    import poseprobe_b200 as pp
    from poseprobe_b200 import dist as ppdist
    from poseprobe_b200 import dvgo, synthetic
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    dist_mode = os.environ.get("PP_DIST_MODE", "auto")
    density, k0 = synthetic.make_dvgo_grids(GRID)
    if world > 1:
        trainer = ppdist.DistributedDVGO(density.to(dev), k0.to(dev), XYZ_MIN, XYZ_MAX, STEPSIZE, mode=dist_mode)
        eng = trainer.eng
    else:
        trainer = None
        eng = dvgo.FusedDVGO(density.to(dev), k0.to(dev), XYZ_MIN, XYZ_MAX, stepsize=STEPSIZE)
    host = [tuple(t.pin_memory() for t in b) for b in make_batches(N_BATCHES, RAYS, rank)]
    resident = [tuple(t.to(dev) for t in b) for b in host]

    def step(o, d, tgt, events=None):
        if trainer is not None:
            return trainer.train_step(o, d, tgt, events)
        eng.loss_buf.zero_()
        if events is not None:
            events[0].record()
        eng.backward_only(o, d, tgt)
        if events is not None:
            events[1].record()
        eng.optimizer_step()
        if events is not None:
            events[2].record()
        return eng.loss_buf

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # samples actually composited per batch (for the algorithmic byte count)
    counts = []
    for o, d, _ in resident:
        _, _, _, cnt = eng.render(o, d, want_depth=False, want_count=True)
        counts.append(int(cnt.sum()))
    m_bar = sum(counts) / len(counts)

    sampler = ClockSampler(local_rank) if rank == 0 else None
    for i in range(max(args.warmup, 3)):
        step(*resident[i % N_BATCHES])
    barrier()

    # ---- timed region (device timed, max over ranks) ----
    # A block = EXACTLY args.steps steps between a barrier + synchronize on both sides.  One block of 20 steps is
    # only ~6 ms of device time, so the block is repeated until >= 0.5 s has been timed and the MEDIAN block is the
    # reported one (min / max alongside); the repeat count is agreed on rank 0's first block.
    def timed_block(with_events):
        k_ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)] if with_events else None
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        a.record()
        for i in range(args.steps):
            step(*resident[i % N_BATCHES], events=k_ev[i] if with_events else None)
        b.record()
        barrier()
        ms = torch.tensor([a.elapsed_time(b)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms), k_ev

    launches0 = pp._lib.launch_count()
    first_ms, k_ev = timed_block(True)
    launches = pp._lib.launch_count() - launches0
    n_blocks = torch.tensor([min(200, max(1, int(500.0 / max(first_ms, 1e-3)) + 1))], device=dev)
    if world > 1:
        dist.broadcast(n_blocks, 0)
    block_ms = [first_ms] + [timed_block(False)[0] for _ in range(int(n_blocks) - 1)]
    block_ms.sort()
    ms_total = block_ms[len(block_ms) // 2]
    march_ms = sum(e[0].elapsed_time(e[1]) for e in k_ev) / args.steps
    tail_ms = sum(e[1].elapsed_time(e[2]) for e in k_ev) / args.steps

    # ---- end to end through the public API: pinned host inputs in, loss out, every step ----
    # a batch (origins, directions, target colours) is one pinned [3,N,3] block: one H2D copy per step
    stream = torch.cuda.current_stream()
    host_packed = [torch.stack(b).pin_memory() for b in host]
    dpack = torch.empty(tuple(host_packed[0].shape), dtype=torch.float32, device=dev)
    loss_host = torch.zeros(2).pin_memory()
    # Three ways of making the same call sequence, all with every step's rays copied from pinned host memory and every
    # step's (loss, mse) read back to the host inside the timed region:
    #   "pipeline" (reported): FusedDVGO.host_pipeline - the copy of batch k+1 runs on a copy stream under step k and
    #               the host reads the loss of step k while step k+1 runs (one step of look-ahead, as a loader has);
    #   "graph":    FusedDVGO.train_step_host - copy in + step + copy out as ONE CUDA-graph launch, stream
    #               synchronised after every step;
    #   "eager":    the same six stream operations issued one by one, stream synchronised after every step.
    def e2e_block(mode):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        if mode == "pipeline":
            pipe.submit(host_packed[0])
            for i in range(args.steps):
                if i + 1 < args.steps:
                    pipe.submit(host_packed[(i + 1) % N_BATCHES])
                loss_seen[0] = pipe.result()      # the caller reads the loss of every step
        else:
            for i in range(args.steps):
                if mode == "graph":
                    eng.train_step_host(host_packed[i % N_BATCHES], loss_host)
                else:
                    dpack.copy_(host_packed[i % N_BATCHES], non_blocking=True)
                    loss_host.copy_(step(dpack[0], dpack[1], dpack[2])[:2], non_blocking=True)
                stream.synchronize()              # the caller reads the loss of every step
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    def e2e_median(mode, blocks):
        e2e_block(mode)                           # untimed: graph capture / stream and slot creation / warm copies
        t = sorted(e2e_block(mode) for _ in range(blocks))
        return t[len(t) // 2], t

    loss_seen = [None]
    e2e_side = {}
    from poseprobe_b200.dvgo import HostPipeline
    if trainer is None:
        e2e_mode = os.environ.get("PP_BENCH_E2E", "pipeline")
        pipe = eng.host_pipeline(RAYS, depth=2)
        for m in ("graph", "eager"):
            e2e_side[m + "_ms_per_step"] = e2e_median(m, max(3, int(n_blocks) // 4))[0] / args.steps
    else:
        # N ranks: the in-line sequence round 2 measured at 2 / 4 / 8 GPUs.  The pipeline also wraps the distributed
        # step (PP_BENCH_E2E=pipeline) but has not been timed on N > 1 GPUs, so it is not the reported form there.
        e2e_mode = os.environ.get("PP_BENCH_E2E", "eager")
        if e2e_mode == "pipeline":
            pipe = HostPipeline(eng, RAYS, depth=2, step_fn=lambda o, d, t: trainer.train_step(o, d, t))
    e2e_ms, e2e_blocks = e2e_median(e2e_mode, int(n_blocks))
    # ---- forward-only render throughput (the metric's "render Mrays/s"): rays are sharded, no communication ----
    render_iters = 40
    for i in range(3):
        eng.render(*resident[i % N_BATCHES][:2], want_depth=True)
    barrier()
    r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    r0.record()
    for i in range(render_iters):
        eng.render(*resident[i % N_BATCHES][:2], want_depth=True)
    r1.record()
    barrier()
    render_ms = torch.tensor([r0.elapsed_time(r1)], device=dev)
    if world > 1:
        dist.all_reduce(render_ms, op=dist.ReduceOp.MAX)
    render_ms = float(render_ms) / render_iters
    # the same through the public API with host buffers: ray chunk in (pinned), rgb + depth out (pinned), every chunk
    rgb_host = torch.empty((RAYS, 3)).pin_memory()
    depth_host = torch.empty(RAYS).pin_memory()
    barrier()
    r0.record()
    for i in range(render_iters):
        dpack.copy_(host_packed[i % N_BATCHES], non_blocking=True)
        out = eng.render(dpack[0], dpack[1], want_depth=True)
        rgb_host.copy_(out[0], non_blocking=True)
        depth_host.copy_(out[2], non_blocking=True)
        stream.synchronize()
    r1.record()
    barrier()
    render_e2e_ms = torch.tensor([r0.elapsed_time(r1)], device=dev)
    if world > 1:
        dist.all_reduce(render_e2e_ms, op=dist.ReduceOp.MAX)
    render_e2e_ms = float(render_e2e_ms) / render_iters
    # the same work as ONE call on host buffers (FusedDVGO.render_host): the chunks' copies in / out run on copy streams
    # under the neighbouring chunks' kernels, one synchronisation at the end of the image
    # (reported at N = 1; N > 1 keeps the per-chunk form round 2 measured there unless PP_BENCH_RENDER_HOST=1)
    render_host_ms, n_img = None, 0
    if world == 1 or os.environ.get("PP_BENCH_RENDER_HOST", "0") == "1":
        img_rays = torch.cat([hp[:2] for hp in host_packed] * (render_iters // N_BATCHES), dim=1).contiguous().pin_memory()
        n_img = img_rays.shape[1]
        img_rgb = torch.empty((n_img, 3)).pin_memory()
        img_depth = torch.empty(n_img).pin_memory()
        eng.render_host(img_rays, img_rgb, img_depth, chunk=RAYS)
        host_ms = []
        for _ in range(5):
            barrier()
            r0.record()
            eng.render_host(img_rays, img_rgb, img_depth, chunk=RAYS)
            r1.record()
            barrier()
            t = torch.tensor([r0.elapsed_time(r1)], device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            host_ms.append(float(t))
        render_host_ms = sorted(host_ms)[len(host_ms) // 2] / (n_img / RAYS)
    # keep the GPU busy a little longer so the clock sampler sees the loaded state (the timed region is short)
    # The step COUNT is derived from the (all-reduced, hence rank-identical) median block time, never from a rank's own
    # wall clock: the distributed step carries cross-rank flag barriers, so every rank must issue the same number of
    # steps - a rank that issued one more would sit in that barrier until its 60 s budget ran out.
    busy_steps = max(args.steps, min(5000, int(600.0 / max(ms_total / args.steps, 1e-3))))
    for i in range(busy_steps):
        step(*resident[i % N_BATCHES])
    barrier()
    clocks = sampler.stop() if sampler else None

    exchange_breakdown = None
    if trainer is not None and trainer.mode in ("p2p", "nvls", "hybrid"):
        evs = [[torch.cuda.Event(enable_timing=True) for _ in range(5)] for _ in range(10)]
        for i in range(10):
            trainer.profile_events = evs[i]
            step(*resident[i % N_BATCHES])
        trainer.profile_events = None
        barrier()
        names = ["sync_gradients_ms", "exchange_adam_kernel_ms", "sync_params_ms", "zero_grad_ms"]
        exchange_breakdown = {nm: sum(e[k].elapsed_time(e[k + 1]) for e in evs) / len(evs) for k, nm in enumerate(names)}

    fine_dist = None
    if world > 1 and not args.no_fine_stage:
        try:     # BASELINE configs[2] on N ranks (every rank takes part: collective inside)
            fine_dist = fine_stage_dist_bench(dev, rank, world)
        except Exception as e:
            fine_dist = {"error": str(e)[:200]}

    extra_configs = {}
    if not args.no_fine_stage:
        for key, fn in (("config_d_render_256", config_d_bench), ("config_e_dual_branch", config_e_bench)):
            try:     # parity-case configs measured for the record; never allowed to take the headline line down
                extra_configs[key] = fn(dev, rank, world)
            except Exception as e:
                extra_configs[key] = {"error": str(e)[:200]}
                try:
                    torch.cuda.empty_cache()
                except Exception:
                    pass

    if rank != 0:
        if trainer is not None:
            trainer.close()
        if world > 1:
            dist.destroy_process_group()
        return

    pk, pk_kind = peaks()
    V = GRID ** 3
    C_d, C_f = 1, 3
    sparse_bytes = 60 * RAYS + 96 * (C_d + C_f) * m_bar            # rays + gathers + scatters
    dense_bytes = 32 * V * (C_d + C_f)                             # Adam 28 B/elem + grad zeroing 4 B/elem
    step_bytes = sparse_bytes + dense_bytes
    march_gbs = sparse_bytes / (march_ms * 1e-3) / 1e9
    step_gbs = step_bytes / (ms_total / args.steps * 1e-3) / 1e9
    value = world * RAYS * args.steps / (ms_total * 1e-3)
    cfg = workload_config(world)
    exchange = None
    if world > 1:
        exchange = {
            "nvls": "ONE kernel: in-switch gradient reduction (multimem.ld_reduce) + masked Adam on 1/N of the "
                    "voxels + multicast parameter store (multimem.st); peer-memory flag barriers, no NCCL on the step",
            "hybrid": "ONE kernel: part of the slice through the switch (multimem.ld_reduce / multimem.st), the rest "
                      "through peer loads / stores, concurrently; masked Adam on 1/N of the voxels",
            "p2p": "ONE kernel: peer-load reduce-scatter + masked Adam on 1/N of the voxels + peer-store all-gather",
            "nccl": "NCCL all-reduce of the gradient records + full Adam on every rank"}[trainer.mode]
    line = {
        "metric": "train_rays_per_s", "value": value, "unit": "rays/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": cfg,
        "timing": {"blocks": len(block_ms), "steps_per_block": args.steps, "reported": "median block",
                   "ms_per_step_min": block_ms[0] / args.steps, "ms_per_step_max": block_ms[-1] / args.steps,
                   "timed_s": sum(block_ms) * 1e-3},
        "samples_per_ray": m_bar / RAYS,
        "roofline": {"bound": "hbm", "kernel": "dvgo_march_kernel<train> (gather + composite + loss + scatter)",
                     "achieved": march_gbs, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": march_gbs / pk["hbm_gbs"],
                     "peak_kind": pk_kind, "algorithmic_bytes_per_launch": sparse_bytes,
                     "ms_per_launch": march_ms, "traffic": MARCH_DRAM_TRAFFIC_BYTES,
                     "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of one launch in the committed ncu "
                                       "--set full capture " + MARCH_TRAFFIC_PROFILE + " (not re-measured by this run)"},
        "step_roofline": {"algorithmic_bytes_per_step": step_bytes, "achieved": step_gbs,
                          "frac": step_gbs / pk["hbm_gbs"], "march_ms": march_ms,
                          "exchange_plus_adam_ms": tail_ms},
        "e2e": {"value": world * RAYS * args.steps / (e2e_ms * 1e-3), "unit": "rays/s",
                "h2d_bytes_per_step": sum(t.numel() * 4 for t in host[0]), "d2h_bytes_per_step": 8,
                "ms_per_step": e2e_ms / args.steps, "blocks": len(e2e_blocks),
                "ms_per_step_min": e2e_blocks[0] / args.steps, "ms_per_step_max": e2e_blocks[-1] / args.steps,
                "api": {"pipeline": "FusedDVGO.host_pipeline(depth=2): submit(batch k+1) then result() of step k - the "
                                    "pinned-host copy of the next batch overlaps the step in flight, the loss of every "
                                    "step is read on the host",
                        "graph": "FusedDVGO.train_step_host: one CUDA-graph launch per step, stream synchronised every step",
                        "eager": "pinned copy in, train_step, pinned loss copy out, stream synchronised every step"}[e2e_mode],
                "other_call_patterns": e2e_side},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "render": {"metric": "render_mrays_per_s", "value": world * RAYS / (render_ms * 1e-3) / 1e6, "unit": "Mrays/s",
                   "ms_per_chunk": render_ms, "workload": "forward only (rgb, depth, alphainv_last), same 160^3 grid, "
                   "%d-ray chunks per GPU, device timed" % RAYS,
                   "achieved_gbs": (48 * RAYS + 32 * (C_d + C_f) * m_bar) / (render_ms * 1e-3) / 1e9,
                   "frac_of_hbm_peak": (48 * RAYS + 32 * (C_d + C_f) * m_bar) / (render_ms * 1e-3) / 1e9 / pk["hbm_gbs"],
                   "frac_note": "ALGORITHMIC bytes (B_render = 48 N + 32 C M, every corner gather counted, no cache reuse) "
                                "over the measured HBM copy peak: neighbouring samples of a ray share corners in L1 / L2, "
                                "so the DRAM traffic is ~3x smaller and this ratio can exceed 1",
                   "e2e": ({"value": world * RAYS / (render_host_ms * 1e-3) / 1e6, "unit": "Mrays/s",
                            "ms_per_chunk": render_host_ms, "h2d_bytes_per_chunk": 2 * RAYS * 12,
                            "d2h_bytes_per_chunk": RAYS * 16,
                            "what": "FusedDVGO.render_host: %d pinned host rays in, rgb + depth out to pinned host in one "
                                    "call; chunk copies on copy streams under the neighbouring chunks' kernels, one "
                                    "synchronisation per image" % n_img,
                            "synchronised_per_chunk": {"value": world * RAYS / (render_e2e_ms * 1e-3) / 1e6,
                                                       "ms_per_chunk": render_e2e_ms}}
                           if render_host_ms is not None else
                           {"value": world * RAYS / (render_e2e_ms * 1e-3) / 1e6, "unit": "Mrays/s",
                            "ms_per_chunk": render_e2e_ms, "h2d_bytes_per_chunk": 2 * RAYS * 12,
                            "d2h_bytes_per_chunk": RAYS * 16,
                            "what": "pinned host rays in, rgb + depth out to pinned host, synchronised per chunk"})},
    }
    if exchange:
        line["exchange"] = exchange
    if exchange_breakdown:
        line["step_roofline"]["exchange_breakdown"] = exchange_breakdown
    if world == 1:
        try:    # the same step with the total-variation term of SURVEY 8(d) switched on (dense mode, fused with Adam)
            tv_eng = dvgo.FusedDVGO(density.to(dev), k0.to(dev), XYZ_MIN, XYZ_MAX, stepsize=STEPSIZE, tv_density=5e-3,
                                    tv_k0=1e-2, tv_dense=True)
            for i in range(5):
                tv_eng.train_step(*resident[i % N_BATCHES])
            torch.cuda.synchronize()
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record()
            for i in range(30):
                tv_eng.train_step(*resident[i % N_BATCHES])
            t1.record()
            torch.cuda.synchronize()
            tv_ms = t0.elapsed_time(t1) / 30
            tv_bytes = step_bytes + 12 * V * (C_d + C_f)
            line["tv_on"] = {"workload": "same step with total_variation_add_grad (dense mode) fused into the Adam pass",
                             "train_rays_per_s": RAYS / (tv_ms * 1e-3), "ms_per_step": tv_ms,
                             "algorithmic_bytes_per_step": tv_bytes,
                             "frac_of_hbm_peak": tv_bytes / (tv_ms * 1e-3) / 1e9 / pk["hbm_gbs"]}
            del tv_eng
        except Exception as e:
            line["tv_on"] = {"error": str(e)[:200]}
    if world == 1 and not args.no_fine_stage:
        try:
            line["fine_stage"] = fine_stage_bench(dev)
        except Exception as e:   # an extra, never allowed to take the headline line down
            line["fine_stage"] = {"error": str(e)[:200]}
    if fine_dist is not None:
        line["fine_stage"] = fine_dist
    line.update(extra_configs)
    if world == 1 and not args.no_cpu_baseline:
        try:
            line["cpu_baseline"] = cpu_baseline()
        except Exception as e:
            line["cpu_baseline"] = {"error": str(e)[:200]}
    print(json.dumps(line), flush=True)
    try:
        if trainer is not None:
            trainer.close()
        if world > 1:
            dist.destroy_process_group()
    except Exception as e:      # the line is out: a teardown problem must not turn the run into a failure
        print("bench.py: teardown: %s" % str(e)[:200], file=sys.stderr)


def fine_stage_bench(dev):
    """BASELINE configs[2] on one rank: Voxurf SDF fine stage (160^3 sdf + 12-channel k0 + rgbnet 57-128x3-3 on
    tcgen05), 8192 rays / step, fwd + bwd + Adam.  Reported beside the headline, not instead of it."""
    import torch.nn as nn
    from poseprobe_b200 import ops, synthetic, voxurf
    torch.manual_seed(777)
    sdf = synthetic.make_sdf_grid(GRID, radius=0.5, noise=0.02)
    k0 = 0.1 * torch.randn(1, 12, GRID, GRID, GRID)
    net = nn.Sequential(nn.Linear(57, 128), nn.ReLU(inplace=True),
                        nn.Sequential(nn.Linear(128, 128), nn.ReLU(inplace=True)),
                        nn.Sequential(nn.Linear(128, 128), nn.ReLU(inplace=True)), nn.Linear(128, 3))
    nn.init.constant_(net[-1].bias, 0)
    ws, bs = ops.rgbnet_from_sequential(net)
    eng = voxurf.FusedVoxurf(sdf.to(dev), k0.to(dev), [w.detach() for w in ws], [b.detach() for b in bs], XYZ_MIN,
                             XYZ_MAX, stepsize=STEPSIZE, s_val=0.05, bg=0.0, sdf_alpha=10.0, sdf_beta=0.35,
                             posbase_pe=5, viewbase_pe=1, lr_sdf=1e-4, lr_k0=1e-2, lr_rgbnet=1e-3, lr_warp=1e-6)
    from poseprobe_b200 import synthetic as syn
    batches = []
    for b in range(4):
        o, d, v = syn.make_rays(RAYS, seed=777 + b)
        t = torch.rand(RAYS, 3, generator=torch.Generator().manual_seed(99 + b))
        batches.append(tuple(x.to(dev) for x in (o, d, v, t)))
    for i in range(4):
        eng.train_step(*batches[i % 4])
    torch.cuda.synchronize()
    m = eng.n_samples()
    steps = 20
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        eng.train_step(*batches[i % 4])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    V = GRID ** 3
    step_bytes = 60 * RAYS + 96 * (4 + 12) * m + 32 * V * 13
    mlp_flops = 3 * 2.0 * m * (57 * 128 + 2 * 128 * 128 + 128 * 3)
    pk, _ = peaks()
    # render with the implicit deformation (Voxurf.inference): warp network 3-128x4-4 + Jacobian on every in-box sample
    dims = [3, 128, 128, 128, 128, 4]
    wws = [torch.randn(dims[i + 1], dims[i]) * (2.0 / dims[i]) ** 0.5 for i in range(5)]
    wws[-1] = 0.005 * torch.randn(4, 128)
    eng.set_warp_network(wws, [0.1 * torch.randn(dims[i + 1]) for i in range(5)], 2.0)
    o, d, v, _ = batches[0]
    for _ in range(3):
        eng.render(o, d, v, use_deform=True)
    e0.record()
    for _ in range(10):
        eng.render(o, d, v, use_deform=True)
    e1.record()
    torch.cuda.synchronize()
    ms_r = e0.elapsed_time(e1) / 10
    n_in = eng._bufs["dcounter"].tolist()[0]
    warp_flops = 4 * n_in * 2.0 * (16 * 128 + 3 * 128 * 128 + 128 * 16)
    render_deform = {"workload": "Voxurf.inference: warp network (3-128x4-4, with input Jacobian, fp16 tcgen05) on every "
                                 "in-box sample + warped SDF lookup + rgbnet, %d-ray chunks, forward only" % RAYS,
                     "render_mrays_per_s": RAYS / (ms_r * 1e-3) / 1e6, "ms_per_chunk": ms_r,
                     "inbox_samples_per_ray": n_in / RAYS, "warp_network_tflop_per_chunk": warp_flops / 1e12}
    # the reference's LIVE train step (lib/recon_scene.py:603): through the deformation network with the full object
    # loss (photometric with mask, entropy, mask BCE, k0 TV, Eikonal over all in-box samples, three deformation priors),
    # SDF grid frozen, dense reference Adam on feature grid / rgbnet / deformation network
    eng.set_loss_weights(w_entropy=0.01, w_mask=0.1, w_tv_k0=0.01, w_eikonal=1.0, w_deform=0.05)
    msk = [(torch.rand(RAYS, generator=torch.Generator().manual_seed(5 + b)) > 0.2).float().to(dev) for b in range(4)]
    for i in range(3):
        eng.train_step(*batches[i % 4], use_deform=True, mask=msk[i % 4])
    torch.cuda.synchronize()
    e0.record()
    for i in range(10):
        eng.train_step(*batches[i % 4], use_deform=True, mask=msk[i % 4])
    e1.record()
    torch.cuda.synchronize()
    ms_d = e0.elapsed_time(e1) / 10
    n_in = eng._bufs["dcounter"].tolist()[0]
    m_c = eng.n_samples()
    # the deformation network's backward alone, on the buffers the last step left behind
    import ctypes
    from poseprobe_b200._lib import check, lib, ptr, stream as cur_stream
    b = eng._bufs
    wws_k, wbs_k, _ = eng.warp
    arr = lambda ts: (ctypes.c_void_p * len(ts))(*[t.data_ptr() for t in ts])  # noqa: E731
    scratch_w = [torch.zeros_like(t) for t in eng.g_ww]
    scratch_b = [torch.zeros_like(t) for t in eng.g_wb]

    def warp_bwd():
        check(lib.pp_warpnet_backward_packed(ptr(b["drows"]), ptr(b["g_dout"]), 4 * b["cap"], ptr(b["dcounter"][2:]), 5,
                                             128, arr(wws_k), arr(wbs_k), arr(scratch_w), arr(scratch_b),
                                             ptr(eng._warp_work), cur_stream()), "warpnet_backward")
    warp_bwd()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(5):
        warp_bwd()
    e1.record()
    torch.cuda.synchronize()
    ms_wb = e0.elapsed_time(e1) / 5
    fwd_flops_row = 2.0 * (16 * 128 + 3 * 128 * 128 + 128 * 16)
    tflops_pk = pk.get("bf16_tflops_sustained", pk["bf16_tflops"])
    train_deform = {
        "workload": "Voxurf.forward(use_deform=True) train step (recon_scene.py:603): warp network fwd+Jacobian and its "
                    "tcgen05 backward on every in-box sample, rgbnet fwd/bwd, full object loss (lib/losses.py:34-74), "
                    "dense Adam; %d rays/step" % RAYS,
        "train_rays_per_s": RAYS / (ms_d * 1e-3), "ms_per_step": ms_d, "inbox_samples_per_ray": n_in / RAYS,
        "composited_samples_per_ray": m_c / RAYS, "stalled": int(lib.pp_warpnet_backward_stalled()),
        "warp_backward": {"ms": ms_wb, "rows": 4 * n_in,
                          "algorithmic_tflop": 3 * fwd_flops_row * 4 * n_in / 1e12,
                          "algorithmic_tflops_per_s": 3 * fwd_flops_row * 4 * n_in / 1e12 / (ms_wb * 1e-3),
                          "frac_of_tensor_peak": 3 * fwd_flops_row * 4 * n_in / 1e12 / (ms_wb * 1e-3) / tflops_pk,
                          "note": "recompute + dH + dW = 3x the forward flops; tensor-pipe %% from ncu in profiles/"}}
    return {"render_deform": render_deform, "train_deform": train_deform,
            "workload": "Voxurf SDF fine stage (BASELINE configs[2], one rank): 160^3 sdf + 12-ch k0 + rgbnet "
                        "57-128-128-128-3 (tcgen05), 8192 rays/step, fwd+bwd+Adam",
            "train_rays_per_s": RAYS / (ms * 1e-3), "ms_per_step": ms, "composited_samples_per_ray": m / RAYS,
            "M_c_over_M": 1.0, "algorithmic_bytes_per_step": step_bytes, "mlp_algorithmic_tflop_per_step": mlp_flops / 1e12,
            "roofline_ms": {"hbm": step_bytes / (pk["hbm_gbs"] * 1e9) * 1e3,
                            "tensor": mlp_flops / (pk.get("bf16_tflops_sustained", pk["bf16_tflops"]) * 1e12) * 1e3},
            "frac_of_hbm_plus_tensor_bound": (step_bytes / (pk["hbm_gbs"] * 1e9) * 1e3 + mlp_flops / (
                pk.get("bf16_tflops_sustained", pk["bf16_tflops"]) * 1e12) * 1e3) / ms}


def fine_stage_dist_bench(dev, rank, world):
    """BASELINE configs[2] on N ranks: Voxurf fine stage ray-sharded over the ranks (8192 rays per rank), ONE flat
    gradient all-reduce over NVLink per step, identical optimiser step on every replica."""
    import torch.distributed as dist
    import torch.nn as nn
    from poseprobe_b200 import ops, synthetic, voxurf
    torch.manual_seed(777)
    sdf = synthetic.make_sdf_grid(GRID, radius=0.5, noise=0.02)
    k0 = 0.1 * torch.randn(1, 12, GRID, GRID, GRID)
    net = nn.Sequential(nn.Linear(57, 128), nn.ReLU(inplace=True),
                        nn.Sequential(nn.Linear(128, 128), nn.ReLU(inplace=True)),
                        nn.Sequential(nn.Linear(128, 128), nn.ReLU(inplace=True)), nn.Linear(128, 3))
    nn.init.constant_(net[-1].bias, 0)
    ws, bs = ops.rgbnet_from_sequential(net)
    tr = voxurf.DistributedVoxurf(sdf.to(dev), k0.to(dev), [w.detach() for w in ws], [b.detach() for b in bs], XYZ_MIN,
                                  XYZ_MAX, stepsize=STEPSIZE, s_val=0.05, bg=0.0, sdf_alpha=10.0, sdf_beta=0.35,
                                  posbase_pe=5, viewbase_pe=1, lr_sdf=1e-4, lr_k0=1e-2, lr_rgbnet=1e-3, pose_grad=True)
    batches = []
    for b in range(4):
        o, d, v = synthetic.make_rays(RAYS, seed=777 + 1000 * rank + b)
        t = torch.rand(RAYS, 3, generator=torch.Generator().manual_seed(99 + 1000 * rank + b))
        batches.append(tuple(x.to(dev) for x in (o, d, v, t)))
    for i in range(3):
        tr.train_step(*batches[i % 4])
    dist.barrier()
    torch.cuda.synchronize()
    steps = 20
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        tr.train_step(*batches[i % 4])
    e1.record()
    dist.barrier()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms)
    mode = tr.mode
    tr.close()
    del tr
    return {"workload": "Voxurf SDF fine stage + per-ray pose gradient (BASELINE configs[2]): 160^3 sdf + 12-ch k0 + "
                        "rgbnet 57-128-128-128-3 (tcgen05), 8192 rays/step/rank, fwd+bwd+Adam, ray-sharded over %d ranks; "
                        "exchange: %s" % (world, {
                            "nvls": "feature records through ONE in-switch reduce-scatter + sharded Adam + multicast "
                                    "kernel, SDF / rgbnet gradients summed in place through the switch, flag barriers "
                                    "in peer memory (no NCCL on the step)",
                            "nccl": "one flat NCCL all-reduce of the gradient buffer, full Adam on every rank"}[mode]),
            "train_rays_per_s": world * RAYS / (ms * 1e-3), "ms_per_step": ms, "n_gpus": world}


def _fine_net():
    import torch.nn as nn
    from poseprobe_b200 import ops
    net = nn.Sequential(nn.Linear(57, 128), nn.ReLU(inplace=True),
                        nn.Sequential(nn.Linear(128, 128), nn.ReLU(inplace=True)),
                        nn.Sequential(nn.Linear(128, 128), nn.ReLU(inplace=True)), nn.Linear(128, 3))
    nn.init.constant_(net[-1].bias, 0)
    ws, bs = ops.rgbnet_from_sequential(net)
    return [w.detach() for w in ws], [b.detach() for b in bs]


def _sphere_sdf(G, dev, seed):
    """distance-to-sphere SDF (r = 0.5) + 0.02 N(0,1) on [-1,1]^3, generated on the device (the 320^3 grids of
    config E are too large to draw on the host within the bench's time budget)."""
    gen = torch.Generator(device=dev).manual_seed(seed)
    lin = torch.linspace(-1, 1, G, device=dev)
    x, y, z = torch.meshgrid(lin, lin, lin, indexing="ij")
    sdf = (x * x + y * y + z * z).sqrt() - 0.5 + 0.02 * torch.randn(G, G, G, device=dev, generator=gen)
    k0 = 0.1 * torch.randn(1, 12, G, G, G, device=dev, generator=gen)
    return sdf[None, None].contiguous(), k0


def config_d_bench(dev, rank, world):
    """BASELINE configs[3]: full-image render throughput, 800x800 = 640 000 rays through a 256^3 SDF grid + 12-channel
    feature grid + rgbnet (tcgen05), forward only, in 8192-ray chunks; the chunks of the image are dealt round-robin to
    the ranks (no communication: every rank writes its own rows), time = max over ranks."""
    import torch.distributed as dist
    from poseprobe_b200 import synthetic, voxurf
    G, chunk = 256, 8192
    n_rays = 800 * 800
    n_chunks = (n_rays + chunk - 1) // chunk
    sdf, k0 = _sphere_sdf(G, dev, 4242)
    ws, bs = _fine_net()
    eng = voxurf.FusedVoxurf(sdf, k0, ws, bs, XYZ_MIN, XYZ_MAX, stepsize=STEPSIZE, s_val=0.05, bg=0.0, sdf_alpha=10.0,
                             sdf_beta=0.35, posbase_pe=5, viewbase_pe=1)
    del sdf, k0
    pool = []
    for c in range(8):
        o, d, v = synthetic.make_rays(chunk, seed=900 + c)
        pool.append(tuple(x.to(dev) for x in (o, d, v)))
    mine = list(range(rank, n_chunks, world))
    sizes = [min(chunk, n_rays - c * chunk) for c in mine]
    out_rgb = torch.empty(chunk, 3, pin_memory=True)
    out_depth = torch.empty(chunk, pin_memory=True)

    def image(to_host):
        for c, n in zip(mine, sizes):
            o, d, v = pool[c % 8]
            res = eng.render(o[:n], d[:n], v[:n], want_depth=True)
            if to_host:
                out_rgb[:n].copy_(res[0], non_blocking=True)
                out_depth[:n].copy_(res[3], non_blocking=True)

    def timed(to_host, reps):
        image(to_host)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            image(to_host)
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / reps], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)
    ms = timed(False, 3)
    ms_host = timed(True, 3)
    m_per_ray = eng.n_samples() / sizes[-1] if sizes else 0.0
    del eng
    torch.cuda.empty_cache()
    return {"workload": "BASELINE configs[3]: 800x800 image (640 000 rays, %d chunks of 8192 dealt round-robin to %d "
                        "rank(s)), 256^3 SDF + 12-ch k0 + rgbnet 57-128-128-128-3 (tcgen05), forward only" % (n_chunks, world),
            "render_mrays_per_s": n_rays / (ms * 1e-3) / 1e6, "ms_per_image": ms, "n_gpus": world, "scaling": "strong",
            "composited_samples_per_ray": m_per_ray,
            "e2e": {"render_mrays_per_s": n_rays / (ms_host * 1e-3) / 1e6, "ms_per_image": ms_host,
                    "what": "rgb + depth of every chunk copied to pinned host memory inside the timed region"}}


def config_e_bench(dev, rank, world):
    """BASELINE configs[4]: dual-branch step, object 160^3 + scene 320^3 fine-stage grids (synthetic: the reference's
    scene branch is an MLP NeRF), 8192 rays per rank through each, fwd + bwd + Adam; ray-sharded over the
    ranks with one flat gradient all-reduce per grid and step (N = 8: 65 536 rays per step)."""
    import torch.distributed as dist
    from poseprobe_b200 import synthetic, voxurf
    ws, bs = _fine_net()
    engines = []
    for G, seed in ((160, 11), (320, 12)):
        sdf, k0 = _sphere_sdf(G, dev, seed)
        kw = dict(stepsize=STEPSIZE, s_val=0.05, bg=0.0, sdf_alpha=10.0, sdf_beta=0.35, posbase_pe=5, viewbase_pe=1,
                  lr_sdf=1e-4, lr_k0=1e-2, lr_rgbnet=1e-3)
        if world > 1:
            engines.append(voxurf.DistributedVoxurf(sdf, k0, ws, bs, XYZ_MIN, XYZ_MAX, **kw))
        else:
            engines.append(voxurf.FusedVoxurf(sdf, k0, ws, bs, XYZ_MIN, XYZ_MAX, **kw))
        del sdf, k0
    batches = []
    for b in range(4):
        o, d, v = synthetic.make_rays(RAYS, seed=3000 + 100 * rank + b)
        t = torch.rand(RAYS, 3, generator=torch.Generator().manual_seed(77 + 100 * rank + b))
        batches.append(tuple(x.to(dev) for x in (o, d, v, t)))

    def step(i):
        for e in engines:
            e.train_step(*batches[i % 4])
    for i in range(3):
        step(i)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    steps = 8
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        step(i)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms)
    gb = torch.cuda.memory_allocated(dev) / 1e9
    del engines
    torch.cuda.empty_cache()
    return {"workload": "BASELINE configs[4]: dual-branch fine-stage train step, object 160^3 + scene 320^3 grids (sdf + "
                        "12-ch k0 + rgbnet each), %d rays/rank through each, fwd+bwd+Adam, %d rank(s)" % (RAYS, world),
            "train_rays_per_s": world * RAYS / (ms * 1e-3), "global_rays_per_step": world * RAYS, "ms_per_step": ms,
            "n_gpus": world, "resident_GB_per_rank": gb}


def cpu_baseline():
    """The reference's pure-PyTorch CPU path on a bounded number of steps of the same workload (rank 0, N=1)."""
    from oracle import dvgo_cpu
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    batches, synthetic = make_batches_cpu(1, RAYS, 0)
    density, k0 = synthetic.make_dvgo_grids(GRID)
    base = dvgo_cpu.DenseBaseline(density, k0, XYZ_MIN, XYZ_MAX, STEPSIZE)
    o, d, tgt = batches[0]
    base.train_step(o, d, tgt)
    t0 = time.perf_counter()
    steps = 0
    while steps < 3 or (time.perf_counter() - t0 < 10 and steps < 20):
        base.train_step(o, d, tgt)
        steps += 1
    dt = time.perf_counter() - t0
    out = {"value": RAYS * steps / dt, "unit": "rays/s", "cores": cores, "kind": "port",
           "sample": "%d steps of the full %d-ray batch (lib/dvgo_ori.py dense formulation, oracle/dvgo_cpu.py)"
                     % (steps, RAYS)}
    # BASELINE configs[0] (the reference's own CPU-runnable case): 64^3 grid, 4096-ray step, same formulation
    density, k0 = synthetic.make_dvgo_grids(64)
    o, d, _ = synthetic.make_rays(4096)
    tgt = torch.rand(4096, 3, generator=torch.Generator().manual_seed(99))
    base = dvgo_cpu.DenseBaseline(density, k0, XYZ_MIN, XYZ_MAX, STEPSIZE)
    base.train_step(o, d, tgt)
    t0 = time.perf_counter()
    steps = 0
    while steps < 3 or (time.perf_counter() - t0 < 4 and steps < 20):
        base.train_step(o, d, tgt)
        steps += 1
    dt = time.perf_counter() - t0
    out["config_a"] = {"workload": "BASELINE configs[0]: 64^3 density+k0(3) grid, 4096-ray fwd+bwd+Adam step, CPU",
                       "value": 4096 * steps / dt, "unit": "rays/s", "ms_per_step": dt / steps * 1e3, "steps": steps}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", dest="no_cpu_baseline", action="store_true")
    ap.add_argument("--no-fine-stage", dest="no_fine_stage", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product has no CPU fallback)")
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
