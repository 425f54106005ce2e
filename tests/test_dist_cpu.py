"""CPU tier (gloo, world_size 2): the voxel-sharded optimiser protocol of poseprobe_b200/dist.py --
slice bounds, "sum the per-rank gradients of my slice, masked-Adam it once, publish it to every replica"
-- equals one process applying masked Adam to the summed gradient.  The arithmetic is the C oracle's;
the CUDA kernel itself is covered by tests/dist_check.py on GPUs."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, out):
    from oracle import oracle
    from poseprobe_b200.dist import shard_bounds
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(5)
    param = torch.randn(n, 4, generator=g)                       # identical replicas
    grads = []
    for r in range(world):                                       # every rank knows its own gradient only
        gr = torch.randn(n, 4, generator=g)
        gr[torch.rand(n, generator=g) < 0.4] = 0
        grads.append(gr)
    mine = grads[rank]
    b, e = shard_bounds(n, world, rank)
    # reduce-scatter: gather every rank's slice of the gradient (peer loads in the kernel)
    pieces = [torch.zeros(n, 4) for _ in range(world)]
    dist.all_gather(pieces, mine)
    gsum = pieces[0][b:e].clone()
    for r in range(1, world):
        gsum += pieces[r][b:e]                                   # fixed rank order, like the kernel
    m, v = torch.zeros(e - b, 4), torch.zeros(e - b, 4)
    slice_p = param[b:e].clone()
    for ch, lr in enumerate((0.1, 0.05, 0.05, 0.05)):            # density / k0 learning rates
        p_c, g_c = slice_p[:, ch].contiguous(), gsum[:, ch].contiguous()
        m_c, v_c = m[:, ch].contiguous(), v[:, ch].contiguous()
        oracle.adam_upd(1, p_c, g_c, m_c, v_c, None, 1, 0.9, 0.99, lr, 1e-8)
        slice_p[:, ch] = p_c
    # all-gather: publish my slice to every replica (peer stores in the kernel)
    padded = torch.zeros(shard_bounds(n, world, 0)[1], 4)
    padded[:e - b] = slice_p
    allp = [torch.zeros_like(padded) for _ in range(world)]
    dist.all_gather(allp, padded)
    new_param = torch.cat([allp[r][:shard_bounds(n, world, r)[1] - shard_bounds(n, world, r)[0]] for r in range(world)])
    # single-process reference
    ref = param.clone()
    total = grads[0].clone()
    for r in range(1, world):
        total += grads[r]
    for ch, lr in enumerate((0.1, 0.05, 0.05, 0.05)):
        p_c, g_c = ref[:, ch].contiguous(), total[:, ch].contiguous()
        oracle.adam_upd(1, p_c, g_c, torch.zeros(n), torch.zeros(n), None, 1, 0.9, 0.99, lr, 1e-8)
        ref[:, ch] = p_c
    ok = torch.equal(new_param, ref)
    out[rank] = bool(ok)
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [1000, 777])
def test_sharded_optimizer_protocol_gloo(n):
    world = 2
    out = mp.Manager().dict()
    mp.spawn(_worker, args=(world, _free_port(), n, out), nprocs=world, join=True)
    assert all(out[r] for r in range(world)), dict(out)


def test_shard_bounds_partition_matches_c_abi():
    from poseprobe_b200 import _lib
    from poseprobe_b200.dist import shard_bounds
    for n in (1, 255, 256, 4096000, 160 ** 3, 13):
        for world in (1, 2, 4, 8):
            s = int(_lib.lib.pp_voxel_shard_size(n, world))
            covered = 0
            for r in range(world):
                b, e = shard_bounds(n, world, r)
                assert b == min(s * r, n) and e == min(b + s, n)
                assert b == covered or b == n
                covered = max(covered, e)
            assert covered == n


def _fine_worker(rank, world, port, n_vox, out):
    """DistributedVoxurf(mode='nvls') protocol on the CPU: the feature records (3 float4 records per voxel) are
    reduce-scattered, updated by the slice's owner with the reference's dense Adam and all-gathered; the SDF / rgbnet
    gradients are all-reduced in place and every rank applies the same update.  Must equal ONE process applying dense
    Adam to the summed gradients, and leave identical replicas."""
    from oracle import oracle
    from poseprobe_b200.dist import shard_bounds
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(9)
    n_rec = 3 * n_vox
    k0 = torch.randn(n_rec, 4, generator=g)
    small = torch.randn(4 * n_vox + 37, generator=g)              # SDF gradient records | rgbnet gradients (flat)
    g_k0 = [torch.randn(n_rec, 4, generator=g) * (torch.rand(n_rec, 1, generator=g) > 0.3) for _ in range(world)]
    g_small = [torch.randn(small.shape, generator=g) for _ in range(world)]
    b, e = shard_bounds(n_rec, world, rank)
    m, v = torch.zeros(max(e - b, 0), 4), torch.zeros(max(e - b, 0), 4)
    params, smalls = k0.clone(), small.clone()
    m_s, v_s = torch.zeros_like(small), torch.zeros_like(small)
    m_ref, v_ref = torch.zeros_like(k0), torch.zeros_like(k0)
    ms_ref, vs_ref = torch.zeros_like(small), torch.zeros_like(small)
    ref_k0, ref_small = k0.clone(), small.clone()
    for step in (1, 2):
        # feature records: reduce-scatter (fixed rank order) + dense Adam on my slice + all-gather
        pieces = [torch.zeros(n_rec, 4) for _ in range(world)]
        dist.all_gather(pieces, g_k0[rank] * step)
        gsum = pieces[0][b:e].clone()
        for r in range(1, world):
            gsum += pieces[r][b:e]
        mine = params[b:e].clone()
        oracle.python_adam_step(mine, gsum, m, v, step, 0.9, 0.99, 1e-2, 1e-8)
        pad = shard_bounds(n_rec, world, 0)[1]
        padded = torch.zeros(pad, 4)
        padded[:e - b] = mine
        allp = [torch.zeros_like(padded) for _ in range(world)]
        dist.all_gather(allp, padded)
        params = torch.cat([allp[r][:shard_bounds(n_rec, world, r)[1] - shard_bounds(n_rec, world, r)[0]]
                            for r in range(world)])
        # small gradients: in-place all-reduce, identical update everywhere
        gs = g_small[rank] * step
        dist.all_reduce(gs)
        oracle.python_adam_step(smalls, gs, m_s, v_s, step, 0.9, 0.99, 1e-3, 1e-8)
        # one process on the summed gradients
        tot_k0, tot_s = sum(x * step for x in g_k0), sum(x * step for x in g_small)
        oracle.python_adam_step(ref_k0, tot_k0, m_ref, v_ref, step, 0.9, 0.99, 1e-2, 1e-8)
        oracle.python_adam_step(ref_small, tot_s, ms_ref, vs_ref, step, 0.9, 0.99, 1e-3, 1e-8)
    same = [params.clone() for _ in range(world)]
    dist.all_gather(same, params)
    out[rank] = bool(torch.allclose(params, ref_k0, rtol=1e-6, atol=1e-7) and
                     torch.allclose(smalls, ref_small, rtol=1e-6, atol=1e-7) and
                     all(torch.equal(same[0], s) for s in same))
    dist.destroy_process_group()


@pytest.mark.parametrize("n_vox", [343, 1000])
def test_fine_stage_exchange_protocol_gloo(n_vox):
    world = 2
    out = mp.Manager().dict()
    mp.spawn(_fine_worker, args=(world, _free_port(), n_vox, out), nprocs=world, join=True)
    assert all(out[r] for r in range(world)), dict(out)
