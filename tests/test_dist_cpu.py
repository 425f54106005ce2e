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
