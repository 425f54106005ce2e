"""Multi-GPU training of the fused DVGO path: ray-sharded replicas, voxel-sharded optimiser.

One process per GPU (torchrun), ``torch.distributed`` (NCCL) for rendezvous, handle exchange and the
two tiny cross-rank stream barriers of a step.  The data path is NOT an NCCL collective: every rank's
parameter and gradient records live in IPC-shared device memory, and ``pp_voxel_adam_exchange`` (one
kernel per rank) reduces the gradient slices over NVLink with peer loads, applies masked Adam once per
voxel and writes the result into every replica with peer stores.

Step on every rank (every item is one launch on the rank's stream; no NCCL call on the critical path):
    march + loss + scatter into the local gradient records          (no communication)
    pp_peer_sync(loss payload)          -> "all gradients complete" barrier; delivers the global loss
    pp_voxel_adam_exchange              -> reduce-scatter + Adam(1/world of the voxels) + all-gather
    pp_peer_sync                        -> "all slices written, all peers done reading" barrier
    zero the local gradient records

``mode="nccl"`` keeps the plain formulation (all-reduce of the gradient records, full Adam on every rank)
as the comparison baseline.  The reference itself is single-GPU; the N-rank result equals
the 1-rank result on the concatenated batch up to fp32 summation order.
"""
import ctypes

import torch
import torch.distributed as dist

from . import dvgo
from ._lib import check, lib, ptr, stream

_IPC_BYTES = 64


def shard_bounds(n_voxels, world, rank):
    """[begin, end) of the voxel slice rank owns (mirror of pp_voxel_shard_size)."""
    s = (n_voxels + world - 1) // world
    s = (s + 255) // 256 * 256
    begin = min(s * rank, n_voxels)
    return begin, min(begin + s, n_voxels)


class _DeviceBuffer:
    """float32 view of raw device memory for torch.as_tensor (CUDA array interface)."""

    def __init__(self, address, n_floats, owner=None):
        self.owner = owner
        self.__cuda_array_interface__ = {"shape": (n_floats,), "typestr": "<f4", "data": (address, False),
                                         "version": 2}


class SharedRecords:
    """[n,4] fp32 records in cudaMalloc'ed memory that peers can map (CUDA IPC)."""

    def __init__(self, n_records, device, zero=False):
        self.device = device
        self.n = n_records
        p = ctypes.c_void_p()
        self.handle = (ctypes.c_ubyte * _IPC_BYTES)()
        with torch.cuda.device(device):
            check(lib.pp_shared_alloc(n_records * 16, ctypes.byref(p), self.handle), "shared_alloc")
        self.address = p.value
        self.tensor = torch.as_tensor(_DeviceBuffer(self.address, n_records * 4, self), device=device).view(n_records, 4)
        if zero:
            self.tensor.zero_()
        self.peers = []

    def exchange(self, group, world, rank):
        """all-gather the IPC handles and map every peer's buffer; returns the pointer table."""
        mine = torch.tensor(list(bytes(self.handle)), dtype=torch.uint8, device=self.device)
        allh = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allh, mine, group=group)
        table = (ctypes.c_void_p * world)()
        for r in range(world):
            if r == rank:
                table[r] = self.address
                continue
            h = (ctypes.c_ubyte * _IPC_BYTES)(*allh[r].cpu().tolist())
            p = ctypes.c_void_p()
            with torch.cuda.device(self.device):
                check(lib.pp_shared_open(h, ctypes.byref(p)), "shared_open")
            self.peers.append(p.value)
            table[r] = p.value
        return table

    def close(self):
        with torch.cuda.device(self.device):
            for p in self.peers:
                lib.pp_shared_close(ctypes.c_void_p(p))
            self.peers = []


class SymmRecords:
    """[n,4] fp32 records in torch symmetric memory: peer pointers AND an NVSwitch multicast address.
    torch.distributed is plumbing here (allocation + handle exchange); the kernels that use the addresses
    are this library's."""

    def __init__(self, n_records, device, group, world, zero=False):
        import torch.distributed._symmetric_memory as symm_mem
        self.device = device
        self.tensor = symm_mem.empty((n_records, 4), dtype=torch.float32, device=device)
        if zero:
            self.tensor.zero_()
        name = (group or dist.group.WORLD).group_name
        self.hdl = symm_mem.rendezvous(self.tensor, name)
        self.address = self.tensor.data_ptr()
        self.multicast_ptr = int(self.hdl.multicast_ptr)
        self.table = (ctypes.c_void_p * world)(*[int(p) for p in self.hdl.buffer_ptrs])

    def exchange(self, group, world, rank):
        return self.table

    def close(self):
        pass


def nvls_available():
    try:
        import torch.distributed._symmetric_memory  # noqa: F401
        return True
    except Exception:
        return False


class DistributedDVGO:
    """FusedDVGO replicated over the ranks of a process group."""

    def __init__(self, density, k0, xyz_min, xyz_max, stepsize, group=None, mode="auto", **kw):
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        auto = mode == "auto"
        if mode == "auto":
            # measured on 8xB200 (profiles/r1_results.md): in-switch reduction wins at 8 ranks, plain peer
            # loads/stores at 2 and 4 (their link traffic 2(w-1)/w is below NVLS's (w+1)/w there)
            mode = "nvls" if (self.world >= 8 and nvls_available()) else "p2p"
        self.mode = mode if self.world > 1 else "single"
        import os
        self.fused_sync = os.environ.get("PP_XCHG_FUSED_SYNC", "1") != "0"
        dev = density.device
        kw = dict(kw)
        self.mc_fraction = float(os.environ.get("PP_XCHG_MC_FRACTION", "0.7"))
        self.symm = self.mode in ("nvls", "hybrid")       # records in torch symmetric memory (multicast address)
        if self.mode in ("p2p", "nvls", "hybrid"):
            # the fused exchange kernels have no total-variation stencil and exist for 1, 2, 4, 8 ranks
            if any(kw.get(k, 0.0) for k in ("tv_density", "tv_k0")):
                if not auto:
                    raise RuntimeError("DistributedDVGO(mode=%r): total variation is only implemented in the "
                                       "'nccl' mode (all-reduce + the single-GPU TV + Adam pass)" % self.mode)
                self.mode = "nccl"
            elif self.world not in (2, 4, 8):
                if not auto:
                    raise RuntimeError("DistributedDVGO(mode=%r): world size %d not supported (2, 4, 8)"
                                       % (self.mode, self.world))
                self.mode = "nccl"
        kw["w_rgb"] = kw.get("w_rgb", 1.0) / self.world          # global mean over all ranks' rays
        for k in ("w_entropy_last", "w_rgbper"):
            if k in kw:
                kw[k] = kw[k] / self.world
        kw["mse_scale"] = 1.0 / self.world
        n = density.numel()
        # two gradient buffers: while step k+1 marches into one, the other (already consumed by step k's
        # exchange) is zeroed on a side stream, off the critical path
        if self.symm:
            ok = True
            try:
                self.vox_mem = SymmRecords(n, dev, group, self.world)
                self.grad_mems = [SymmRecords(n, dev, group, self.world) for _ in range(2)]
                ok = bool(self.vox_mem.multicast_ptr and all(g.multicast_ptr for g in self.grad_mems))
            except Exception as e:      # no symmetric memory / multicast on this system
                ok = False
                err = e
            # every rank must take the same path: agree on the outcome before continuing
            flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
            if not int(flag):
                if not auto:
                    raise RuntimeError("DistributedDVGO(mode=%r): no multicast address "
                                       "(NVSwitch/NVLS or torch symmetric memory not available)" % self.mode)
                self.mode = "p2p"       # auto mode: fall back to peer loads / stores over NVLink
                self.symm = False
                self.vox_mem = self.grad_mems = None
        if self.mode == "p2p":
            self.vox_mem = SharedRecords(n, dev)
            self.grad_mems = [SharedRecords(n, dev) for _ in range(2)]
        if self.mode in ("p2p", "nvls", "hybrid"):
            b, e = shard_bounds(n, self.world, self.rank)
            self.eng = dvgo.FusedDVGO(density, k0, xyz_min, xyz_max, stepsize, vox_buf=self.vox_mem.tensor,
                                      grad_buf=self.grad_mems[0].tensor, moment_records=max(e - b, 1), **kw)
            self.grad_mems[1].tensor.zero_()
            n_sync = int(lib.pp_peer_sync_bytes()) // 16
            self.sync_mem = (SymmRecords(n_sync, dev, group, self.world, zero=True) if self.symm
                             else SharedRecords(n_sync, dev, zero=True))
            torch.cuda.synchronize(dev)
            self.vox_ptrs = self.vox_mem.exchange(group, self.world, self.rank)
            self.grad_ptrs = [g.exchange(group, self.world, self.rank) for g in self.grad_mems]
            self.cur = 0
            self.side = torch.cuda.Stream(device=dev)
            self.zeroed = [torch.cuda.Event(), torch.cuda.Event()]
            self.consumed = torch.cuda.Event()
            self.sync_ptrs = self.sync_mem.exchange(group, self.world, self.rank)
            self.epoch = 0
            self.loss_global = torch.zeros(2, dtype=torch.float32, device=dev)
        else:
            self.eng = dvgo.FusedDVGO(density, k0, xyz_min, xyz_max, stepsize, **kw)
        self.token = torch.zeros(1, dtype=torch.float32, device=dev)
        self.profile_events = None      # optional: 5 CUDA events (sync1 start, exchange start/end, sync2 end, zero end)
        dist.barrier(group=group)

    def train_step(self, rays_o, rays_d, target, events=None):
        """One optimisation step on this rank's ray batch; returns [global loss, global mse] (device).
        `events` (optional): three CUDA events recorded before the march, after it, and at the end."""
        try:
            return self._train_step(rays_o, rays_d, target, events)
        finally:
            if events is not None:
                events[2].record()

    def _train_step(self, rays_o, rays_d, target, events):
        eng = self.eng
        eng.loss_buf.zero_()
        if self.mode in ("p2p", "nvls", "hybrid"):
            eng.grad = self.grad_mems[self.cur].tensor
            torch.cuda.current_stream().wait_event(self.zeroed[self.cur])     # no-op until first recorded
        if events is not None:
            events[0].record()
        eng.backward_only(rays_o, rays_d, target)
        if events is not None:
            events[1].record()
        if self.mode == "single":
            eng.optimizer_step()
            return eng.loss_buf
        if self.mode == "nccl":
            dist.all_reduce(eng.loss_buf, group=self.group)
            dist.all_reduce(eng.grad, group=self.group)
            eng.optimizer_step()
            return eng.loss_buf
        b1, b2, lr_d, lr_k, eps = eng.hp
        eng.step += 1
        prof = self.profile_events
        with torch.cuda.device(eng.device):
            s = stream()
            if prof is not None:
                prof[0].record()
            # barrier: every rank finished scattering; the local [loss, mse] rides along and comes back summed.
            # NVLS mode fuses it into the exchange kernel's prologue (one launch less on the critical path).
            self.epoch += 1
            fused = self.symm and self.fused_sync
            if not fused:
                check(lib.pp_peer_sync(self.sync_ptrs, self.world, self.rank, self.epoch, ptr(eng.loss_buf),
                                       ptr(self.loss_global), s), "peer_sync")
            if prof is not None:
                prof[1].record()
            if self.mode == "nvls":
                check(lib.pp_voxel_adam_exchange_mc(ptr(eng.vox), self.vox_mem.multicast_ptr,
                                                    self.grad_mems[self.cur].multicast_ptr, ptr(eng.exp_avg),
                                                    ptr(eng.exp_avg_sq), eng.n_voxels, self.world, self.rank,
                                                    eng.step, b1, b2, lr_d, lr_k, eps,
                                                    self.sync_ptrs, self.epoch if fused else 0,
                                                    ptr(eng.loss_buf) if fused else None,
                                                    ptr(self.loss_global) if fused else None, s),
                      "voxel_adam_exchange_mc")
            elif self.mode == "hybrid":
                check(lib.pp_voxel_adam_exchange_hybrid(self.vox_ptrs, self.grad_ptrs[self.cur],
                                                        self.vox_mem.multicast_ptr,
                                                        self.grad_mems[self.cur].multicast_ptr, ptr(eng.exp_avg),
                                                        ptr(eng.exp_avg_sq), eng.n_voxels, self.world, self.rank,
                                                        eng.step, b1, b2, lr_d, lr_k, eps, self.sync_ptrs,
                                                        self.epoch if fused else 0,
                                                        ptr(eng.loss_buf) if fused else None,
                                                        ptr(self.loss_global) if fused else None, self.mc_fraction, s),
                      "voxel_adam_exchange_hybrid")
            else:
                check(lib.pp_voxel_adam_exchange(self.vox_ptrs, self.grad_ptrs[self.cur], ptr(eng.exp_avg),
                                                 ptr(eng.exp_avg_sq), eng.n_voxels, self.world, self.rank,
                                                 eng.step, b1, b2, lr_d, lr_k, eps, self.sync_ptrs[self.rank], s),
                      "voxel_adam_exchange")
            if prof is not None:
                prof[2].record()
            # barrier: all slices written, all peers done reading my gradient
            self.epoch += 1
            check(lib.pp_peer_sync(self.sync_ptrs, self.world, self.rank, self.epoch, None, None, s), "peer_sync")
            if prof is not None:
                prof[3].record()
        # every peer has finished reading this gradient buffer: clear it on the side stream while the next
        # step marches into the other buffer
        self.consumed.record()
        with torch.cuda.stream(self.side):
            self.side.wait_event(self.consumed)
            self.grad_mems[self.cur].tensor.zero_()
            self.zeroed[self.cur].record()
        self.cur ^= 1
        if prof is not None:
            prof[4].record()
        return self.loss_global

    def check_health(self):
        """Raises if a cross-rank barrier of this rank ever timed out (a peer died or stalled for ~60 s): the exchange
        kernels stop updating from that point on and the reported loss is NaN.  Synchronises the device."""
        if self.mode in ("p2p", "nvls", "hybrid"):
            with torch.cuda.device(self.eng.device):
                if lib.pp_peer_sync_timed_out(self.sync_ptrs[self.rank]) != 0:
                    raise RuntimeError("DistributedDVGO: a cross-rank barrier timed out on rank %d - a peer stalled "
                                       "or died; parameters were not updated since" % self.rank)

    def render(self, rays_o, rays_d, **kw):
        return self.eng.render(rays_o, rays_d, **kw)

    def export_grids(self):
        self.check_health()
        return self.eng.export_grids()

    def close(self):
        if self.symm:
            torch.cuda.synchronize(self.eng.device)
            dist.barrier(group=self.group)
        if self.mode == "p2p":
            torch.cuda.synchronize(self.eng.device)
            dist.barrier(group=self.group)
            self.vox_mem.close()
            for g in self.grad_mems:
                g.close()
            self.sync_mem.close()
            dist.barrier(group=self.group)      # nobody maps our buffers any more
            self.eng.vox = self.eng.grad = None
            for mem in [self.vox_mem, self.sync_mem] + self.grad_mems:
                mem.tensor = None
                with torch.cuda.device(mem.device):
                    lib.pp_shared_free(ctypes.c_void_p(mem.address))
