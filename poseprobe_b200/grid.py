"""Host-side mirror of the reference's ``lib/grid.py`` for the dense-grid path.

``DenseGrid`` and ``MaskGrid`` answer to the reference's constructor arguments, parameter / buffer
names (``grid`` as ``[1,C,X,Y,Z]`` fp32 so ``state_dict`` keys such as ``k0.grid`` / ``sdf.grid`` stay
interchangeable, lib/recon_scene.py:779-791) and method names; the trilinear gather/scatter and the
TV / mask lookups run on the sm_100a kernels instead of ATen's grid_sampler and the JIT extension.
``TensoRFGrid`` is out of scope (never constructed by the reference, voxurf_coarse.py:121-123).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops, render_utils_cuda, total_variation_cuda


def _vec3(v):
    return torch.as_tensor(v, dtype=torch.float32).detach().clone().reshape(3)


def create_grid(type, **kwargs):
    """lib/grid.py:20-30 -- only the dense representation exists on this path."""
    if type != 'DenseGrid':
        raise NotImplementedError(type)
    return DenseGrid(**kwargs)


class DenseGrid(nn.Module):
    """A [1,C,X,Y,Z] fp32 parameter spanning the box [xyz_min, xyz_max] (lib/grid.py:38-89)."""

    def __init__(self, channels, world_size, xyz_min, xyz_max, **kwargs):
        super().__init__()
        self.channels = int(channels)
        self.world_size = world_size
        for name, value in (('xyz_min', xyz_min), ('xyz_max', xyz_max)):
            self.register_buffer(name, _vec3(value))
        self.grid = nn.Parameter(self._blank(world_size))

    def _blank(self, world_size):
        return torch.zeros(1, self.channels, *[int(n) for n in world_size])

    def forward(self, xyz):
        """xyz [...,3] in world coordinates -> [...,C] ([...] for C == 1); samples outside the box read 0
        (zero padding, align_corners=True -- the F.grid_sample call of lib/grid.py:47-58)."""
        lead = xyz.shape[:-1]
        vals = ops.trilinear(self.grid, xyz, self.xyz_min, self.xyz_max, padding_mode="zeros")
        return vals.reshape(*lead) if self.channels == 1 else vals.reshape(*lead, self.channels)

    def scale_volume_grid(self, new_world_size):
        """Resample to a new resolution, corner-aligned trilinear (lib/grid.py:60-65)."""
        size = tuple(int(n) for n in new_world_size)
        if self.channels > 0:
            resized = F.interpolate(self.grid.detach(), size=size, mode='trilinear', align_corners=True)
        else:
            resized = self._blank(size).to(self.grid.device)
        self.grid = nn.Parameter(resized)

    def total_variation_add_grad(self, wx, wy, wz, dense_mode, mask=None):
        """Accumulate the TV gradient into ``grid.grad`` in place (lib/grid.py:67-78); with ``mask`` the
        masked variant (total_variation_add_grad_new) runs, the mask broadcast over channels."""
        if mask is None:
            total_variation_cuda.total_variation_add_grad(self.grid, self.grid.grad, wx, wy, wz, dense_mode)
            return
        m = mask.detach()
        if m.shape != self.grid.shape and self.channels > 1:
            m = m.expand(-1, self.channels, -1, -1, -1) if m.shape[1] == 1 else m.repeat(1, self.channels, 1, 1, 1)
        if m.shape != self.grid.shape:
            raise AssertionError(f"mask {tuple(m.shape)} does not match grid {tuple(self.grid.shape)}")
        total_variation_cuda.total_variation_add_grad_new(
            self.grid, self.grid.grad, m.float().contiguous(), wx, wy, wz, dense_mode)

    def get_dense_grid(self):
        return self.grid

    @torch.no_grad()
    def __isub__(self, val):
        self.grid.sub_(val)
        return self

    def extra_repr(self):
        return f'channels={self.channels}, world_size={list(self.world_size)}'


def _occupancy_from_checkpoint(path, thres):
    """Coarse-stage checkpoint -> (bool [X,Y,Z], xyz_min, xyz_max): density dilated by a 3^3 max-pool, turned
    into alpha with the checkpoint's act_shift / voxel_size_ratio, thresholded (lib/grid.py:215-222)."""
    # (the reference's checkpoints hold numpy arrays in model_kwargs: load them the way the reference does, :215)
    ckpt = torch.load(path, map_location="cpu", weights_only=False)
    state, kwargs = ckpt['model_state_dict'], ckpt['model_kwargs']
    dilated = F.max_pool3d(state['density.grid'], kernel_size=3, stride=1, padding=1)
    alpha = 1 - torch.exp(-F.softplus(dilated + state['act_shift']) * kwargs['voxel_size_ratio'])
    return (alpha >= thres)[0, 0], kwargs['xyz_min'], kwargs['xyz_max']


class MaskGrid(nn.Module):
    """Bool occupancy volume with the world -> index affine map as buffers (lib/grid.py:212-245)."""

    def __init__(self, path=None, mask_cache_thres=None, mask=None, xyz_min=None, xyz_max=None):
        super().__init__()
        if path is None:
            occupied = mask.bool()
        else:
            self.mask_cache_thres = mask_cache_thres
            occupied, xyz_min, xyz_max = _occupancy_from_checkpoint(path, mask_cache_thres)
        lo, hi = _vec3(xyz_min).to(occupied.device), _vec3(xyz_max).to(occupied.device)
        cells = torch.tensor([n - 1 for n in occupied.shape], dtype=torch.float32, device=occupied.device)
        scale = cells / (hi - lo)
        self.register_buffer('mask', occupied)
        self.register_buffer('xyz2ijk_scale', scale)
        self.register_buffer('xyz2ijk_shift', -lo * scale)

    @torch.no_grad()
    def forward(self, xyz):
        """xyz [...,3] world coordinates -> bool [...]: True where the nearest voxel is occupied, False for
        free space and for points outside the volume (render_utils_kernel.cu:246-268)."""
        hit = render_utils_cuda.maskcache_lookup(self.mask, xyz.reshape(-1, 3).contiguous(),
                                                 self.xyz2ijk_scale, self.xyz2ijk_shift)
        return hit.reshape(xyz.shape[:-1])

    def extra_repr(self):
        return 'mask.shape=list(self.mask.shape)'
