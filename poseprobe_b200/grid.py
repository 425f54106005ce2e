"""Host-side mirror of the reference's ``lib/grid.py`` for the dense-grid path.

``DenseGrid`` and ``MaskGrid`` keep the reference's constructor arguments, parameter / buffer names
(``grid`` as ``[1,C,X,Y,Z]`` fp32 so ``state_dict`` keys such as ``k0.grid`` / ``sdf.grid`` stay
interchangeable, lib/recon_scene.py:779-791) and method names; the trilinear gather/scatter and the
TV / mask lookups run on the sm_100a kernels instead of ATen's grid_sampler and the JIT extension.
``TensoRFGrid`` is out of scope (never constructed by the reference, voxurf_coarse.py:121-123).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops, render_utils_cuda, total_variation_cuda


def create_grid(type, **kwargs):
    if type == 'DenseGrid':
        return DenseGrid(**kwargs)
    raise NotImplementedError(type)


class DenseGrid(nn.Module):
    """lib/grid.py:38-89"""

    def __init__(self, channels, world_size, xyz_min, xyz_max, **kwargs):
        super().__init__()
        self.channels = channels
        self.world_size = world_size
        self.register_buffer('xyz_min', torch.Tensor(xyz_min))
        self.register_buffer('xyz_max', torch.Tensor(xyz_max))
        self.grid = nn.Parameter(torch.zeros([1, channels, *world_size]))

    def forward(self, xyz):
        """xyz: [...,3] global coordinates -> [...,C] (squeezed when C == 1), zero padding."""
        shape = xyz.shape[:-1]
        out = ops.trilinear(self.grid, xyz, self.xyz_min, self.xyz_max, padding_mode="zeros")
        out = out.reshape(*shape, self.channels)
        if self.channels == 1:
            out = out.squeeze(-1)
        return out

    def scale_volume_grid(self, new_world_size):
        if self.channels == 0:
            self.grid = nn.Parameter(torch.zeros([1, self.channels, *new_world_size]))
        else:
            self.grid = nn.Parameter(
                F.interpolate(self.grid.data, size=tuple(new_world_size), mode='trilinear', align_corners=True))

    def total_variation_add_grad(self, wx, wy, wz, dense_mode, mask=None):
        """Add gradients by total variation loss in-place (lib/grid.py:67-78)."""
        if mask is None:
            total_variation_cuda.total_variation_add_grad(self.grid, self.grid.grad, wx, wy, wz, dense_mode)
        else:
            mask = mask.detach()
            if self.grid.size(1) > 1 and mask.size() != self.grid.size():
                mask = mask.repeat(1, self.grid.size(1), 1, 1, 1).contiguous()
            assert mask.size() == self.grid.size()
            total_variation_cuda.total_variation_add_grad_new(
                self.grid, self.grid.grad, mask.float(), wx, wy, wz, dense_mode)

    def get_dense_grid(self):
        return self.grid

    @torch.no_grad()
    def __isub__(self, val):
        self.grid.data -= val
        return self

    def extra_repr(self):
        return f'channels={self.channels}, world_size={list(self.world_size)}'


class MaskGrid(nn.Module):
    """lib/grid.py:212-245"""

    def __init__(self, path=None, mask_cache_thres=None, mask=None, xyz_min=None, xyz_max=None):
        super().__init__()
        if path is not None:
            st = torch.load(path)
            self.mask_cache_thres = mask_cache_thres
            density = F.max_pool3d(st['model_state_dict']['density.grid'], kernel_size=3, padding=1, stride=1)
            alpha = 1 - torch.exp(-F.softplus(density + st['model_state_dict']['act_shift']) * st['model_kwargs']['voxel_size_ratio'])
            mask = (alpha >= self.mask_cache_thres).squeeze(0).squeeze(0)
            xyz_min = torch.Tensor(st['model_kwargs']['xyz_min'])
            xyz_max = torch.Tensor(st['model_kwargs']['xyz_max'])
        else:
            mask = mask.bool()
            xyz_min = torch.Tensor(xyz_min)
            xyz_max = torch.Tensor(xyz_max)
        self.register_buffer('mask', mask)
        xyz_len = xyz_max - xyz_min
        self.register_buffer('xyz2ijk_scale', (torch.Tensor(list(mask.shape)) - 1) / xyz_len)
        self.register_buffer('xyz2ijk_shift', -xyz_min * self.xyz2ijk_scale)

    @torch.no_grad()
    def forward(self, xyz):
        """Skip known free space. xyz: [...,3] -> bool [...]"""
        shape = xyz.shape[:-1]
        xyz = xyz.reshape(-1, 3)
        mask = render_utils_cuda.maskcache_lookup(self.mask, xyz.contiguous(), self.xyz2ijk_scale,
                                                  self.xyz2ijk_shift)
        return mask.reshape(shape)

    def extra_repr(self):
        return 'mask.shape=list(self.mask.shape)'
