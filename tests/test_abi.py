"""CPU tier: the C-ABI library loads, exports every symbol include/*.h declares, and the Python shims
mirror the reference's four extension modules (names, arity, error behaviour).  No kernel runs here."""
import ctypes
import inspect
import os
import re
import subprocess

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# the pybind surface of the reference: lib/cuda/render_utils.cpp:170-184, total_variation.cpp:29-32,
# adam_upd.cpp:79-86, ub360_utils.cpp:21-23  (name -> positional arg count)
REFERENCE_SURFACE = {
    "render_utils_cuda": {
        "infer_t_minmax": 6, "infer_n_samples": 4, "infer_ray_start_dir": 3, "sample_pts_on_rays": 7,
        "sample_ndc_pts_on_rays": 5, "sample_bg_pts_on_rays": 5, "maskcache_lookup": 4, "raw2alpha": 3,
        "raw2alpha_backward": 3, "raw2alpha_nonuni": 3, "raw2alpha_nonuni_backward": 3, "alpha2weight": 3,
        "alpha2weight_backward": 9},
    "total_variation_cuda": {"total_variation_add_grad": 6, "total_variation_add_grad_new": 7},
    "adam_upd_cuda": {"adam_upd": 9, "masked_adam_upd": 9, "adam_upd_with_perlr": 10},
    "ub360_utils_cuda": {"cumdist_thres": 2},
}


def test_library_exports_every_declared_symbol():
    from poseprobe_b200 import _lib
    protos = _lib.parse_header()
    assert len(protos) >= 28
    out = subprocess.check_output(["nm", "-D", "--defined-only", _lib.LIB_PATH], text=True)
    exported = set(re.findall(r"\bT (pp_\w+)", out))
    missing = set(protos) - exported
    assert not missing, "declared but not exported: %s" % sorted(missing)
    undeclared = exported - set(protos)
    assert not undeclared, "exported but not declared in include/: %s" % sorted(undeclared)
    lib = ctypes.CDLL(_lib.LIB_PATH)
    assert lib.pp_abi_version() == 1
    lib.pp_build_info.restype = ctypes.c_char_p
    assert b"sm_100a" in lib.pp_build_info()


def test_library_contains_sm100a_code():
    from poseprobe_b200 import _lib
    out = subprocess.check_output(["cuobjdump", "-lelf", _lib.LIB_PATH], text=True)
    assert "sm_100a" in out


def test_shim_surface_matches_reference():
    import poseprobe_b200 as pp
    total = 0
    for mod_name, fns in REFERENCE_SURFACE.items():
        mod = getattr(pp, mod_name)
        for fn_name, n_args in fns.items():
            fn = getattr(mod, fn_name)
            sig = inspect.signature(fn)
            assert len(sig.parameters) == n_args, (mod_name, fn_name, sig)
            total += 1
    assert total == 19


@pytest.mark.skipif(not os.path.isdir("/root/reference/lib/cuda"), reason="reference checkout not present")
def test_surface_table_agrees_with_reference_sources():
    """The table above is checked against the reference's pybind definitions where it is available."""
    for mod_name, fns in REFERENCE_SURFACE.items():
        stem = mod_name.replace("_cuda", "")
        src = open("/root/reference/lib/cuda/%s.cpp" % stem).read()
        defs = re.findall(r'm\.def\("(\w+)"', src)
        assert sorted(defs) == sorted(fns), (mod_name, defs)


def test_no_cpu_fallback():
    import poseprobe_b200 as pp
    ru = pp.render_utils_cuda
    with pytest.raises(RuntimeError, match="alpha must be a CUDA tensor"):
        ru.alpha2weight(torch.rand(4), torch.zeros(4, dtype=torch.int64), 1)
    with pytest.raises(RuntimeError, match="rays_o must be a CUDA tensor"):
        ru.sample_pts_on_rays(torch.rand(4, 3), torch.rand(4, 3), torch.zeros(3), torch.ones(3), 0.1, 1.0, 0.1)
    with pytest.raises(RuntimeError, match="param must be a CUDA tensor"):
        pp.total_variation_cuda.total_variation_add_grad(torch.rand(1, 1, 2, 2, 2), torch.rand(1, 1, 2, 2, 2), 1, 1, 1, True)
    with pytest.raises(RuntimeError, match="param must be a CUDA tensor"):
        pp.adam_upd_cuda.adam_upd(torch.rand(4), torch.rand(4), torch.rand(4), torch.rand(4), 1, 0.9, 0.99, 0.1, 1e-8)
    with pytest.raises(RuntimeError, match="dist must be a CUDA tensor"):
        pp.ub360_utils_cuda.cumdist_thres(torch.rand(2, 2), 0.1)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        pp.ops.segment_sum(torch.rand(4, 3), torch.zeros(4, dtype=torch.int64), 1)
    dims = [3, 128, 128, 128, 128, 4]
    ws = [torch.rand(dims[i + 1], dims[i]) for i in range(5)]
    bs = [torch.rand(dims[i + 1]) for i in range(5)]
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        pp.ops.warpnet_forward(torch.rand(8, 3), ws, bs)
    with pytest.raises(RuntimeError, match="expected layers 3 -> 128"):
        pp.ops.warpnet_params(ws[1:], bs[1:])
    from poseprobe_b200.deform import DeformedImplicitField
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        DeformedImplicitField(range_shape=torch.ones(3))(torch.rand(8, 3))


def test_product_does_not_import_the_oracle():
    """The product path must never route through oracle/ (or any CPU restatement)."""
    pkg = os.path.join(ROOT, "poseprobe_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f
                assert "libpp_oracle" not in text, f


def test_c_abi_argument_validation():
    """Status codes and error text without launching anything (null pointers / bad extents)."""
    from poseprobe_b200 import _lib
    lib = _lib.lib
    assert lib.pp_raw2alpha(0, None, 0.0, 0.5, None, 16, None, None, None) == -1
    assert b"null pointer" in lib.pp_last_error()
    assert lib.pp_raw2alpha(7, None, 0.0, 0.5, None, 0, None, None, None) == 0          # n == 0 is a no-op
    assert lib.pp_alpha2weight(0, None, None, -1, 4, None, None, None, None, None, None) == -1
    assert lib.pp_adam_upd(0, 9, None, None, None, None, None, 4, 1, 0.9, 0.99, 0.1, 1e-8, None) == -1
    assert lib.pp_sample_pts_fill(0, None, None, None, None, None, None, 0.1, 4, 2**31, None, None, None, None, None) == -3
    assert lib.pp_check_device_ptr(None) == -1
    # deformation path: extents and null pointers are rejected before anything is launched
    assert lib.pp_warpnet_forward_packed(None, 6, 5, 128, None, None, None, None, None) == -1      # rows % 4 != 0
    assert b"multiple of 4" in lib.pp_last_error()
    assert lib.pp_warpnet_forward_packed(None, 8, 5, 64, None, None, None, None, None) == -5       # width != 128
    assert lib.pp_warpnet_forward_packed(None, 8, 6, 128, None, None, None, None, None) == -5      # depth > 5
    assert lib.pp_warpnet_forward_packed(None, 0, 5, 128, None, None, None, None, None) == 0       # empty: no-op
    assert lib.pp_warpnet_forward_packed(None, 8, 5, 128, None, None, None, None, None) == -1
    assert lib.pp_warpnet_pack_points(None, 4, None, None) == -1
    assert lib.pp_voxurf_deform_rows(None, None, None, None, 4, 4, None, None, None, None, None, None) == -1
    assert lib.pp_voxurf_deform_sample_backward(None, None, None, None, None, None, 4, 1.0, 0, 0, 0, 0, None, None,
                                                None, None, None) == -1
    assert lib.pp_voxurf_deform_sample_forward(None, None, None, None, None, 4, 1.0, None, None, None) == -1
    assert lib.pp_voxurf_deform_ray_grad(None, None, None, None, 4, None, None, None, None) == -1
    assert lib.pp_voxurf_sample_records(None, None, None, None, 4, None, None) == -1
    assert lib.pp_voxurf_sample_records_backward(None, None, None, None, 4, None, 0.0, None, None, None) == -1
    assert lib.pp_voxurf_object_loss(4, None, None, None, None, None, 0.0, 1.0, 0.0, 0.0, None, None, None, None, None,
                                     None, None) == -1
    assert lib.pp_voxurf_k0_tv(None, None, None, 1.0, None, None) == -1
    # deformation network backward: depth 5 / width 128 only, rows in fours, null pointers rejected
    assert lib.pp_warpnet_backward_packed(None, None, 6, None, 5, 128, None, None, None, None, None, None) == -1
    assert lib.pp_warpnet_backward_packed(None, None, 8, None, 4, 128, None, None, None, None, None, None) == -5
    assert lib.pp_warpnet_backward_packed(None, None, 8, None, 5, 64, None, None, None, None, None, None) == -5
    assert lib.pp_warpnet_backward_packed(None, None, 0, None, 5, 128, None, None, None, None, None, None) == 0
    assert lib.pp_warpnet_backward_packed(None, None, 8, None, 5, 128, None, None, None, None, None, None) == -1
    assert lib.pp_warpnet_backward_workspace_bytes() >= 4 and lib.pp_warpnet_backward_stalled() == 0
    assert lib.pp_peer_sync_timed_out(None) == 0
    assert lib.pp_adam_upd(0, 4, None, None, None, None, None, 4, 1, 0.9, 0.99, 0.1, 1e-8, None) == -1   # modes 0..3


def test_host_fed_entry_points_validate_before_touching_the_device():
    """pp_dvgo_render_host / pp_dvgo_step_graph_*: extents, null pointers and the scene are checked before any CUDA
    call; empty work is a no-op; a null graph is rejected, destroying one is harmless."""
    import ctypes
    from poseprobe_b200 import _lib, dvgo
    lib = _lib.lib
    sc = dvgo.make_scene((8, 8, 8), [-1.0, -1, -1], [1.0, 1, 1], 0.2, 1e9, 0.125, -4.6, 0.5, 0.0, 0.0)
    bad = dvgo.make_scene((8, 8, 8), [-1.0, -1, -1], [1.0, -1, 1], 0.2, 1e9, 0.125, -4.6, 0.5, 0.0, 0.0)
    ref = ctypes.byref
    assert lib.pp_dvgo_render_host(ref(sc), None, None, 0, 8192, None, None, None, None) == 0          # no rays: no-op
    assert lib.pp_dvgo_render_host(ref(sc), None, None, -1, 8192, None, None, None, None) == -1
    assert lib.pp_dvgo_render_host(ref(sc), None, None, 16, 0, None, None, None, None) == -1
    assert b"chunk" in lib.pp_last_error()
    assert lib.pp_dvgo_render_host(ref(sc), None, None, 16, 8, None, None, None, None) == -1
    assert b"null pointer" in lib.pp_last_error()
    assert lib.pp_dvgo_render_host(ref(bad), None, None, 16, 8, None, None, None, None) == -1
    assert b"bounding box" in lib.pp_last_error()
    assert lib.pp_dvgo_render_host(None, None, None, 16, 8, None, None, None, None) == -1
    out = ctypes.c_void_p()
    args = (None, None, None, None, None, None)
    assert lib.pp_dvgo_step_graph_create(ref(sc), None, *args, 16, None, None, None, 0, 0.9, 0.99, 0.1, 0.1, 1e-8, 1,
                                         None) == -1
    assert b"graph_out" in lib.pp_last_error()
    assert lib.pp_dvgo_step_graph_create(ref(sc), None, *args, 0, None, None, None, 0, 0.9, 0.99, 0.1, 0.1, 1e-8, 1,
                                         ref(out)) == -1
    assert b"n_rays" in lib.pp_last_error()
    assert lib.pp_dvgo_step_graph_create(ref(sc), None, *args, 16, None, None, None, 0, 0.9, 0.99, 0.1, 0.1, 1e-8, 1,
                                         ref(out)) == -1
    assert b"null pointer" in lib.pp_last_error() and out.value is None
    assert lib.pp_dvgo_step_graph_launch(None, None, 1, None) == -1
    assert lib.pp_dvgo_step_graph_destroy(None) == 0


def test_maskcache_loads_a_reference_style_checkpoint(tmp_path):
    """dvgo.MaskCache(path=...) reads checkpoints as the reference writes them (numpy arrays in MaskCache_kwargs,
    lib/dvgo_ori.py:143-152, :436-451); construction is host logic only."""
    import numpy as np
    import torch
    from poseprobe_b200 import dvgo
    ck = {"MaskCache_kwargs": {"xyz_min": np.array([-1., -1, -1], dtype=np.float32),
                               "xyz_max": np.array([1., 1, 1], dtype=np.float32), "act_shift": -4.6,
                               "voxel_size_ratio": 1.0, "nearest": False, "pre_act_density": False,
                               "in_act_density": False},
          "model_state_dict": {"density": torch.randn(1, 1, 8, 8, 8)}}
    path = str(tmp_path / "coarse_last.tar")
    torch.save(ck, path)
    m = dvgo.MaskCache(path=path, mask_cache_thres=1e-3)
    assert tuple(m.density.shape) == (1, 1, 8, 8, 8) and m.xyz_max.tolist() == [1.0, 1.0, 1.0]
    pooled = torch.nn.functional.max_pool3d(ck["model_state_dict"]["density"], kernel_size=3, padding=1, stride=1)
    assert torch.equal(m.density, pooled)
