"""TEST INFRASTRUCTURE ONLY -- builds the reference's own CUDA extensions into oracle/_ref/.

The four reference extension modules (lib/cuda/{render_utils,total_variation,adam_upd,ub360_utils}
[.cpp + _kernel.cu]) are compiled *from where they lie* under /root/reference -- nothing is copied
into this repo.  The only accommodation is the force-included oracle/ref_compat.h (one overload
that modern torch dropped).  Outputs (ninja files, objects, .so) land in oracle/_ref/, which is
git-ignored but travels to the GPU box with gpurun.

The resulting modules are named ref_<name> so they never shadow the product's modules.
They are used by tests/ (parity: product kernels vs the reference kernels on the same B200) and
by tests/golden/make_golden.py (fixture generation).  The product never imports them.

Usage:  python oracle/build_ref.py [module ...]
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_CUDA = os.environ.get("POSEPROBE_REFERENCE", "/root/reference") + "/lib/cuda"
OUT = os.path.join(HERE, "_ref")
MODULES = {
    "render_utils_cuda": ["render_utils.cpp", "render_utils_kernel.cu"],
    "total_variation_cuda": ["total_variation.cpp", "total_variation_kernel.cu"],
    "adam_upd_cuda": ["adam_upd.cpp", "adam_upd_kernel.cu"],
    "ub360_utils_cuda": ["ub360_utils.cpp", "ub360_utils_kernel.cu"],
}


def ref_available() -> bool:
    return os.path.isdir(REF_CUDA)


def so_path(name: str) -> str:
    return os.path.join(OUT, "ref_" + name, "ref_" + name + ".so")


def build(names=None, verbose=False):
    if not ref_available():
        raise RuntimeError("reference sources not present at %s" % REF_CUDA)
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    os.environ.setdefault("MAX_JOBS", "4")
    from torch.utils.cpp_extension import load
    compat = os.path.join(HERE, "ref_compat.h")
    for name in (names or MODULES):
        if os.path.exists(so_path(name)):
            continue
        bdir = os.path.join(OUT, "ref_" + name)
        os.makedirs(bdir, exist_ok=True)
        load(name="ref_" + name,
             sources=[os.path.join(REF_CUDA, s) for s in MODULES[name]],
             extra_cflags=["-include", compat, "-w"],
             extra_cuda_cflags=["-include", compat, "-w"],   # reference passes no flags: default -fmad=true
             build_directory=bdir, verbose=verbose, is_python_module=False)


def load_ref(name: str):
    """Import oracle/_ref/ref_<name>.so (prebuilt).  Returns None when it is absent."""
    p = so_path(name)
    if not os.path.exists(p):
        return None
    import importlib.util
    import torch  # noqa: F401  (libtorch must be loaded first)
    spec = importlib.util.spec_from_file_location("ref_" + name, p)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    build(sys.argv[1:] or None, verbose=True)
    print("built:", [so_path(n) for n in (sys.argv[1:] or MODULES)])
