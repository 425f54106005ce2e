"""ctypes binding of libposeprobe_b200.so (the C-ABI CUDA library).

`include/poseprobe_b200.h` is the single source of truth: the prototypes are parsed from it at
import time and turned into ctypes signatures, so a function that is declared but not exported (or
the other way round) fails loudly here instead of silently mis-calling.

There is no fallback of any kind: if the shared library is missing or a symbol cannot be resolved
the import raises.  Build the library with `python poseprobe_b200/build.py` (or
`__graft_entry__.build()`).
"""
import ctypes
import os
import re

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libposeprobe_b200.so")
HEADER_PATHS = [os.path.join(os.path.dirname(_HERE), "include", "poseprobe_b200.h"),
                os.path.join(os.path.dirname(_HERE), "include", "poseprobe_b200_fused.h")]

PP_F32, PP_F64 = 0, 1

_CTYPES = {
    "float": ctypes.c_float, "double": ctypes.c_double, "int": ctypes.c_int,
    "int32_t": ctypes.c_int32, "int64_t": ctypes.c_int64, "uint64_t": ctypes.c_uint64, "uint32_t": ctypes.c_uint32,
    "long": ctypes.c_longlong, "pp_dtype": ctypes.c_int, "pp_stream_t": ctypes.c_void_p, "size_t": ctypes.c_size_t,
}
_PROTO = re.compile(r"PP_API\s+([\w\s\*]+?)\s*\b(pp_\w+)\s*\(([^;]*?)\)\s*;", re.S)


def _ctype_of(decl):
    decl = decl.strip()
    if decl == "void":
        return None
    if "*" in decl:
        return ctypes.c_char_p if decl.replace(" ", "") == "constchar*" else ctypes.c_void_p
    base = decl.replace("const", "").split()
    # drop the parameter name if present
    tname = base[0]
    if tname not in _CTYPES:
        raise RuntimeError("poseprobe_b200: unknown C type %r in header" % decl)
    return _CTYPES[tname]


def parse_header(paths=None):
    """-> {name: (restype, [argtypes])} for every PP_API prototype."""
    protos = {}
    for path in (paths or HEADER_PATHS):
        if not os.path.exists(path):
            continue
        text = re.sub(r"/\*.*?\*/", "", open(path).read(), flags=re.S)
        for ret, name, args in _PROTO.findall(text):
            args = args.strip()
            argtypes = [] if args in ("", "void") else [_ctype_of(a) for a in args.split(",")]
            protos[name] = (_ctype_of(ret), argtypes)
    return protos


class PoseProbeError(RuntimeError):
    pass


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "poseprobe_b200: %s is missing - the CUDA library has not been built "
            "(run `python poseprobe_b200/build.py`). There is no CPU fallback." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    protos = parse_header()
    if not protos:
        raise ImportError("poseprobe_b200: no prototypes found in %s" % HEADER_PATHS[0])
    for name, (restype, argtypes) in protos.items():
        try:
            fn = getattr(lib, name)
        except AttributeError:
            raise ImportError("poseprobe_b200: %s is declared in the header but not exported by %s"
                              % (name, LIB_PATH))
        fn.restype = restype
        fn.argtypes = argtypes
    return lib, protos


lib, PROTOS = _load()
if lib.pp_abi_version() != 1:
    raise ImportError("poseprobe_b200: ABI version mismatch")


def check(rc, what=""):
    if rc != 0:
        msg = lib.pp_last_error()
        raise PoseProbeError("poseprobe_b200 %s failed (status %d): %s"
                             % (what, rc, msg.decode() if msg else "?"))


def stream():
    """torch's current CUDA stream as a raw cudaStream_t."""
    return torch.cuda.current_stream().cuda_stream


def ptr(t):
    return None if t is None else t.data_ptr()


_ATEN_NAMES = {torch.float16: "Half", torch.bfloat16: "BFloat16", torch.int64: "Long", torch.int32: "Int",
               torch.int16: "Short", torch.int8: "Char", torch.uint8: "Byte", torch.bool: "Bool"}


def dtype_code(t, name):
    if t.dtype == torch.float32:
        return PP_F32
    if t.dtype == torch.float64:
        return PP_F64
    # mirrors AT_DISPATCH_FLOATING_TYPES' error text
    raise RuntimeError('"%s" not implemented for \'%s\'' % (name, _ATEN_NAMES.get(t.dtype, str(t.dtype))))


def check_input(t, name):
    """CHECK_INPUT of the reference (render_utils.cpp:46-48): same error text, no CPU fallback."""
    if not t.is_cuda:
        raise RuntimeError("%s must be a CUDA tensor" % name)
    if not t.is_contiguous():
        raise RuntimeError("%s must be contiguous" % name)


def launch_count():
    return int(lib.pp_launch_count())
