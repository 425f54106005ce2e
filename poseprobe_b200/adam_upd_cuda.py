"""Drop-in for the reference's ``adam_upd_cuda`` extension (lib/cuda/adam_upd.cpp:79-86).

All three functions update ``param``, ``exp_avg`` and ``exp_avg_sq`` in place.  Note the kernel's
``eps`` placement (added to the un-bias-corrected sqrt(v); adam_upd_kernel.cu:20-22,72) differs from
the Python optimiser in lib/utils.py:176-198 -- this module follows the kernel.
"""
import torch

from ._lib import check, check_input, dtype_code, lib, ptr, stream


def _run(mode, name, param, grad, exp_avg, exp_avg_sq, perlr, step, beta1, beta2, lr, eps):
    for t, n in ((param, "param"), (grad, "grad"), (exp_avg, "exp_avg"), (exp_avg_sq, "exp_avg_sq")):
        check_input(t, n)
    dt = dtype_code(param, name)
    with torch.cuda.device(param.device):
        check(lib.pp_adam_upd(dt, mode, ptr(param), ptr(grad), ptr(exp_avg), ptr(exp_avg_sq), ptr(perlr),
                              param.numel(), step, beta1, beta2, lr, eps, stream()), name)


def adam_upd(param, grad, exp_avg, exp_avg_sq, step, beta1, beta2, lr, eps):
    """adam_upd.cpp:36-48"""
    _run(0, "adam_upd_cuda", param, grad, exp_avg, exp_avg_sq, None, step, beta1, beta2, lr, eps)


def masked_adam_upd(param, grad, exp_avg, exp_avg_sq, step, beta1, beta2, lr, eps):
    """adam_upd.cpp:50-62 -- elements with grad == 0 are left untouched"""
    _run(1, "masked_adam_upd_cuda", param, grad, exp_avg, exp_avg_sq, None, step, beta1, beta2, lr, eps)


def adam_upd_with_perlr(param, grad, exp_avg, exp_avg_sq, perlr, step, beta1, beta2, lr, eps):
    """adam_upd.cpp:64-77 -- per-element learning-rate multiplier"""
    _run(2, "adam_upd_with_perlr_cuda", param, grad, exp_avg, exp_avg_sq, perlr, step, beta1, beta2, lr, eps)
