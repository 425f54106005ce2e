"""rgbnet forward on tcgen05: parity against torch.  The kernel rounds inputs, weights and hidden
activations to bf16 and accumulates in fp32 (TMEM), so it is compared (a) tightly against a torch model that
applies the same roundings and (b) against the plain fp32 nn.Sequential of the reference
(lib/voxurf_coarse.py:208-216) with the bf16 tolerance stated below."""
import pytest
import torch
import torch.nn as nn

pytestmark = pytest.mark.gpu


def make_rgbnet(d0, depth, width=128, seed=0):
    torch.manual_seed(seed)
    net = nn.Sequential(
        nn.Linear(d0, width), nn.ReLU(inplace=True),
        *[nn.Sequential(nn.Linear(width, width), nn.ReLU(inplace=True)) for _ in range(depth - 2)],
        nn.Linear(width, 3))
    nn.init.constant_(net[-1].bias, 0)
    return net


def bf16_emulated(net, x):
    lin = [m for m in net.modules() if isinstance(m, nn.Linear)]
    h = x.to(torch.bfloat16).float()
    for i, m in enumerate(lin):
        w = m.weight.to(torch.bfloat16).float()
        h = h.double() @ w.double().T + m.bias.double()
        if i < len(lin) - 1:
            h = torch.relu(h).float().to(torch.bfloat16).float()
    return torch.sigmoid(h).float()


@pytest.mark.parametrize("d0,depth,n", [(57, 4, 1000), (72, 3, 128), (39, 2, 5), (128, 5, 4099), (57, 4, 300000)])
def test_rgbnet_forward(cuda, d0, depth, n):
    from poseprobe_b200 import ops
    net = make_rgbnet(d0, depth, seed=d0 + depth)
    with torch.no_grad():
        for m in net.modules():
            if isinstance(m, nn.Linear):
                m.bias.add_(0.05 * torch.randn_like(m.bias))
    g = torch.Generator().manual_seed(n)
    x = torch.randn(n, d0, generator=g)
    weights, biases = ops.rgbnet_from_sequential(net)
    got = ops.rgbnet_forward(x.cuda(), [w.cuda() for w in weights], [b.cuda() for b in biases]).cpu()
    with torch.no_grad():
        want_bf16 = bf16_emulated(net, x)
        want_fp32 = torch.sigmoid(net(x))
    assert got.shape == (n, 3)
    # same roundings: only the fp32 accumulation order differs
    torch.testing.assert_close(got, want_bf16, rtol=0, atol=2e-3)
    # bf16 operands (8 mantissa bits) through `depth` layers vs the fp32 reference network
    torch.testing.assert_close(got, want_fp32, rtol=0, atol=2e-2)
    assert float((got - want_fp32).abs().mean()) < 3e-3


def _rel(a, b):
    return float((a - b).norm() / (b.norm() + 1e-12))


def _ste_bf16(t):
    """round to bf16 in the forward, identity in the backward"""
    return t + (t.to(torch.bfloat16).float() - t).detach()


def bf16_emulated_autograd(net, x, gout):
    """The kernel's numerics model in torch: inputs, weights and hidden activations rounded to bf16, fp32
    accumulation; backward through the same (bf16-forward) ReLU masks.  Returns (gx, [gW], [gb])."""
    lin = [m for m in net.modules() if isinstance(m, nn.Linear)]
    xr = x.clone().requires_grad_(True)
    ws = [m.weight.detach().clone().requires_grad_(True) for m in lin]
    bs = [m.bias.detach().clone().requires_grad_(True) for m in lin]
    h = _ste_bf16(xr)
    for i in range(len(lin)):
        h = h @ _ste_bf16(ws[i]).T + bs[i]
        if i < len(lin) - 1:
            h = _ste_bf16(torch.relu(h))
    torch.sigmoid(h).backward(gout)
    return xr.grad, [w.grad for w in ws], [b.grad for b in bs]


@pytest.mark.parametrize("d0,depth,n", [(57, 4, 1000), (72, 3, 128), (39, 2, 5), (57, 4, 200000), (64, 3, 4099)])
def test_rgbnet_backward(cuda, d0, depth, n):
    """Gradients of sigmoid(rgbnet(x)) w.r.t. x, weights and biases from the fused tcgen05 backward.
    (a) against torch autograd on the kernel's numerics model (bf16 operands, fp32 accumulation): the kernel
        additionally rounds every dZ_l to bf16 (8 mantissa bits, 2^-9 relative per element) -> 1e-2 relative
        (Frobenius) per gradient tensor;
    (b) against the plain fp32 network of the reference: 1e-1 relative, the price of bf16 operands through
        `depth` layers (a hidden unit whose pre-activation rounds across zero flips its ReLU mask)."""
    from poseprobe_b200 import ops
    net = make_rgbnet(d0, depth, seed=3 * d0 + depth)
    with torch.no_grad():
        for m in net.modules():
            if isinstance(m, nn.Linear):
                m.bias.add_(0.05 * torch.randn_like(m.bias))
    g = torch.Generator().manual_seed(n + 1)
    x = torch.randn(n, d0, generator=g)
    gout = torch.randn(n, 3, generator=g) / n
    lin = [m for m in net.modules() if isinstance(m, nn.Linear)]
    xr = x.clone().requires_grad_(True)
    torch.sigmoid(net(xr)).backward(gout)
    gx_e, gw_e, gb_e = bf16_emulated_autograd(net, x, gout)
    xc = x.cuda().requires_grad_(True)
    ws = [m.weight.detach().cuda().requires_grad_(True) for m in lin]
    bs = [m.bias.detach().cuda().requires_grad_(True) for m in lin]
    out = ops.rgbnet(xc, ws, bs)
    out.backward(gout.cuda())
    torch.cuda.synchronize()
    errs = {"gx": (_rel(xc.grad.cpu(), gx_e), _rel(xc.grad.cpu(), xr.grad))}
    for i, m in enumerate(lin):
        errs["gW%d" % i] = (_rel(ws[i].grad.cpu(), gw_e[i]), _rel(ws[i].grad.cpu(), m.weight.grad))
        errs["gb%d" % i] = (_rel(bs[i].grad.cpu(), gb_e[i]), _rel(bs[i].grad.cpu(), m.bias.grad))
    print(errs)
    for k, (e_model, e_fp32) in errs.items():
        assert e_model < 1e-2, "%s vs the bf16 numerics model: %g (all: %s)" % (k, e_model, errs)
        assert e_fp32 < 1e-1, "%s vs the fp32 network: %g (all: %s)" % (k, e_fp32, errs)
    # gradients are accumulated into zero-initialised buffers by the shim: a fresh call gives the same result
    ws2 = [w.detach().clone().requires_grad_(True) for w in ws]
    bs2 = [b.detach().clone().requires_grad_(True) for b in bs]
    ops.rgbnet(x.cuda(), ws2, bs2).backward(gout.cuda())
    for a, b in zip(ws, ws2):
        torch.testing.assert_close(a.grad, b.grad, rtol=1e-3, atol=1e-6)
    from poseprobe_b200._lib import lib
    assert lib.pp_rgbnet_backward_stalled() == 0      # no barrier wait ran out of its cycle budget


def test_rgbnet_forward_composite_epilogue(cuda):
    """rgb_marched[ray] += w * rgb fused into the forward kernel's epilogue == segment sum of the plain output."""
    import ctypes
    from poseprobe_b200 import ops
    from poseprobe_b200._lib import check, lib, ptr, stream
    d0, depth, n, n_rays = 57, 4, 70001, 300
    net = make_rgbnet(d0, depth, seed=11)
    weights, biases = ops.rgbnet_from_sequential(net)
    ws = [w.detach().cuda().contiguous() for w in weights]
    bs = [b.detach().cuda().contiguous() for b in biases]
    g = torch.Generator().manual_seed(5)
    x = torch.randn(n, d0, generator=g).cuda()
    w = torch.rand(n, generator=g).cuda()
    ray_id = torch.sort(torch.randint(0, n_rays, (n,), generator=g))[0].to(torch.int32).cuda()
    k0p = int(lib.pp_rgbnet_padded_width(d0))
    xp = torch.empty(n, k0p, dtype=torch.bfloat16, device="cuda")
    check(lib.pp_rgbnet_pack_input(ptr(x), n, d0, ptr(xp), stream()), "pack")
    out = torch.empty(n, 3, device="cuda")
    marched = torch.zeros(n_rays, 3, device="cuda")
    arr = lambda ts: (ctypes.c_void_p * depth)(*[t.data_ptr() for t in ts])
    check(lib.pp_rgbnet_forward_packed(ptr(xp), n, d0, depth, 128, arr(ws), arr(bs), 1, ptr(out), ptr(w), ptr(ray_id),
                                       ptr(marched), None, stream()), "fwd")
    want = torch.zeros(n_rays, 3, device="cuda").index_add_(0, ray_id.long(), w[:, None] * out)
    torch.testing.assert_close(marched, want, rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(out, ops.rgbnet_forward(x, ws, bs), rtol=0, atol=0)
