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
