"""Host-side mirror of the reference's deformation network (lib/deformation/deform_net.py, modules.py).

``DeformedImplicitField`` keeps the reference's constructor arguments and ``state_dict`` keys
(``progress``, ``deform_net.net.net.<i>.0.weight|bias`` - the nesting of SingleBVPNet -> FCBlock -> MetaSequential,
modules.py:43-124), so the ``warp_network.*`` entries of a reference checkpoint load unchanged.  Evaluation runs on the
tcgen05 kernel of mlp.cu (kJac): one pass yields the network output AND its input Jacobian, which the reference obtains
with four ``torch.autograd.grad`` calls (lib/voxurf_coarse.py:966-985).  Forward / inference only in this round: the
kernel does not build an autograd graph.
"""
import torch
import torch.nn as nn

from . import ops


class _Holder(nn.Module):
    """a module whose only child is ``net`` (reproduces one level of the reference's key nesting)"""

    def __init__(self, net):
        super().__init__()
        self.net = net


class DeformedImplicitField(nn.Module):
    """3 -> hidden_num x (1 + num_hidden_layers) -> 4, ReLU; outputs scaled by ``range_shape.max()``
    (deform_net.py:12-29).  Initialisation as the reference: Kaiming-normal weights (fan_in, ReLU), the last layer
    all zeros (modules.py:127-130, :166-170), so a fresh network deforms nothing."""

    def __init__(self, model_type='relu', hyper_hidden_layers=1, range_shape=None, hyper_hidden_features=256,
                 hidden_num=128, num_hidden_layers=3, **kwargs):
        super().__init__()
        if model_type != 'relu' or hidden_num != 128:
            raise NotImplementedError("DeformedImplicitField: the tcgen05 kernel covers the ReLU, width-128 network")
        self.progress = nn.Parameter(torch.tensor(0.))
        self.output_range = float(torch.as_tensor(range_shape).max())
        dims = [3] + [hidden_num] * (num_hidden_layers + 1) + [4]
        blocks = []
        for i in range(len(dims) - 1):
            lin = nn.Linear(dims[i], dims[i + 1])
            nn.init.kaiming_normal_(lin.weight, a=0.0, nonlinearity='relu', mode='fan_in')
            last = i == len(dims) - 2
            if last:
                nn.init.zeros_(lin.weight)
                nn.init.zeros_(lin.bias)
            blocks.append(nn.Sequential(lin) if last else nn.Sequential(lin, nn.ReLU(inplace=True)))
        self.deform_net = _Holder(_Holder(nn.Sequential(*blocks)))

    def linears(self):
        return [blk[0] for blk in self.deform_net.net.net]

    def parameters_for_kernel(self):
        lin = self.linears()
        return [q.weight for q in lin], [q.bias for q in lin]

    @torch.no_grad()
    def forward_with_jacobian(self, model_input):
        """x [n,3] (CUDA) -> deformation [n,3], correction [n,1], jac [n,4,3] = d (deformation, correction) / d x."""
        ws, bs = self.parameters_for_kernel()
        return ops.warpnet_forward(model_input, ws, bs, self.output_range)

    def forward(self, model_input):
        """-> (deformation [n,3], correction [n,1]), deform_net.py:23-29"""
        d, c, _ = self.forward_with_jacobian(model_input)
        return d, c
