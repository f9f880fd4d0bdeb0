"""Drop-in mirror of the reference's ``odenet`` module (ode_net/code/odenet.py) for the B200 path.

Same class names (pickles written by ``ODENet.save`` reference ``odenet.SoftsignMod`` / ``LogShiftedSoftSignMod``),
same constructor, attributes, parameter order, init distribution and four-file checkpoint format; ``forward`` /
``prior_only_forward`` run the fused CUDA kernels of libphoenix_b200.so (no ATen math on the hot path, no CPU
fallback).  Differences, both papering over reference defects:
  * every parameter is really placed on ``device`` (the reference discards ``gene_multipliers.to(device)``,
    odenet.py:80, and its ``to()`` refers to a non-existent ``self.net``, odenet.py:150-151);
  * ``load_model`` keeps the loaded modules on the model's device instead of forcing 'cpu'.
"""
import sys

import torch
import torch.nn as nn

from . import engine


class SoftsignMod(nn.Module):
    """(x - 0.5) / (1 + |x - 0.5|)   (odenet.py:16-25).  Kept as a module so checkpoints stay loadable; the fused
    kernels evaluate it inside the operand load, this ``forward`` is only used if someone calls the module itself."""

    def __init__(self):
        super().__init__()

    def forward(self, input):
        z = input - 0.5
        return z / (1 + torch.abs(z))


class LogShiftedSoftSignMod(nn.Module):
    """log1p((x - 0.5) / (1 + |x - 0.5|))   (odenet.py:27-35)."""

    def __init__(self):
        super().__init__()

    def forward(self, input):
        z = input - 0.5
        return torch.log1p(z / (1 + torch.abs(z)))


class _RHSFunction(torch.autograd.Function):
    """ODENet.forward / prior_only_forward as one fused CUDA op with its hand-written VJP (odenet.py:85-98)."""

    @staticmethod
    def forward(ctx, net, decay, y, *params):
        ctx.net = net
        ctx.decay = decay
        ctx.save_for_backward(y)
        return engine.rhs_forward(net, y, decay).view_as(y)

    @staticmethod
    def backward(ctx, g):
        (y,) = ctx.saved_tensors
        need_y = ctx.needs_input_grad[2]
        need_p = any(ctx.needs_input_grad[3:])
        ybar, grads = engine.rhs_vjp(ctx.net, y, g, ctx.decay, need_ybar=need_y, need_grads=need_p)
        out = [None, None, ybar.view_as(y) if need_y else None]
        for i, need in enumerate(ctx.needs_input_grad[3:]):
            out.append(grads[i] if (need and grads is not None) else None)
        return tuple(out)


class ODENet(nn.Module):
    ''' ODE-Net class implementation (B200-native) '''

    def __init__(self, device, ndim, explicit_time=False, neurons=100):
        ''' Initialize a new ODE-Net '''
        super(ODENet, self).__init__()
        self.ndim = ndim
        self.explicit_time = explicit_time
        if explicit_time:
            raise NotImplementedError("explicit_time=True is never used by the reference (read_config.py hard-wires "
                                      "False) and is outside the accelerated path")

        self.net_prods = nn.Sequential()
        self.net_prods.add_module('activation_0', LogShiftedSoftSignMod())
        self.net_prods.add_module('linear_out', nn.Linear(ndim, neurons, bias=True))

        self.net_sums = nn.Sequential()
        self.net_sums.add_module('activation_0', SoftsignMod())
        self.net_sums.add_module('linear_out', nn.Linear(ndim, neurons, bias=True))

        self.net_alpha_combine = nn.Sequential()
        self.net_alpha_combine.add_module('linear_out', nn.Linear(2 * neurons, ndim, bias=False))

        self.gene_multipliers = nn.Parameter(torch.rand(1, ndim), requires_grad=True)

        # same init as odenet.py:64-75
        for seq in (self.net_sums, self.net_prods, self.net_alpha_combine):
            for n in seq.modules():
                if isinstance(n, nn.Linear):
                    nn.init.sparse_(n.weight, sparsity=0.95, std=0.05)

        self._device = torch.device(device)
        nn.Module.to(self, self._device)

    def forward(self, t, y):
        # sums/prods branches, their learned combination and the decay term, fused in libphoenix_b200 (odenet.py:85-91)
        final = _RHSFunction.apply(self, True, y, *engine.net_params(self))
        return(final)

    def prior_only_forward(self, t, y):
        # the combination before the decay term (odenet.py:93-98)
        joint = _RHSFunction.apply(self, False, y, *engine.net_params(self))
        return(joint)

    def save(self, fp):
        ''' Save the model to file (four pickles, odenet.py:100-111) '''
        idx = fp.index('.')
        alpha_comb_path = fp[:idx] + '_alpha_comb' + fp[idx:]
        gene_mult_path = fp[:idx] + '_gene_multipliers' + fp[idx:]
        prod_path = fp[:idx] + '_prods' + fp[idx:]
        sum_path = fp[:idx] + '_sums' + fp[idx:]
        torch.save(self.net_prods, prod_path)
        torch.save(self.net_sums, sum_path)
        torch.save(self.net_alpha_combine, alpha_comb_path)
        torch.save(self.gene_multipliers, gene_mult_path)

    def load_dict(self, fp):
        ''' Load a model from a dict file '''
        self.load_state_dict(torch.load(fp))

    def load_model(self, fp):
        ''' Load a model from a file (odenet.py:118-133) '''
        idx = fp.index('.pt')
        gene_mult_path = fp[:idx] + '_gene_multipliers' + fp[idx:]
        prod_path = fp[:idx] + '_prods' + fp[idx:]
        sum_path = fp[:idx] + '_sums' + fp[idx:]
        alpha_comb_path = fp[:idx] + '_alpha_comb' + fp[idx:]
        dev = self.gene_multipliers.device
        self.net_prods = torch.load(prod_path, weights_only=False, map_location=dev)
        self.net_sums = torch.load(sum_path, weights_only=False, map_location=dev)
        self.gene_multipliers = torch.load(gene_mult_path, weights_only=False, map_location=dev)
        self.net_alpha_combine = torch.load(alpha_comb_path, weights_only=False, map_location=dev)

    def load(self, fp):
        ''' General loading from a file '''
        try:
            print('Trying to load model from file= {}'.format(fp))
            self.load_model(fp)
            print('Done')
        except Exception:
            print('Failed! Trying to load parameters from file...')
            try:
                self.load_dict(fp)
                print('Done')
            except Exception:
                print('Failed! Network structure is not correct, cannot load parameters from file, exiting!')
                sys.exit(0)
