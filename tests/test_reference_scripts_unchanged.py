"""INTEGRATION option B on the reference's OWN script code: `train_insilico.training_step` (train_insilico.py:124-140) is
imported UNMODIFIED from /root/reference with the module names `odenet` and `torchdiffeq` shadowed by phoenix_b200's
mirrors (and matplotlib stubbed: it is not installed here and only draws plots), then run for one optimiser step and
compared with the same function run over the reference's own modules.

/root/reference exists only in the build container, which has no GPU, so the CUDA engine underneath the mirrors is
replaced for this test by the CPU oracle (test infrastructure): what is exercised is everything ABOVE the C ABI -- the
import surface, `ODENet` attribute / method names, `odeint_adjoint` argument normalisation, the autograd plumbing that
lands the six gradients in `.grad`, `prior_only_forward`, the Adam parameter groups.  The kernels themselves are checked
against the same oracle on the GPU (tests/test_gpu_*.py)."""
import importlib
import os
import sys
import types

import pytest
import torch

from oracle import phoenix_oracle as O

REF = "/root/reference/ode_net/code"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="needs the reference checkout (build container only)")


class FakeDataHandler:
    """get_batch of a `single`-type batch (datahandler.py:95-120): (y_i [B,1,G], t [B,2], y_{i+1} [B,1,G])."""

    def __init__(self, batch, t, target):
        self.device = "cpu"
        self._b = (batch, t, target)

    def get_batch(self, batch_size):
        return self._b


def _import_training_step(shadow):
    saved = {k: sys.modules.get(k) for k in list(sys.modules)
             if k in ("odenet", "train_insilico", "datahandler", "csvreader", "visualization", "figure_saver",
                      "read_config") or k == "torchdiffeq" or k.startswith("torchdiffeq.")}
    for k in saved:
        sys.modules.pop(k, None)
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.lines", "matplotlib.patches", "matplotlib.font_manager",
                 "mpl_toolkits", "mpl_toolkits.axes_grid1"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["matplotlib.lines"].Line2D = object
    sys.modules["mpl_toolkits.axes_grid1"].make_axes_locatable = object
    sys.path.insert(0, REF)
    argv, sys.argv = sys.argv, ["train_insilico.py"]     # the script parses its command line at import time (:149-154)
    try:
        if shadow:
            import phoenix_b200.odenet as pod
            import phoenix_b200.torchdiffeq as ptd
            sys.modules["odenet"] = pod
            sys.modules["torchdiffeq"] = ptd
            sys.modules["torchdiffeq.__init__"] = ptd
        ti = importlib.import_module("train_insilico")
        return ti, sys.modules["odenet"].ODENet
    finally:
        sys.argv = argv
        sys.path.remove(REF)
        for k in ("odenet", "train_insilico", "datahandler", "csvreader", "visualization", "figure_saver", "read_config",
                  "torchdiffeq", "torchdiffeq.__init__"):
            sys.modules.pop(k, None)
        for k in list(sys.modules):
            if k.startswith("torchdiffeq."):
                sys.modules.pop(k, None)
        for k, v in saved.items():
            if v is not None:
                sys.modules[k] = v


def _oracle_engine(monkeypatch):
    """phoenix_b200.engine entry points on the CPU oracle (same signatures, same return conventions)."""
    from phoenix_b200 import engine

    def weights(net):
        return O.Weights(*[p.detach() for p in engine.net_params(net)])

    def solve_forward(net, y0, t_list, t_is_f32, reversed_time, method, rtol, atol, max_num_steps):
        t = torch.tensor(t_list, dtype=torch.float32 if t_is_f32 else torch.float64)
        return O.odeint(weights(net), y0.detach(), -t if reversed_time else t, method=method, rtol=rtol, atol=atol)[0]

    def solve_adjoint(net, t_list, t_is_f32, method, rtol, atol, max_num_steps, y_saved, grad_y):
        t = torch.tensor(t_list, dtype=torch.float32 if t_is_f32 else torch.float64)
        ady, grads, _ = O.adjoint_backward(weights(net), t, y_saved.detach(), grad_y.detach(), method=method, rtol=rtol,
                                           atol=atol)
        return ady, grads

    def solve_adjoint_many(net, t_rows, t_is_f32, method, rtol, atol, max_num_steps, y_saved, grad_y):
        # the sibling nodes of one backward pass share one call (_api.py, "backward side"): sum over the samples
        parts = [solve_adjoint(net, t_rows[i], t_is_f32, method, rtol, atol, max_num_steps, y_saved[i], grad_y[i])
                 for i in range(len(t_rows))]
        return torch.stack([p[0] for p in parts]), [sum(p[1][k] for p in parts) for k in range(6)]

    def rhs_forward(net, y, decay):
        return O.rhs(weights(net), y.detach(), decay=bool(decay))

    def rhs_vjp(net, y, g, decay, need_ybar=True, need_grads=True):
        _, ybar, pbar = O.rhs_vjp(weights(net), y.detach(), g.detach(), decay=bool(decay))
        return ybar, pbar

    for name, fn in (("solve_forward", solve_forward), ("solve_adjoint", solve_adjoint),
                     ("solve_adjoint_many", solve_adjoint_many), ("rhs_forward", rhs_forward),
                     ("rhs_vjp", rhs_vjp)):
        monkeypatch.setattr(engine, name, fn)


def _one_step(ti, ODENet, method):
    G, H, B, K = 61, 9, 3, 40
    torch.manual_seed(5)
    net = ODENet("cpu", G, explicit_time=False, neurons=H)
    w = O.make_weights(G, H, 77, dense=True)
    with torch.no_grad():
        net.gene_multipliers.copy_(w.gene_multipliers)
        net.net_prods.linear_out.weight.copy_(w.Wp)
        net.net_prods.linear_out.bias.copy_(w.bp)
        net.net_sums.linear_out.weight.copy_(w.Ws)
        net.net_sums.linear_out.bias.copy_(w.bs)
        net.net_alpha_combine.linear_out.weight.copy_(w.Wa)
    opt = torch.optim.Adam([
        {'params': net.net_sums.linear_out.weight}, {'params': net.net_sums.linear_out.bias},
        {'params': net.net_prods.linear_out.weight}, {'params': net.net_prods.linear_out.bias},
        {'params': net.net_alpha_combine.linear_out.weight},
        {'params': net.gene_multipliers, 'lr': 5 * 1e-3}], lr=1e-3, weight_decay=0.0)
    gen = torch.Generator().manual_seed(6)
    batch = torch.rand(B, 1, G, generator=gen)
    target = torch.rand(B, 1, G, generator=gen)
    t = torch.tensor([[0.0, 2.0], [2.0, 3.0], [3.0, 7.0]])
    bfp = torch.rand(K, 1, G, generator=gen) - 0.5
    pgrad = torch.randn(K, 1, G, generator=gen) * 0.1
    losses = ti.training_step(net, FakeDataHandler(batch, t, target), opt, method, B, False, False, bfp, pgrad, 0.99)
    return [float(x) for x in losses], [p.detach().clone() for p in net.parameters()]


@pytest.mark.parametrize("method", ["rk4", "dopri5"])
def test_unmodified_training_step_over_the_shadowed_modules(monkeypatch, method):
    ti_ref, ODENet_ref = _import_training_step(shadow=False)
    ref_losses, ref_params = _one_step(ti_ref, ODENet_ref, method)
    _oracle_engine(monkeypatch)
    ti_new, ODENet_new = _import_training_step(shadow=True)
    import phoenix_b200
    assert ODENet_new is phoenix_b200.ODENet and ti_new.odeint is phoenix_b200.odeint_adjoint
    new_losses, new_params = _one_step(ti_new, ODENet_new, method)
    for a, b in zip(new_losses, ref_losses):
        assert abs(a - b) <= 1e-5 * abs(b)
    # the parameters AFTER the Adam step: the update is lr * g / (|g| + eps) on the first step, so agreement of the
    # updated weights to 1e-6 absolute (lr = 1e-3) means the gradients agree in sign and are not lost anywhere
    for a, b in zip(new_params, ref_params):
        assert float((a - b).abs().max()) <= 2e-6
