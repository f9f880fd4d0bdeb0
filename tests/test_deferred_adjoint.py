"""The per-sample ``odeint_adjoint`` nodes of one backward pass hand their cotangents to the last of them, which solves
all the sweeps in one lock-step call (phoenix_b200/torchdiffeq/_api.py, "the per-sample loop of training_step, backward
side").  CPU part: the bookkeeping, with the three engine entry points replaced by a closed-form toy (no kernels run);
GPU part: the real kernels, deferred against one-sweep-per-node."""
import pytest
import torch

import phoenix_b200 as pb
from phoenix_b200 import engine
from phoenix_b200.torchdiffeq import _api


class _Toy:
    """y(t) = y0 * (1 + c t) with c = mean of the multipliers; parameter cotangents are simple functions of (y, grad_y) so
    that a wrong pairing, a missing or a doubled sample changes the result."""

    def __init__(self):
        self.single = self.many = 0
        self.many_sizes = []

    def forward(self, net, y0, tl, t_is_f32, rev, method, rtol, atol, max_steps):
        c = float(engine.net_params(net)[0].mean())
        return torch.stack([y0 * (1.0 + c * t) for t in tl])

    def _one(self, net, tl, y, gy):
        adj = (gy * torch.tensor(tl, dtype=torch.float32).view(-1, *[1] * (gy.dim() - 1))).sum(0)
        s = float((y * gy).sum())
        return adj, [torch.full_like(p, s * (i + 1)) for i, p in enumerate(engine.net_params(net))]

    def adjoint(self, net, tl, t_is_f32, method, rtol, atol, max_steps, y, gy):
        self.single += 1
        return self._one(net, tl, y, gy)

    def adjoint_many(self, net, t_rows, t_is_f32, method, rtol, atol, max_steps, ys, gys):
        self.many += 1
        self.many_sizes.append(len(t_rows))
        assert ys.shape[0] == gys.shape[0] == len(t_rows)
        parts = [self._one(net, t_rows[i], ys[i], gys[i]) for i in range(len(t_rows))]
        grads = [sum(p[1][k] for p in parts) for k in range(6)]
        return torch.stack([p[0] for p in parts]), grads


@pytest.fixture
def toy(monkeypatch):
    t = _Toy()
    monkeypatch.setattr(engine, "solve_forward", t.forward)
    monkeypatch.setattr(engine, "solve_adjoint", t.adjoint)
    monkeypatch.setattr(engine, "solve_adjoint_many", t.adjoint_many)
    yield t
    _api.set_deferred_adjoint(True)


def _net():
    torch.manual_seed(0)
    return pb.ODENet("cpu", 12, neurons=4)


def _loss(net, y0s, ts, idx=None):
    preds = [pb.odeint_adjoint(net, y0, t, method="rk4")[-1] for y0, t in zip(y0s, ts)]
    if idx is not None:
        preds = [preds[i] for i in idx]
    w = torch.arange(1, len(preds) + 1, dtype=torch.float32).view(-1, 1, 1)
    return (torch.stack(preds) * w).pow(2).sum()


def _data(n, T=2, seed=1):
    g = torch.Generator().manual_seed(seed)
    y0s = [torch.rand(1, 12, generator=g) for _ in range(n)]
    ts = [torch.tensor([0.0] + [0.5 + 0.25 * i + j for j in range(T - 1)]) for i in range(n)]
    return y0s, ts


def _grads(net):
    return [p.grad.clone() for p in net.parameters()]


def test_backward_batches_the_sibling_nodes(toy):
    net = _net()
    y0s, ts = _data(5)
    _api.set_deferred_adjoint(False)
    _loss(net, y0s, ts).backward()
    ref = _grads(net)
    assert (toy.single, toy.many) == (5, 0)
    net.zero_grad()
    _api.set_deferred_adjoint(True)
    _loss(net, y0s, ts).backward()
    assert (toy.single, toy.many, toy.many_sizes) == (5, 1, [5])
    for a, b in zip(_grads(net), ref):
        torch.testing.assert_close(a, b, rtol=1e-6, atol=0)
    assert not _api._deferred


def test_only_the_nodes_of_this_backward_pass_are_counted(toy):
    net = _net()
    y0s, ts = _data(6)
    stale = pb.odeint_adjoint(net, y0s[0], ts[0], method="rk4")      # a live node that no backward pass reaches
    _loss(net, y0s, ts, idx=[0, 2, 5]).backward()                     # three of six nodes are in the graph of the loss
    assert toy.many_sizes == [3] and toy.single == 0
    got = _grads(net)
    net.zero_grad()
    _api.set_deferred_adjoint(False)
    _loss(net, y0s, ts, idx=[0, 2, 5]).backward()
    for a, b in zip(got, _grads(net)):
        torch.testing.assert_close(a, b, rtol=1e-6, atol=0)
    assert stale.grad_fn is not None and not _api._deferred


def test_autograd_grad_and_retained_graphs(toy):
    net = _net()
    y0s, ts = _data(4)
    params = list(net.parameters())
    loss = _loss(net, y0s, ts)
    g1 = torch.autograd.grad(loss, params, retain_graph=True)
    g2 = torch.autograd.grad(loss, params[3:5], retain_graph=True)     # a subset of the parameters
    loss.backward()
    assert toy.many_sizes == [4, 4, 4] and toy.single == 0
    assert all(p.grad is not None for p in params)
    for a, b in zip(g1, _grads(net)):
        torch.testing.assert_close(a, b, rtol=0, atol=0)
    for a, b in zip(g2, g1[3:5]):
        torch.testing.assert_close(a, b, rtol=0, atol=0)


def test_nodes_that_cannot_wait_run_their_own_sweep(toy):
    net = _net()
    y0s, ts = _data(3)
    y0s[1].requires_grad_(True)                                        # its adj_y0 is wanted: solved on the spot
    _loss(net, y0s, ts).backward()
    assert toy.single == 1 and toy.many_sizes == [2]
    assert y0s[1].grad is not None and y0s[0].grad is None
    # a lone node, and a two-row state, are not deferred either
    net.zero_grad()
    toy.single = toy.many = 0
    pb.odeint_adjoint(net, y0s[0], ts[0], method="rk4")[-1].sum().backward()
    yb = torch.rand(2, 12)
    (pb.odeint_adjoint(net, yb, ts[0], method="rk4")[-1].sum()
     + pb.odeint_adjoint(net, yb, ts[1], method="rk4")[-1].sum()).backward()
    assert (toy.single, toy.many) == (3, 0)


def test_groups_do_not_mix(toy):
    """Different numbers of output times (or solver settings) cannot share a lock-step call: one batch per group."""
    net = _net()
    y0s, ts = _data(4)
    y0b, tb = _data(2, T=3, seed=2)
    (_loss(net, y0s, ts) + _loss(net, y0b, tb)).backward()
    assert sorted(toy.many_sizes) == [2, 4] and toy.single == 0
    got = _grads(net)
    net.zero_grad()
    _api.set_deferred_adjoint(False)
    (_loss(net, y0s, ts) + _loss(net, y0b, tb)).backward()
    for a, b in zip(got, _grads(net)):
        torch.testing.assert_close(a, b, rtol=1e-6, atol=0)


def test_a_sibling_that_never_runs_is_reported(toy, monkeypatch):
    """If the count of expected siblings were ever wrong the waiting cotangents must not vanish silently."""
    net = _net()
    y0s, ts = _data(3)
    real = torch._C._will_engine_execute_node
    extra = pb.odeint_adjoint(net, y0s[0], ts[0], method="rk4")        # counted by the faulty predicate, never run

    monkeypatch.setattr(torch._C, "_will_engine_execute_node",
                        lambda node: True if node is extra.grad_fn else real(node))
    with pytest.raises(RuntimeError, match="never ran"):
        _loss(net, y0s, ts).backward()
    assert not _api._deferred


@pytest.mark.gpu
def test_deferred_adjoint_matches_one_sweep_per_node_on_the_gpu():
    torch.manual_seed(4)
    G, H, N = 1037, 56, 7
    net = pb.ODENet("cuda", G, neurons=H)
    y0 = torch.rand(N, 1, G, device="cuda")
    t = torch.tensor([[0.0, 0.4 + 0.1 * i] for i in range(N)], device="cuda")
    target = torch.rand(N, 1, G, device="cuda")

    def loop():
        net.zero_grad()
        preds = [pb.odeint_adjoint(net, y0[i], t[i], method="dopri5")[1] for i in range(N)]
        loss = torch.mean((torch.stack(preds) - target) ** 2)
        loss.backward()
        return float(loss), [p.grad.clone() for p in net.parameters()]

    _api.set_deferred_adjoint(False)
    try:
        l0, g0 = loop()
    finally:
        _api.set_deferred_adjoint(True)
    l1, g1 = loop()
    net.zero_grad()
    ym = pb.odeint_adjoint_many(net, y0, t, method="dopri5")
    torch.mean((ym[:, 1] - target) ** 2).backward()
    assert l0 == l1
    for a, b, p in zip(g1, g0, net.parameters()):
        assert float((a - b).norm() / b.norm()) < 2e-6      # the same sweeps, summed in another order
        assert torch.equal(a, p.grad)                       # exactly the backward of odeint_adjoint_many
    pb.check_errors()
