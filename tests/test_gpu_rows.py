"""GPU tests of the rows kernels (phx_rows.cuh): N independent one-row problems as the rows of one persistent launch.

  * every method / several output times / ragged shapes against the CPU oracle, sample by sample;
  * a problem's results are bit-identical whether it runs alone or beside others (any position in a pass, any N);
  * the speculative final-step accumulation and its fall-back give the same gradient sum as single calls;
  * a failing problem is reported for that problem only and its unwritten outputs are NaN;
  * engine.invalidate after an in-place `.data` edit, lazy error mode, model on a non-current device."""
import pytest
import torch

from golden_util import rel_l2
from oracle import phoenix_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pb():
    import phoenix_b200 as pb
    pb.set_sync_errors(True)
    # this file is about the rows kernels: one-sample calls go through them too (by default a one-sample call of a
    # GPU-filling model takes the one-problem resident kernels, engine.SINGLE_CALL_ENGINE), which is what makes
    # "one call at a time" bit-identical to odeint_adjoint_many
    saved = pb.engine.SINGLE_CALL_ENGINE
    pb.engine.SINGLE_CALL_ENGINE = "rows"
    yield pb
    pb.engine.SINGLE_CALL_ENGINE = saved
    pb.set_sync_errors(True)


def make_net(pb, w, device="cuda"):
    net = pb.ODENet(device, w.G, neurons=w.H)
    with torch.no_grad():
        for p, src in zip(net.parameters(), w.as_list()):
            p.copy_(src)
    return net


def _problem(G, H, N, T, dt, seed, dense=True):
    w = O.make_weights(G, H, seed, dense=dense, neg_mult_frac=0.1)
    gen = torch.Generator().manual_seed(seed + 1)
    y0 = torch.rand(N, 1, G, generator=gen)
    tau = torch.rand(N, generator=gen)
    t = torch.stack([tau + dt * i * (1 + 0.3 * i) for i in range(T)], dim=1)
    target = torch.rand(N, 1, G, generator=gen)
    return w, y0, t, target


def _oracle(w, y0, t, target, method):
    ys, adys, gsum = [], [], None
    for i in range(y0.shape[0]):
        y, _ = O.odeint(w, y0[i], t[i], method=method)
        gy = torch.zeros_like(y)
        gy[-1] = 2.0 * (y[-1] - target[i]) / target.numel()
        ady, grads, _ = O.adjoint_backward(w, t[i], y, gy, method=method)
        ys.append(y)
        adys.append(ady)
        gsum = grads if gsum is None else [a + b for a, b in zip(gsum, grads)]
    return torch.stack(ys), torch.stack(adys), gsum


@pytest.mark.parametrize("G,H,N,T,method,dt,tol", [
    (37, 5, 1, 2, "euler", 0.7, 1e-5), (37, 5, 3, 3, "euler", 0.4, 1e-5), (129, 33, 2, 3, "midpoint", 0.3, 1e-5),
    (129, 33, 5, 3, "rk4", 0.3, 1e-5), (350, 40, 4, 2, "dopri5", 1.0, 2e-4), (350, 40, 6, 3, "dopri5", 0.5, 2e-4),
    (1001, 100, 3, 2, "rk4", 0.2, 1e-5), (3551, 120, 5, 2, "rk4", 0.5, 1e-5), (3551, 120, 4, 2, "dopri5", 0.5, 4e-4),
    (11165, 40, 3, 2, "dopri5", 0.0051, 5e-5), (11165, 200, 6, 2, "dopri5", 0.0051, 5e-5),
    (11165, 200, 2, 2, "dopri5", 0.3, 2e-4),
])
def test_rows_against_oracle_and_single_calls(pb, G, H, N, T, method, dt, tol):
    """tol: relative L2 of y / adj_y0 / the six gradient sums (fixed grid 1e-5; dopri5 follows its own, noise-decided
    step sequence: 5e-5 for single-step solves, 2e-4 .. 4e-4 over tens of steps, see profiles/r02_step_sequence_noise.txt)."""
    w, y0, t, target = _problem(G, H, N, T, dt, 7000 + G + N, dense=G < 5000)
    net = make_net(pb, w)
    y_ref, ady_ref, g_ref = _oracle(w, y0, t, target, method)
    ya = y0.cuda().requires_grad_(True)
    many = pb.odeint_adjoint_many(net, ya, t, method=method)
    torch.mean((many[:, -1] - target.cuda()) ** 2).backward()
    assert rel_l2(many.detach().cpu(), y_ref) < min(tol, 1e-5)
    assert rel_l2(ya.grad.cpu(), ady_ref) < tol
    g_many = [p.grad.detach().clone() for p in net.parameters()]
    for i, (g, r) in enumerate(zip(g_many, g_ref)):
        assert rel_l2(g.cpu(), r) < tol, (i, rel_l2(g.cpu(), r))
    # the same problems one call at a time: bit-identical trajectories and state cotangents, the gradient sum up to the
    # order of the cross-sample additions
    net.zero_grad()
    yb = y0.cuda().requires_grad_(True)
    one = torch.stack([pb.odeint_adjoint(net, yb[i], t[i], method=method) for i in range(N)])
    torch.mean((one[:, -1] - target.cuda()) ** 2).backward()
    assert torch.equal(one, many)
    assert torch.equal(yb.grad, ya.grad)
    for g, p in zip(g_many, net.parameters()):
        assert rel_l2(p.grad.cpu(), g.cpu()) < 2e-6


def test_row_position_does_not_change_a_problem(pb):
    """Problem 0 of a 5-problem call re-run as problem 3 of a 7-problem call (other neighbours, another pass, another row):
    identical bits."""
    w, y0, t, _ = _problem(690, 40, 7, 2, 1.0, 7100)
    net = make_net(pb, w)
    with torch.no_grad():
        a = pb.odeint_adjoint_many(net, y0[:5].cuda(), t[:5], method="dopri5")
        perm = [4, 2, 6, 0, 1, 5, 3]
        b = pb.odeint_adjoint_many(net, y0[perm].cuda(), t[perm], method="dopri5")
    assert torch.equal(a[0], b[3]) and torch.equal(a[2], b[1]) and torch.equal(a[4], b[0])


def test_speculative_sum_fallback_on_a_rejected_final_step(pb):
    """Rows that finish at different steps and rows whose last step is rejected take the fall-back of the speculative
    final-step accumulation: the gradient sum must still equal the sum of single calls.  Long intervals with dense weights
    give both rejected steps and different step counts per row (checked through the status records)."""
    w, y0, t, target = _problem(350, 40, 8, 2, 3.0, 7200)
    net = make_net(pb, w)
    ya = y0.cuda().requires_grad_(True)
    many = pb.odeint_adjoint_many(net, ya, t, method="dopri5")
    torch.mean((many[:, -1] - target.cuda()) ** 2).backward()
    g_many = [p.grad.detach().clone() for p in net.parameters()]
    net.zero_grad()
    yb = y0.cuda().requires_grad_(True)
    outs = []
    for i in range(8):
        outs.append(pb.odeint_adjoint(net, yb[i], t[i], method="dopri5"))
    torch.mean((torch.stack(outs)[:, -1] - target.cuda()) ** 2).backward()
    for g, p in zip(g_many, net.parameters()):
        assert rel_l2(p.grad.cpu(), g.cpu()) < 2e-6
    assert torch.equal(yb.grad, ya.grad)


def test_failing_row_is_reported_alone_and_reads_nan(pb):
    w, y0, t, _ = _problem(3551, 120, 6, 2, 0.3, 7300)
    net = make_net(pb, w)
    bad = y0.clone()
    bad[4, 0, 11] = float("inf")
    pb.set_sync_errors(False)          # lazy mode: the call returns, the failure surfaces at the next call
    try:
        with torch.no_grad():
            out = pb.odeint_adjoint_many(net, bad.cuda(), t, method="dopri5")
        torch.cuda.synchronize()
        assert torch.isnan(out[4, 1]).all()                     # never reached: NaN, not uninitialised memory
        assert torch.isfinite(out[[0, 1, 2, 3, 5]]).all()
        with pytest.raises(AssertionError):
            pb.check_errors()
    finally:
        pb.set_sync_errors(True)
    with torch.no_grad():
        ref = pb.odeint(net, y0[5].cuda(), t[5], method="dopri5")
    assert torch.equal(out[5], ref)
    with pytest.raises(AssertionError):                         # default mode: raised at the call site
        with torch.no_grad():
            pb.odeint_adjoint_many(net, bad.cuda(), t, method="dopri5")


def test_invalidate_after_data_edit(pb):
    """`.data` edits do not bump the autograd version counter the packed-weight cache is keyed on:
    engine.invalidate(net) must make the next solve see the new weights."""
    w, y0, t, _ = _problem(350, 40, 1, 2, 0.5, 7400)
    net = make_net(pb, w)
    with torch.no_grad():
        y_a = pb.odeint(net, y0[0].cuda(), t[0], method="rk4")
    net.net_alpha_combine.linear_out.weight.data.mul_(0.5)
    pb.engine.invalidate(net)
    with torch.no_grad():
        y_b = pb.odeint(net, y0[0].cuda(), t[0], method="rk4")
    w2 = O.Weights(w.gene_multipliers, w.Wp, w.bp, w.Ws, w.bs, w.Wa * 0.5)
    y_ref, _ = O.odeint(w2, y0[0], t[0], method="rk4")
    assert rel_l2(y_b.cpu(), y_ref) < 1e-5
    assert rel_l2(y_a.cpu(), y_ref) > 1e-4


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_model_on_a_device_other_than_the_current_one(pb):
    w, y0, t, target = _problem(350, 40, 3, 2, 0.5, 7500)
    torch.cuda.set_device(0)
    net = make_net(pb, w, "cuda:1")
    ya = y0.to("cuda:1").requires_grad_(True)
    many = pb.odeint_adjoint_many(net, ya, t, method="dopri5")
    torch.mean((many[:, -1] - target.to("cuda:1")) ** 2).backward()
    y_ref, ady_ref, g_ref = _oracle(w, y0, t, target, "dopri5")
    assert rel_l2(many.detach().cpu(), y_ref) < 1e-5
    for g, r in zip([p.grad for p in net.parameters()], g_ref):
        assert rel_l2(g.cpu(), r) < 2e-4
    f = net(None, y0[:, 0].to("cuda:1"))
    assert rel_l2(f.cpu(), O.rhs(w, y0[:, 0])) < 1e-5
    assert torch.cuda.current_device() == 0
