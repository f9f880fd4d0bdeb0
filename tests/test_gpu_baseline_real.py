"""GPU parity at the BASELINE configs (VERDICT r01 "untested BASELINE configs"):

  * C3 yeast (3 551 genes x 120) on the real 24-point Pramila series, dt = 5 and 10, batch 4, init-distribution weights,
    dopri5 and rk4;  C4 breast (11 165 genes, H = 200 and 40) on real Desmedt rows, dt = 0.0051, dopri5 -- against goldens
    produced by the unmodified reference (tests/golden/make_golden_big.py), through the per-sample loop the scripts
    issue (train_insilico.py:128-130) and through odeint_adjoint_many;
  * C5 (20 000 genes x 200): RHS-VJP, rk4 solve + adjoint and dopri5 forward against the oracle on a 128-row batch
    (the dopri5 norm is global over the rows of a batched call, misc.py:10-11, so oracle and CUDA path see the same
    rows), and the 4 096-row call against its own 128-row slices (rows of a fixed-grid solve are independent).

Tolerances: fixed-step y / loss / gradients rel-L2 <= 1e-5; dopri5 y <= 1e-5, gradients <= 5e-5 where the reference
reproduces its own step sequence (C4), 2e-4 where it does not (C3: `stable` = 0, the reference's own 1-thread vs
8-thread gradients differ by `self_grad_rel`)."""
import pytest
import torch

from golden_util import (assert_logs_close, big_case_inputs, check_big_case, compare_logs, load, manifest_big, rel_l2)
from oracle import phoenix_oracle as O

pytestmark = pytest.mark.gpu
BIG = manifest_big()


@pytest.fixture(scope="module")
def pb():
    import phoenix_b200 as pb
    pb.set_sync_errors(True)
    yield pb
    pb.set_step_logging(False)
    pb.set_sync_errors(False)


def make_net(pb, w):
    net = pb.ODENet("cuda", w.G, neurons=w.H)
    with torch.no_grad():
        for p, src in zip(net.parameters(), w.as_list()):
            p.copy_(src)
    return net


@pytest.mark.parametrize("m", BIG, ids=[m["name"] for m in BIG])
def test_real_data_goldens_per_sample_loop_and_many(pb, m):
    d = load(m["name"])
    w, y0, t, target = big_case_inputs(m, d)
    net = make_net(pb, w)
    dop = m["method"] == "dopri5"
    ytol = 1e-5
    gtol = (5e-5 if m["stable"] else 2e-4) if dop else 1e-5
    # (1) the literal loop of training_step, with step logs
    pb.set_step_logging(True)
    net.zero_grad()
    y0g = y0.cuda().requires_grad_(True)
    preds, flogs = [], []
    for i in range(m["N"]):
        preds.append(pb.odeint_adjoint(net, y0g[i], t[i], method=m["method"])[1])
        flogs.append(pb.last_step_log())
    pred = torch.stack(preds)
    loss = torch.mean((pred - target.cuda()) ** 2)
    loss.backward()
    pb.set_step_logging(False)
    check_big_case(m, d, pred.detach().cpu(), loss.item(), y0g.grad.cpu(), [p.grad for p in net.parameters()], ytol, gtol)
    if dop:
        for i in range(m["N"]):
            if m["stable"]:
                # one attempted step per sample (dt = 0.0051 is below the first step): identical sequence, dt to 1e-3
                assert compare_logs(flogs[i], d["flog%d" % i], 1e-3)[1] == "identical", (m["name"], i)
            else:
                assert_logs_close(flogs[i], d["flog%d" % i], 0.25, "%s fwd %d" % (m["name"], i))
    # (2) the same samples through ONE odeint_adjoint_many call
    net.zero_grad()
    y0m = y0.cuda().requires_grad_(True)
    many = pb.odeint_adjoint_many(net, y0m, t, method=m["method"])
    lossm = torch.mean((many[:, 1] - target.cuda()) ** 2)
    lossm.backward()
    check_big_case(m, d, many[:, 1].detach().cpu(), lossm.item(), y0m.grad.cpu(), [p.grad for p in net.parameters()],
                   ytol, gtol)


def test_real_data_adjoint_step_logs_c4(pb):
    """C4, H = 200: the adjoint sweep of every sample is ONE attempted step in the reference; so is ours, same dt."""
    m = [x for x in BIG if x["name"] == "c4_breast_real_h200_dopri5"][0]
    d = load(m["name"])
    w, y0, t, target = big_case_inputs(m, d)
    net = make_net(pb, w)
    pb.set_step_logging(True)
    try:
        for i in range(m["N"]):
            net.zero_grad()
            y0g = y0[i].cuda().requires_grad_(True)
            y = pb.odeint_adjoint(net, y0g, t[i], method="dopri5")
            # the golden's loss is the mean over all N samples' elements
            (torch.sum((y[1] - target[i].cuda()) ** 2) / target.numel()).backward()
            blog = pb.last_step_log()
            assert compare_logs(blog, d["blog%d" % i], 1e-3)[1] == "identical", (i, blog, d["blog%d" % i])
    finally:
        pb.set_step_logging(False)


# ---- C5: 20 000 genes x 200 hidden units -------------------------------------------------------------------------------
G5, H5, B5 = 20000, 200, 128


@pytest.fixture(scope="module")
def c5(pb):
    w = O.make_weights(G5, H5, 5001, dense=False)
    gen = torch.Generator().manual_seed(5002)
    y = torch.rand(B5, G5, generator=gen)
    g = torch.randn(B5, G5, generator=gen)
    return w, make_net(pb, w), y, g


def test_c5_rhs_vjp_against_oracle(pb, c5):
    w, net, y, g = c5
    net.zero_grad()
    yg = y.cuda().requires_grad_(True)
    f = net(None, yg)
    f.backward(g.cuda())
    f_ref, ybar_ref, pbar_ref = O.rhs_vjp(w, y, g, decay=True)
    assert rel_l2(f.detach().cpu(), f_ref) < 1e-5
    assert rel_l2(yg.grad.cpu(), ybar_ref) < 1e-5
    for i, (p, ref) in enumerate(zip(net.parameters(), pbar_ref)):
        assert rel_l2(p.grad.cpu(), ref) < 2e-5, (i, rel_l2(p.grad.cpu(), ref))


def test_c5_rk4_solve_and_adjoint_against_oracle(pb, c5):
    w, net, y0, _ = c5
    t = torch.tensor([0.0, 0.1])
    target = torch.rand(B5, G5, generator=torch.Generator().manual_seed(5003))
    y_ref, _ = O.odeint(w, y0, t, method="rk4")
    gy = torch.zeros_like(y_ref)
    gy[1] = 2.0 * (y_ref[1] - target) / target.numel()
    ady_ref, g_ref, _ = O.adjoint_backward(w, t, y_ref, gy, method="rk4")
    net.zero_grad()
    y0g = y0.cuda().requires_grad_(True)
    y = pb.odeint_adjoint(net, y0g, t, method="rk4")
    torch.mean((y[1] - target.cuda()) ** 2).backward()
    assert rel_l2(y.detach().cpu(), y_ref) < 1e-5
    assert rel_l2(y0g.grad.cpu(), ady_ref) < 1e-5
    for i, (p, ref) in enumerate(zip(net.parameters(), g_ref)):
        assert rel_l2(p.grad.cpu(), ref) < 2e-5, (i, rel_l2(p.grad.cpu(), ref))
    # the full 4 096-row call: its first 128 rows are the same solve (fixed grid: rows are independent)
    gen = torch.Generator().manual_seed(5004)
    big = torch.cat([y0, torch.rand(4096 - B5, G5, generator=gen)]).cuda()
    with torch.no_grad():
        yb = pb.odeint(net, big, t, method="rk4")
    assert rel_l2(yb[1, :B5].cpu(), y_ref[1]) < 1e-5
    assert torch.isfinite(yb).all()


def test_c5_dopri5_forward_against_oracle(pb, c5):
    """Reference default rtol 1e-7 / atol 1e-9, global RMS norm over the 128 rows on both sides."""
    w, net, y0, _ = c5
    t = torch.tensor([0.0, 0.4], dtype=torch.float64)
    y_ref, flog = O.odeint(w, y0, t, method="dopri5")
    pb.set_step_logging(True)
    try:
        with torch.no_grad():
            y = pb.odeint(net, y0.cuda(), t, method="dopri5")
        mine = pb.last_step_log()
    finally:
        pb.set_step_logging(False)
    assert rel_l2(y.cpu(), y_ref) < 1e-5
    assert_logs_close(mine, flog.steps, 0.25, "c5 dopri5 fwd")
