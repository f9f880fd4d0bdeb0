"""End-to-end parity of one optimiser step's gradients: the reference's `training_step` (train_insilico.py:124-140) --
per-sample `odeint` solves, data loss, prior-constrained loss through `prior_only_forward` on a batch of random rows,
`composed_loss.backward()` -- on the CUDA path (both the literal per-sample loop and `odeint_adjoint_many`) against
the same gradients assembled from the CPU oracle (adjoint sweeps + RHS VJP).  Tolerance: relative L2 <= 5e-5 per
parameter tensor (dopri5 at the reference default rtol 1e-7; the two loss terms are scaled like the reference's)."""
import pytest
import torch

from golden_util import rel_l2
from oracle import phoenix_oracle as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("method", ["dopri5", "rk4"])
def test_training_step_gradients_match_the_oracle(method):
    import phoenix_b200 as pb
    G, H, batch, K, lam = 350, 40, 4, 300, 0.99
    w = O.make_weights(G, H, 40, dense=True, neg_mult_frac=0.05)
    net = pb.ODENet("cuda", G, neurons=H)
    with torch.no_grad():
        for p, src in zip(net.parameters(), w.as_list()):
            p.copy_(src)
    gen = torch.Generator().manual_seed(41)
    y0 = torch.rand(batch, 1, G, generator=gen)
    target = torch.rand(batch, 1, G, generator=gen)
    tau = torch.rand(batch, generator=gen)
    t = torch.stack([tau, tau + 2.0], dim=1)
    batch_for_prior = torch.rand(K, 1, G, generator=gen) - 0.5
    prior_grad = torch.randn(K, 1, G, generator=gen) * 0.1

    # ---- oracle: d(composed_loss)/d(theta) = lam * sum_i adjoint_i + (1 - lam) * VJP of the prior term
    ref = [torch.zeros_like(p) for p in w.as_list()]
    loss_data = 0.0
    for i in range(batch):
        y, _ = O.odeint(w, y0[i], t[i], method=method)
        gy = torch.zeros_like(y)
        gy[1] = lam * 2.0 * (y[1] - target[i]) / (batch * G)
        _, grads, _ = O.adjoint_backward(w, t[i], y, gy, method=method)
        for r, g in zip(ref, grads):
            r += g.reshape(r.shape)
        loss_data += float(((y[1] - target[i]) ** 2).sum()) / (batch * G)
    J = O.rhs(w, batch_for_prior, decay=False)
    gJ = (1.0 - lam) * 2.0 * (J - prior_grad) / J.numel()
    _, _, pbar = O.rhs_vjp(w, batch_for_prior, gJ, decay=False)
    for r, g in zip(ref, pbar):
        r += g.reshape(r.shape)
    loss_prior = float(((J - prior_grad) ** 2).mean())

    def training_step(many):
        net.zero_grad()
        b, tg = y0.cuda(), target.cuda()
        if many:
            predictions = pb.odeint_adjoint_many(net, b, t, method=method)[:, 1]
        else:
            predictions = torch.zeros(b.shape, device="cuda")
            for index, (time, batch_point) in enumerate(zip(t, b)):
                predictions[index, :, :] = pb.odeint_adjoint(net, batch_point, time, method=method)[1]
        ld = torch.mean((predictions - tg) ** 2)
        pred_grad = net.prior_only_forward(t, batch_for_prior.cuda())
        lp = torch.mean((pred_grad - prior_grad.cuda()) ** 2)
        (lam * ld + (1 - lam) * lp).backward()
        return float(ld.detach()), float(lp.detach()), [p.grad.detach().cpu().clone() for p in net.parameters()]

    for many in (False, True):
        ld, lp, grads = training_step(many)
        assert abs(ld - loss_data) <= 1e-5 * abs(loss_data)
        assert abs(lp - loss_prior) <= 1e-5 * abs(loss_prior)
        for i, (g, r) in enumerate(zip(grads, ref)):
            assert rel_l2(g, r) < 5e-5, (method, many, i, rel_l2(g, r))


def test_mse_grad_entry_point_matches_autograd():
    """phx_mse_grad = the cotangent autograd forms for torch.mean((predictions - targets)**2) (train_insilico.py:132), on
    the t1 slices of a [N][T][G] solver output (strided rows), bit for bit the same arithmetic (one sub, one mul)."""
    import ctypes
    from phoenix_b200 import _lib
    lib, ctx = _lib.load(), _lib.ctx(0)
    N, T, G = 5, 2, 1037
    gen = torch.Generator().manual_seed(3)
    yout = torch.rand(N, T, 1, G, generator=gen).cuda()
    target = torch.rand(N, 1, G, generator=gen).cuda()
    pred = yout[:, 1].clone().requires_grad_(True)
    torch.mean((pred - target) ** 2).backward()
    gy = torch.zeros_like(yout)
    sp = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    rc = lib.phx_mse_grad(ctx, N, G, ctypes.c_void_p(yout.data_ptr() + 4 * G), T * G, ctypes.c_void_p(target.data_ptr()),
                          2.0 / (N * G), ctypes.c_void_p(gy.data_ptr() + 4 * G), T * G, sp)
    assert rc == 0
    torch.cuda.synchronize()
    assert float(gy[:, 0].abs().max()) == 0.0
    assert float((gy[:, 1] - pred.grad).abs().max()) <= 1e-7 * float(pred.grad.abs().max())
