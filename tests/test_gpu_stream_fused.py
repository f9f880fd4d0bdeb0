"""Streaming engine, forward solves on the tensor-core path: the RK stage algebra that follows a stage derivative is
evaluated in the epilogue of the joint contraction producing it (PhxRhsPost, csrc/phx_common.cuh; csrc/phx_stream.cu
`fuse`; opt-in with PHX_STREAM_FUSE=1 -- measured slower than the stand-alone kernels).  Same formulas
in the same operation order as the stand-alone elementwise kernels (the default), so the two must agree BIT FOR BIT -- trajectories at every output time and, for dopri5, the step sequence -- and both are held to
the oracle (fixed_grid.py:6-38, rk_common.py:39-77, interp.py)."""
import os

import pytest
import torch

from golden_util import rel_l2
from oracle import phoenix_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pb():
    import phoenix_b200 as pb
    pb.set_sync_errors(True)
    yield pb
    pb.engine.FORCE_ENGINE = None
    os.environ.pop("PHX_STREAM_FUSE", None)
    pb.set_step_logging(False)
    pb.set_sync_errors(False)


def make_net(pb, w):
    net = pb.ODENet("cuda", w.G, neurons=w.H)
    with torch.no_grad():
        for p, src in zip(net.parameters(), w.as_list()):
            p.copy_(src)
    return net


def _solve(pb, net, y0, t, method, fuse, **kw):
    os.environ["PHX_STREAM_FUSE"] = "1" if fuse else "0"
    pb.engine.FORCE_ENGINE = "stream"
    pb.set_step_logging(method == "dopri5")
    try:
        with torch.no_grad():
            y = pb.odeint(net, y0, t, method=method, **kw)
        log = pb.last_step_log() if method == "dopri5" else None
        st = pb.last_status()
    finally:
        pb.engine.FORCE_ENGINE = None
        pb.set_step_logging(False)
        os.environ.pop("PHX_STREAM_FUSE", None)
    return y, log, st


# ragged shapes: G, H, B multiples of no tile; 300 rows = two batch tiles, 1037 genes = nine gene tiles
CASES = [("euler", 129, 33, 5, 3), ("midpoint", 350, 40, 17, 3), ("rk4", 1037, 56, 300, 4), ("rk4", 37, 5, 128, 2),
         ("dopri5", 350, 40, 60, 4), ("dopri5", 1037, 56, 300, 2)]


@pytest.mark.parametrize("method,G,H,B,T", CASES, ids=["%s_g%d_b%d" % (c[0], c[1], c[3]) for c in CASES])
def test_fused_stage_algebra_is_bit_identical_and_matches_the_oracle(pb, method, G, H, B, T):
    w = O.make_weights(G, H, 900 + G + B, dense=True, neg_mult_frac=0.1)
    net = make_net(pb, w)
    gen = torch.Generator().manual_seed(G + B)
    y0 = torch.rand(B, 1, G, generator=gen)
    t = torch.linspace(0.0, 1.5, T)
    kw = {"rtol": 1e-5, "atol": 1e-7} if method == "dopri5" else {}
    yf, logf, stf = _solve(pb, net, y0.cuda(), t, method, True, **kw)
    yu, logu, stu = _solve(pb, net, y0.cuda(), t, method, False, **kw)
    assert torch.equal(yf, yu)
    assert stf == stu and logf == logu
    if method == "dopri5":
        assert stf["n_accepted"] >= 2          # more than one attempt: FSAL slot swap and several dense outputs
    y_ref, _ = O.odeint(w, y0, t, method=method, **kw)
    assert rel_l2(yf.cpu(), y_ref) < (5e-5 if method == "dopri5" else 1e-5)   # dopri5 at rtol 1e-5: controller noise
    assert torch.equal(yf[0].cpu(), y0)


def test_fused_forward_under_odeint_adjoint(pb):
    """The forward half of odeint_adjoint on the streaming engine is fused as well; the backward sweep reads its saved
    states: gradients unchanged to the bit."""
    G, H, B = 350, 40, 24
    w = O.make_weights(G, H, 77, dense=True)
    net = make_net(pb, w)
    gen = torch.Generator().manual_seed(5)
    y0 = torch.rand(B, 1, G, generator=gen).cuda()
    t = torch.tensor([0.0, 0.7, 1.1])
    out = []
    for fuse in (True, False):
        os.environ["PHX_STREAM_FUSE"] = "1" if fuse else "0"
        pb.engine.FORCE_ENGINE = "stream"
        try:
            net.zero_grad()
            yg = y0.clone().requires_grad_(True)
            y = pb.odeint_adjoint(net, yg, t, method="rk4")
            (y[1:] ** 2).mean().backward()
            out.append((y.detach().clone(), yg.grad.clone(), [p.grad.clone() for p in net.parameters()]))
        finally:
            pb.engine.FORCE_ENGINE = None
            os.environ.pop("PHX_STREAM_FUSE", None)
    assert torch.equal(out[0][0], out[1][0]) and torch.equal(out[0][1], out[1][1])
    for a, b in zip(out[0][2], out[1][2]):
        assert torch.equal(a, b)
