"""GPU parity tests of the tcgen05 (tensor-core) batched RHS: calls with B >= 5 rows run the branch and joint
contractions as 3xTF32 MMAs (phx_tc_min_rows() = 5: everything the resident solver kernels do not take) with TMEM accumulators (csrc/phx_tc.cu).

Tolerances:
  * '3xtf32' (default): same bar as the fp32 CUDA-core path, relative L2 <= 1e-5 against the fp32 oracle
    (observed 3e-7 .. 9e-6 up to 20 000 genes; the accumulation chain inside tensor memory is capped at 96 MMAs because
    the tensor core adds with truncation);
  * 'tf32' (single pass, reported separately): relative L2 <= 2e-3 (observed 2e-4 .. 6e-4);
  * 'fp32': the CUDA-core contractions, <= 1e-5.
"""
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
    pb.set_precision("3xtf32")
    pb.set_sync_errors(False)


def make_net(pb, w):
    net = pb.ODENet("cuda", w.G, neurons=w.H)
    with torch.no_grad():
        for p, src in zip(net.parameters(), w.as_list()):
            p.copy_(src)
    return net


SHAPES = [(350, 40, 300), (37, 5, 128), (1001, 100, 257), (690, 40, 1000), (3551, 120, 256), (129, 33, 5),
          (350, 40, 17), (1001, 100, 60)]


@pytest.mark.parametrize("G,H,B", SHAPES)
def test_tc_rhs_matches_oracle(pb, G, H, B):
    """ODENet.forward and prior_only_forward (odenet.py:85-98) on ragged shapes: G, H, B not multiples of any tile."""
    w = O.make_weights(G, H, 300 + G, dense=True, neg_mult_frac=0.2)
    net = make_net(pb, w)
    y = torch.rand(B, 1, G, generator=torch.Generator().manual_seed(B)) * 1.5 - 0.25
    pb.set_precision("3xtf32")
    with torch.no_grad():
        f = net(None, y.cuda())
        J = net.prior_only_forward(None, y.cuda())
    assert rel_l2(f.cpu(), O.rhs(w, y)) < 1e-5
    assert rel_l2(J.cpu(), O.rhs(w, y, decay=False)) < 1e-5


def test_tc_precision_modes(pb):
    G, H, B = 1001, 100, 384
    w = O.make_weights(G, H, 17, dense=True)
    net = make_net(pb, w)
    y = torch.rand(B, G, generator=torch.Generator().manual_seed(3))
    ref = O.rhs(w, y)
    err = {}
    try:
        for mode in ("fp32", "3xtf32", "tf32"):
            pb.set_precision(mode)
            with torch.no_grad():
                err[mode] = rel_l2(net(None, y.cuda()).cpu(), ref)
    finally:
        pb.set_precision("3xtf32")
    assert err["fp32"] < 1e-5 and err["3xtf32"] < 1e-5
    assert 1e-6 < err["tf32"] < 2e-3          # the single-pass mode really is a different arithmetic
    with pytest.raises(ValueError):
        pb.set_precision("bf16")


def test_tc_below_threshold_uses_fp32_path(pb):
    """B < phx_tc_min_rows() = 5 rows never touches the tensor cores: identical bits in every precision mode."""
    G, H, B = 350, 40, 4
    w = O.make_weights(G, H, 18, dense=True)
    net = make_net(pb, w)
    y = torch.rand(B, G, generator=torch.Generator().manual_seed(4)).cuda()
    outs = []
    try:
        for mode in ("fp32", "3xtf32", "tf32"):
            pb.set_precision(mode)
            with torch.no_grad():
                outs.append(net(None, y).clone())
    finally:
        pb.set_precision("3xtf32")
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])


def test_tc_operand_images_follow_weight_updates(pb):
    """The operand images are rebuilt lazily after phx_pack_weights: an in-place optimiser-style update must be seen."""
    G, H, B = 350, 40, 256
    w = O.make_weights(G, H, 19, dense=True)
    net = make_net(pb, w)
    y = torch.rand(B, G, generator=torch.Generator().manual_seed(5))
    with torch.no_grad():
        f0 = net(None, y.cuda()).cpu()
        for p in net.parameters():
            p.mul_(0.5)
        f1 = net(None, y.cuda()).cpu()
    w2 = O.make_weights(G, H, 19, dense=True)
    for t in w2.as_list():
        t.mul_(0.5)
    assert rel_l2(f0, O.rhs(w, y)) < 1e-5
    assert rel_l2(f1, O.rhs(w2, y)) < 1e-5


def test_tc_bitwise_determinism_and_exact_scaling(pb):
    """Fixed MMA issue order and fixed K-split reduction order => run-to-run identical bits; and because the hi/lo TF32
    split commutes with powers of two, doubling Wa doubles prior_only_forward exactly (size-independent property)."""
    G, H, B = 3551, 120, 640
    w = O.make_weights(G, H, 20, dense=True)
    net = make_net(pb, w)
    y = torch.rand(B, G, generator=torch.Generator().manual_seed(6)).cuda()
    with torch.no_grad():
        a = net.prior_only_forward(None, y).clone()
        b = net.prior_only_forward(None, y).clone()
        net.net_alpha_combine.linear_out.weight.mul_(2.0)
        c = net.prior_only_forward(None, y).clone()
    assert torch.equal(a, b)
    assert torch.equal(c, 2.0 * a)


def test_tc_vjp_and_batched_solves(pb):
    """VJP (branch contraction on tensor cores, cotangent contractions fp32) and the streaming solvers on B = 256."""
    G, H, B = 350, 40, 256
    w = O.make_weights(G, H, 21, dense=True, neg_mult_frac=0.1)
    net = make_net(pb, w)
    gen = torch.Generator().manual_seed(7)
    y = torch.rand(B, G, generator=gen)
    g = torch.randn(B, G, generator=gen)
    yg = y.cuda().requires_grad_(True)
    net(None, yg).backward(g.cuda())
    _, ybar, pbar = O.rhs_vjp(w, y, g, decay=True)
    assert rel_l2(yg.grad.cpu(), ybar) < 1e-5
    for p, ref in zip(net.parameters(), pbar):
        assert rel_l2(p.grad.cpu(), ref) < 2e-5
    t = torch.tensor([0.0, 0.5, 1.0])
    with torch.no_grad():
        yr = pb.odeint(net, y.cuda(), t, method="rk4")
        yd = pb.odeint(net, y.cuda(), t, method="dopri5", rtol=1e-5, atol=1e-7)
    y_ref, _ = O.odeint(w, y, t, method="rk4")
    yd_ref, _ = O.odeint(w, y, t, method="dopri5", rtol=1e-5, atol=1e-7)
    assert rel_l2(yr.cpu(), y_ref) < 1e-5
    assert rel_l2(yd.cpu(), yd_ref) < 2e-5


def test_tc_full_size_sweep_shape_against_float64_rows(pb):
    """BASELINE config 5 shape (20 000 genes x 4 096 trajectories, H = 200): the full batch runs on the tensor cores;
    256 of its rows are checked against a float64 evaluation of odenet.py:85-91 on the GPU."""
    G, H, B = 20000, 200, 4096
    torch.manual_seed(22)
    net = pb.ODENet("cuda", G, neurons=H)
    with torch.no_grad():
        for p in net.parameters():
            if p.dim() == 2 and p.shape[0] != 1:
                p.copy_(torch.randn_like(p) * 0.02)
    y = torch.rand(B, G, device="cuda")
    with torch.no_grad():
        f = net(None, y)
        rows = torch.arange(0, B, 16, device="cuda")
        yd = y[rows].double()
        z = yd - 0.5
        s = z / (1 + z.abs())
        l = torch.log1p(s)
        S = s @ net.net_sums.linear_out.weight.double().t() + net.net_sums.linear_out.bias.double()
        P = torch.exp(l @ net.net_prods.linear_out.weight.double().t() + net.net_prods.linear_out.bias.double())
        J = torch.cat([S, P], -1) @ net.net_alpha_combine.linear_out.weight.double().t()
        ref = torch.relu(net.gene_multipliers.double()) * (J - yd)
    assert torch.isfinite(f).all()
    assert rel_l2(f[rows].cpu(), ref.cpu()) < 1e-5


def test_wide_hidden_layer_falls_back_to_fp32_contractions(pb):
    """H > 256 does not fit the 512 tensor-memory columns (chunk accumulator + running sum): CUDA-core path, same bar."""
    G, H, B = 97, 300, 40
    w = O.make_weights(G, H, 23, dense=True)
    net = make_net(pb, w)
    y = torch.rand(B, G, generator=torch.Generator().manual_seed(9))
    with torch.no_grad():
        f = net(None, y.cuda())
    assert rel_l2(f.cpu(), O.rhs(w, y)) < 1e-5


def test_vjp_reuses_forward_activations_only_when_they_are_still_there(pb):
    """engine.rhs_vjp skips the [S|P] recomputation when the workspace still holds the forward of the same (weights, y);
    an intervening forward on other data, or a weight update, must invalidate that."""
    G, H, B = 350, 40, 200
    w = O.make_weights(G, H, 24, dense=True)
    net = make_net(pb, w)
    gen = torch.Generator().manual_seed(10)
    y1, y2 = torch.rand(B, G, generator=gen), torch.rand(B, G, generator=gen)
    g = torch.randn(B, G, generator=gen)
    _, ybar_ref, pbar_ref = O.rhs_vjp(w, y1, g, decay=True)

    def check(yg):
        assert rel_l2(yg.grad.cpu(), ybar_ref) < 1e-5
        for p, ref in zip(net.parameters(), pbar_ref):
            assert rel_l2(p.grad.cpu(), ref) < 2e-5

    # plain forward -> backward (reuse path)
    net.zero_grad()
    a = y1.cuda().requires_grad_(True)
    net(None, a).backward(g.cuda())
    check(a)
    # another forward in between overwrites the workspace: the first graph's backward must recompute
    net.zero_grad()
    a = y1.cuda().requires_grad_(True)
    fa = net(None, a)
    with torch.no_grad():
        net(None, y2.cuda())
    fa.backward(g.cuda())
    check(a)
    # two graphs alive at once, backward in creation order
    net.zero_grad()
    a, b = y1.cuda().requires_grad_(True), y2.cuda().requires_grad_(True)
    fa, fb = net(None, a), net(None, b)
    fa.backward(g.cuda())
    check(a)


def test_cta_pair_variant_is_bit_identical(pb):
    """Experimental cta_group::2 form of the branch-type contractions (off by default): same accumulation order."""
    from phoenix_b200 import _lib
    G, H, B = 1001, 100, 300
    w = O.make_weights(G, H, 25, dense=True)
    net = make_net(pb, w)
    y = torch.rand(B, G, generator=torch.Generator().manual_seed(12)).cuda()
    lib = _lib.load()
    with torch.no_grad():
        f0 = net(None, y).clone()
        lib.phx_tc_set_pair(1)
        try:
            f1 = net(None, y).clone()
        finally:
            lib.phx_tc_set_pair(0)
    assert torch.equal(f0, f1)


@pytest.mark.parametrize("G,H,B,decay", [(129, 33, 5, True), (1001, 100, 257, True), (350, 40, 1000, False),
                                         (37, 5, 130, True)])
def test_tc_vjp_ragged_shapes(pb, G, H, B, decay):
    """State and parameter cotangents on tcgen05 (gSP, u|v passes, the K = batch contractions with transposed loads) at
    shapes where nothing is a multiple of a tile: genes % 128, rows % 16, hidden % 16 all non-zero."""
    w = O.make_weights(G, H, 500 + G, dense=True, neg_mult_frac=0.2)
    net = make_net(pb, w)
    gen = torch.Generator().manual_seed(G + B)
    y = torch.rand(B, G, generator=gen) * 1.4 - 0.2
    g = torch.randn(B, G, generator=gen)
    yg = y.cuda().requires_grad_(True)
    fn = net.forward if decay else net.prior_only_forward
    fn(None, yg).backward(g.cuda())
    _, ybar, pbar = O.rhs_vjp(w, y, g, decay=decay)
    assert rel_l2(yg.grad.cpu(), ybar) < 1e-5
    for i, (p, ref) in enumerate(zip(net.parameters(), pbar)):
        if i == 0 and not decay:
            assert p.grad is None or not p.grad.any()
        else:
            assert rel_l2(p.grad.cpu(), ref) < 2e-5, (i, rel_l2(p.grad.cpu(), ref))
