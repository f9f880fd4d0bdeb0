"""The prior-constrained loss path (train_insilico.py:64-68, 134-138, 208-209) against a golden produced by
the unmodified reference on the SHIPPED prior matrix edge_prior_matrix_G690_noise_0.0.csv (tests/golden/
make_golden_prior.py).  CPU: the oracle restatement; GPU: the sparse set-up product and the fused loss + backward.
Tolerance: relative L2 <= 1e-5 (fp32; 3xTF32 contractions)."""
import numpy as np
import pytest
import torch

from golden_util import load, rel_l2
from oracle import phoenix_oracle as O


def _golden():
    d = load("prior_g690_h40_k256")
    G, H = int(d["G"]), int(d["H"])
    w = O.make_weights(G, H, int(d["seed"]), dense=False)
    prior = torch.zeros(G, G)
    prior[torch.from_numpy(d["prior_rows"]).long(), torch.from_numpy(d["prior_cols"]).long()] = torch.from_numpy(d["prior_vals"])
    return d, w, prior


def test_oracle_prior_loss_matches_reference():
    d, w, prior = _golden()
    x = torch.from_numpy(d["batch_for_prior"])
    pg = torch.matmul(x, prior)
    assert torch.equal(pg, torch.from_numpy(d["prior_grad"]))      # same ATen matmul as the reference
    J = O.rhs(w, x, decay=False)
    loss = torch.mean((J - pg) ** 2)
    assert abs(float(loss) - float(d["loss_prior"])) <= 1e-6 * float(d["loss_prior"])
    _, _, pbar = O.rhs_vjp(w, x, 2.0 * (J - pg) / J.numel(), decay=False)
    for i, g in enumerate(pbar):
        ref = d["grad%d" % i]
        if not ref.any():
            assert not g.any()
        else:
            assert rel_l2(g, ref) < 2e-6, (i, rel_l2(g, ref))


@pytest.mark.gpu
def test_prior_setup_and_fused_loss_match_reference():
    import phoenix_b200 as pb
    d, w, prior = _golden()
    net = pb.ODENet("cuda", w.G, neurons=w.H)
    with torch.no_grad():
        for p, src in zip(net.parameters(), w.as_list()):
            p.copy_(src)
    x = torch.from_numpy(d["batch_for_prior"]).cuda()
    # set-up product: dense and sparse inputs, CPU or CUDA prior matrix
    for pm in (prior, prior.cuda(), prior.to_sparse()):
        pg = pb.prior_grad_from_matrix(x, pm)
        assert pg.shape == x.shape and rel_l2(pg.cpu(), d["prior_grad"]) < 1e-6
    # fused loss + backward
    net.zero_grad()
    loss = pb.prior_loss(net, x, pg)
    (0.01 * loss).backward()          # the (1 - lambda) weight of composed_loss, train_insilico.py:137
    assert abs(float(loss) - float(d["loss_prior"])) <= 1e-5 * float(d["loss_prior"])
    for i, p in enumerate(net.parameters()):
        ref = d["grad%d" % i]
        if not ref.any():
            assert p.grad is None or not p.grad.any()
        else:
            assert rel_l2(p.grad.cpu(), 0.01 * ref) < 1e-5, (i, rel_l2(p.grad.cpu(), 0.01 * ref))
    # the reference's two unfused lines on the same inputs give the same thing
    net.zero_grad()
    loss2 = torch.mean((net.prior_only_forward(None, x) - pg) ** 2)
    (0.01 * loss2).backward()
    assert abs(float(loss2) - float(loss)) <= 1e-5 * float(loss)
    for i, p in enumerate(net.parameters()):
        if d["grad%d" % i].any():
            assert rel_l2(p.grad.cpu(), 0.01 * d["grad%d" % i]) < 1e-5


@pytest.mark.gpu
def test_fused_prior_loss_small_batch_and_large_shape():
    """Below the tensor-core threshold (B = 3: fp32 path) and at the breast shape with 1 024 rows."""
    import phoenix_b200 as pb
    for G, H, K in ((97, 12, 3), (11165, 200, 1024)):
        w = O.make_weights(G, H, 6100 + K, dense=False)
        net = pb.ODENet("cuda", G, neurons=H)
        with torch.no_grad():
            for p, src in zip(net.parameters(), w.as_list()):
                p.copy_(src)
        gen = torch.Generator().manual_seed(K)
        x = torch.rand(K, 1, G, generator=gen) - 0.5
        pg = torch.randn(K, 1, G, generator=gen) * 0.1
        loss = pb.prior_loss(net, x.cuda(), pg.cuda())
        loss.backward()
        J = O.rhs(w, x, decay=False)
        _, _, pbar = O.rhs_vjp(w, x, 2.0 * (J - pg) / J.numel(), decay=False)
        assert abs(float(loss) - float(torch.mean((J - pg) ** 2))) <= 1e-5 * float(loss)
        for i, (p, ref) in enumerate(zip(net.parameters(), pbar)):
            if i == 0:
                assert p.grad is None or not p.grad.any()
            else:
                assert rel_l2(p.grad.cpu(), ref) < 2e-5, (G, i, rel_l2(p.grad.cpu(), ref))


@pytest.mark.gpu
def test_cached_activations_of_the_prior_batch_change_nothing():
    """prior_loss caches soft-sign / log1p of the constant prior batch (phx_hill_planes) and feeds the contractions from
    the planes: same loss and gradients as the uncached path (rounding level: the K ranges of the two launches differ),
    recomputed when the batch is modified in place."""
    import phoenix_b200 as pb
    from phoenix_b200 import engine as prior      # the cache lives in the engine
    torch.manual_seed(3)
    G, H, K = 1037, 56, 700
    net = pb.ODENet("cuda", G, neurons=H)
    x = (torch.rand(K, 1, G, device="cuda") - 0.5) * 3.0       # beyond the series range of log1p as well
    pg = torch.randn(K, 1, G, device="cuda") * 0.1

    def run():
        net.zero_grad()
        loss = pb.prior_loss(net, x, pg)
        loss.backward()
        return float(loss), [p.grad.clone() for p in net.parameters() if p.grad is not None]

    prior.CACHE_ACTIVATIONS = False
    l0, g0 = run()
    prior.CACHE_ACTIVATIONS = True
    try:
        l1, g1 = run()
        l2, g2 = run()                       # second call: planes come from the cache
        assert len(prior._hill_planes) == 1
        assert abs(l1 - l0) <= 1e-6 * abs(l0) and l2 == l1
        for a, b, c in zip(g0, g1, g2):
            assert float((a - b).norm() / a.norm()) < 2e-6 and torch.equal(b, c)
        x.mul_(0.5)                           # in-place edit: the planes are recomputed
        l3, _ = run()
        prior.CACHE_ACTIVATIONS = False
        l4, _ = run()
        assert abs(l3 - l4) <= 1e-6 * abs(l4) and abs(l3 - l1) > 1e-3 * abs(l1)
        # the reference's own two lines (train_insilico.py:134-135): prior_only_forward caches from the second sighting on
        prior.CACHE_ACTIVATIONS = True
        prior._hill_planes.clear()
        prior._hill_seen.clear()
        outs = []
        for _ in range(3):
            net.zero_grad()
            lp = torch.mean((net.prior_only_forward(None, x) - pg) ** 2)
            lp.backward()
            outs.append((float(lp), net.net_sums.linear_out.weight.grad.clone()))
        assert len(prior._hill_planes) == 1
        assert abs(outs[1][0] - outs[0][0]) <= 1e-6 * abs(outs[0][0]) and outs[2][0] == outs[1][0]
        assert float((outs[1][1] - outs[0][1]).norm() / outs[0][1].norm()) < 2e-6 and torch.equal(outs[1][1], outs[2][1])
        assert abs(outs[0][0] - l3) <= 1e-5 * abs(l3)
        # ODENet.forward (the ODE right-hand side, a new state every call) never caches
        y = torch.rand(K, G, device="cuda")
        for _ in range(3):
            net(None, y)
        assert len(prior._hill_planes) == 1
    finally:
        prior.CACHE_ACTIVATIONS = True


@pytest.mark.gpu
@pytest.mark.parametrize("G,B", [(129, 7), (690, 1), (11165, 10), (20000, 5), (26000, 3)],
                         ids=["g129_b7", "g690_b1", "g11165_b10_4rows", "g20000_b5_2rows", "g26000_b3_1row"])
def test_prior_setup_ragged_rows(G, B):
    """phx_prior_setup puts 4 / 2 / 1 batch rows side by side in a CTA's shared memory depending on G; the last CTA of
    a batch that is not a multiple of that holds fewer.  Against the float64 product (train_insilico.py:209)."""
    import phoenix_b200 as pb
    gen = torch.Generator().manual_seed(G + B)
    nnz = 20 * G
    rows = torch.randint(0, G, (nnz,), generator=gen)
    cols = torch.randint(0, G, (nnz,), generator=gen)
    vals = torch.rand(nnz, generator=gen) - 0.5
    prior = torch.sparse_coo_tensor(torch.stack([rows, cols]), vals, (G, G)).coalesce()
    x = torch.rand(B, 1, G, generator=gen) - 0.5
    pg = pb.prior_grad_from_matrix(x.cuda(), prior)
    ref = torch.sparse.mm(prior.double().t(), x.view(B, G).double().t()).t().reshape(B, 1, G)
    assert pg.shape == x.shape and rel_l2(pg.cpu(), ref) < 1e-6
