"""Gene-influence scan (find_gene_influences.py:60-76): scores of a few genes at the SIM350 shape against
the CPU oracle run the same way (two batched dopri5 solves of 60 rows over 10 float64 output times, global error norm
over the batch), same random draws.  Tolerance: 5e-3 relative -- with init-distribution weights a score is ~6e-6, the
mean absolute difference of two fp32 trajectories of magnitude ~0.3 whose last bit is 3e-8: fp32 resolution alone is
~3e-3 of the score (observed 1.4e-3)."""
import numpy as np
import pytest
import torch

from oracle import phoenix_oracle as O

pytestmark = pytest.mark.gpu


def test_influence_scores_match_the_oracle():
    import phoenix_b200 as pb
    from phoenix_b200 import influence
    G, H, n = 350, 40, 60
    w = O.make_weights(G, H, 8001, dense=False)
    net = pb.ODENet("cuda", G, neurons=H)
    with torch.no_grad():
        for p, src in zip(net.parameters(), w.as_list()):
            p.copy_(src)
    genes = [0, 17, 349]
    gen = torch.Generator().manual_seed(8002)
    inits = [(torch.rand(n, 1, G, generator=gen) - 0.5, torch.rand(n, generator=gen) - 0.5) for _ in genes]
    t = torch.from_numpy(np.arange(0, 1, 0.1))
    ref = []
    for g, (init, col) in zip(genes, inits):
        a, _ = O.odeint(w, init, t, method="dopri5")
        pert = init.clone()
        pert[:, 0, g] = col
        b, _ = O.odeint(w, pert, t, method="dopri5")
        others = [i for i in range(G) if i != g]
        ref.append(float(torch.mean(abs(a[1:, :, :, others] - b[1:, :, :, others]))))
    mine = influence.gene_influence_scores(net, genes, inits=inits).cpu().tolist()
    for m, r in zip(mine, ref):
        assert abs(m - r) <= 5e-3 * abs(r), (mine, ref)
    assert list(influence.shard_genes(10, 1, 3)) == [4, 5, 6]
    # genes scanned side by side on several streams / host threads: the very same scores as the sequential scan
    seq = influence.gene_influence_scores(net, genes, inits=inits, workers=1)
    par = influence.gene_influence_scores(net, genes, inits=inits, workers=3)
    assert torch.equal(seq, par) and torch.equal(seq.cpu(), torch.tensor(mine))
    g1, g2 = (torch.Generator(device="cuda").manual_seed(5) for _ in range(2))
    a = influence.gene_influence_scores(net, range(6), generator=g1, workers=1)
    b = influence.gene_influence_scores(net, range(6), generator=g2, workers=4)
    assert torch.equal(a, b)
