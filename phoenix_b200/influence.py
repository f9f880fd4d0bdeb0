"""Gene-influence scan of the reference (find_gene_influences.py:60-76) on the GPU path.

For every gene g the reference draws 60 random initial states in [-0.5, 0.5)^G, integrates them over
``t = np.arange(0, 1, 0.1)`` (float64, 10 output times, ONE batched ``odeint`` call: the dopri5 error norm is global over
the 60 rows), overwrites column g of the initial states with fresh random values, integrates again, and scores the gene
by the mean absolute difference of the two trajectory bundles over the nine later output times and all OTHER genes.
Here the two batched solves run on the streaming engine (tcgen05 contractions, 60 rows), the score is reduced on the
device, and the genes of a scan are independent: ``shard`` them over ranks (tools/gene_influence.py)."""
import numpy as np
import torch

from . import parallel
from .torchdiffeq import odeint


def gene_influence_scores(odenet, genes, n_random_inputs_per_gene=60, time_pts=None, method="dopri5", generator=None,
                          inits=None):
    """Scores of ``genes`` (iterable of gene indices).  ``inits`` (optional, for parity tests): a list of
    ``(this_init [n,1,G], this_pert_col [n])`` CPU tensors per gene; otherwise they are drawn on the model's device from
    ``generator`` exactly like find_gene_influences.py:65,67 (``torch.rand(...) - 0.5``)."""
    p0 = odenet.gene_multipliers
    dev, G = p0.device, odenet.ndim
    t = torch.from_numpy(np.arange(0, 1, 0.1)) if time_pts is None else time_pts
    scores = []
    with torch.no_grad():
        for k, g in enumerate(genes):
            if inits is not None:
                this_init, pert_col = inits[k][0].to(dev), inits[k][1].to(dev)
            else:
                this_init = torch.rand(n_random_inputs_per_gene, 1, G, device=dev, generator=generator) - 0.5
                pert_col = torch.rand(n_random_inputs_per_gene, device=dev, generator=generator) - 0.5
            unpert_out = odeint(odenet, this_init, t, method=method)
            this_init = this_init.clone()
            this_init[:, 0, g] = pert_col
            pert_out = odeint(odenet, this_init, t, method=method)
            # mean over times 1.., rows and all genes but g (find_gene_influences.py:71-72)
            d = (unpert_out[1:] - pert_out[1:]).abs()
            total = d.sum(dtype=torch.float64) - d[..., g].sum(dtype=torch.float64)
            scores.append(total / (d[..., 0].numel() * (G - 1)))
    return torch.stack(scores).to(torch.float32) if scores else torch.empty(0)


def shard_genes(n_genes, rank, world_size):
    """Contiguous slice of the gene list for this rank (the scan's only multi-GPU structure: no collective)."""
    lo, hi = parallel.shard_range(n_genes, rank, world_size)
    return range(lo, hi)
