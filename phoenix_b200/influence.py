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


_worker_streams = {}


def _score_one(odenet, g, this_init, pert_col, t, method, G):
    unpert_out = odeint(odenet, this_init, t, method=method)
    this_init = this_init.clone()
    this_init[:, 0, g] = pert_col
    pert_out = odeint(odenet, this_init, t, method=method)
    # mean over times 1.., rows and all genes but g (find_gene_influences.py:71-72)
    d = (unpert_out[1:] - pert_out[1:]).abs()
    total = d.sum(dtype=torch.float64) - d[..., g].sum(dtype=torch.float64)
    return total / (d[..., 0].numel() * (G - 1))


def gene_influence_scores(odenet, genes, n_random_inputs_per_gene=60, time_pts=None, method="dopri5", generator=None,
                          inits=None, workers=4):
    """Scores of ``genes`` (iterable of gene indices).  ``inits`` (optional, for parity tests): a list of
    ``(this_init [n,1,G], this_pert_col [n])`` CPU tensors per gene; otherwise they are drawn on the model's device from
    ``generator`` exactly like find_gene_influences.py:65,67 (``torch.rand(...) - 0.5``), gene by gene in scan order.

    A 60-row solve keeps about half of the GPU busy and its adaptive step controller runs on the host (one small read-back
    per attempted step), so ``workers`` genes are scanned side by side: one host thread and one CUDA stream per worker (the
    solver's host loop runs outside the GIL).  Every gene's two solves are still their own ``odeint`` calls with their own
    step sequences: the scores do not depend on ``workers``."""
    import concurrent.futures as cf
    p0 = odenet.gene_multipliers
    dev, G = p0.device, odenet.ndim
    t = torch.from_numpy(np.arange(0, 1, 0.1)) if time_pts is None else time_pts
    genes = list(genes)
    if not genes:
        return torch.empty(0)
    workers = max(1, min(int(workers), len(genes)))
    main = torch.cuda.current_stream(dev)
    if workers > 1:   # the worker streams are kept: the solver workspaces are cached per stream (engine._workspace)
        key = (dev.index if dev.index is not None else torch.cuda.current_device(), workers)
        streams = _worker_streams.get(key)
        if streams is None:
            streams = _worker_streams[key] = [torch.cuda.Stream(device=dev) for _ in range(workers)]
    else:
        streams = [main]

    def draw(k):
        if inits is not None:
            return inits[k][0].to(dev), inits[k][1].to(dev)
        return (torch.rand(n_random_inputs_per_gene, 1, G, device=dev, generator=generator) - 0.5,
                torch.rand(n_random_inputs_per_gene, device=dev, generator=generator) - 0.5)

    def run(k, g, this_init, pert_col, ready):
        s = streams[k % workers]
        with torch.cuda.device(dev), torch.cuda.stream(s), torch.no_grad():
            s.wait_event(ready)
            out = _score_one(odenet, g, this_init, pert_col, t, method, G)
            this_init.record_stream(s)
            pert_col.record_stream(s)
            return out

    scores = [None] * len(genes)
    with torch.no_grad():
        if workers == 1:
            for k, g in enumerate(genes):
                this_init, pert_col = draw(k)
                scores[k] = _score_one(odenet, g, this_init, pert_col, t, method, G)
        else:
            # the draws stay on the caller's stream, in scan order (same random numbers as the sequential scan); a worker
            # is handed a gene only when its stream's previous gene has been submitted, so each stream sees its genes
            # in order and at most `workers` bundles of trajectories are alive
            with cf.ThreadPoolExecutor(max_workers=workers) as ex:
                pending = [None] * workers
                for k, g in enumerate(genes):
                    w = k % workers
                    if pending[w] is not None:
                        kk, fut = pending[w]
                        scores[kk] = fut.result()
                    this_init, pert_col = draw(k)
                    ready = torch.cuda.Event()
                    ready.record(main)
                    pending[w] = (k, ex.submit(run, k, g, this_init, pert_col, ready))
                for item in pending:
                    if item is not None:
                        scores[item[0]] = item[1].result()
            for s in streams:
                main.wait_stream(s)
    return torch.stack(scores).to(torch.float32)


def shard_genes(n_genes, rank, world_size):
    """Contiguous slice of the gene list for this rank (the scan's only multi-GPU structure: no collective)."""
    lo, hi = parallel.shard_range(n_genes, rank, world_size)
    return range(lo, hi)
