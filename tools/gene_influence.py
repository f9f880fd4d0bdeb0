"""Gene-influence scan (find_gene_influences.py:60-76) with the genes sharded over the ranks.

    python tools/gene_influence.py [--genes 11165 --neurons 200 --count 64] [--out scores.csv]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/gene_influence.py ...

--count limits the scan to the first COUNT genes of every rank's shard (throughput measurement); without it the whole
gene list is scanned.  Prints one JSON line: genes/s over all ranks (max-over-ranks time), the projected time of a full
scan, and the first scores.  Random-init weights (there is no trained checkpoint on the box)."""
import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import phoenix_b200 as pb  # noqa: E402
from phoenix_b200 import influence, parallel  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--genes", type=int, default=11165)
    ap.add_argument("--neurons", type=int, default=200)
    ap.add_argument("--count", type=int, default=0)
    ap.add_argument("--method", default="dopri5")
    ap.add_argument("--out", default="")
    ap.add_argument("--workers", type=int, default=4, help="genes scanned side by side (host threads / CUDA streams)")
    a = ap.parse_args()
    import torch.distributed as dist
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    dev = torch.device("cuda", torch.cuda.current_device())
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(0)
    odenet = pb.ODENet(dev, a.genes, neurons=a.neurons)
    parallel.broadcast_parameters(odenet)
    genes = list(influence.shard_genes(a.genes, rank, world))
    if a.count:
        genes = genes[:a.count]
    gen = torch.Generator(device=dev).manual_seed(100 + rank)
    influence.gene_influence_scores(odenet, genes[:2 * a.workers], method=a.method, generator=gen,
                                    workers=a.workers)     # warm-up (every worker stream gets its workspaces)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    scores = influence.gene_influence_scores(odenet, genes, method=a.method, generator=gen, workers=a.workers)
    torch.cuda.synchronize()
    secs = time.perf_counter() - t0
    stats = torch.tensor([secs, float(len(genes))], dtype=torch.float64, device=dev)
    if world > 1:
        mx = stats.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = stats.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        secs, n_done = float(mx[0]), float(sm[1])
        gathered = [torch.zeros_like(scores) for _ in range(world)] if len(set([len(genes)])) == 1 else None
        if gathered is not None:
            dist.all_gather(gathered, scores)
            scores = torch.cat(gathered)
    else:
        n_done = float(len(genes))
    if rank == 0:
        rate = n_done / secs
        print(json.dumps({"metric": "gene-influence scan", "genes": a.genes, "neurons": a.neurons, "n_gpus": world,
                          "method": a.method, "workers": a.workers, "rows_per_solve": 60, "output_times": 10, "genes_scanned": int(n_done),
                          "seconds": secs, "genes_per_s": rate, "solves_per_s": 2 * rate,
                          "full_scan_projected_s": a.genes / rate,
                          "scores_head": [float(x) for x in scores[:4].tolist()]}), flush=True)
        if a.out:
            import numpy as np
            np.savetxt(a.out, scores.cpu().numpy(), delimiter=",")
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
