"""CPU, world_size 2, gloo: the multi-GPU host logic (shard the samples, one sum-allreduce of the flat gradient)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import phoenix_b200 as pb
from phoenix_b200 import parallel


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(1234 + rank)                      # different weights per rank before the broadcast
    net = pb.ODENet("cpu", 12, neurons=4)
    parallel.broadcast_parameters(net, src=0)
    lo, hi = parallel.shard_range(5, rank, world)       # 5 samples over 2 ranks -> 3 + 2
    for i, p in enumerate(net.parameters()):            # fake per-rank gradients: (i+1) * number of local samples
        p.grad = torch.full_like(p, float((i + 1) * (hi - lo)))
    loss = torch.tensor([float(hi - lo)])
    total = parallel.allreduce_grads(net, extra=loss, average=False)
    ok = all(torch.allclose(p.grad, torch.full_like(p, float((i + 1) * 5))) for i, p in enumerate(net.parameters()))
    # default: the average over the ranks (the reference's losses are means over the LOCAL samples)
    for i, p in enumerate(net.parameters()):
        p.grad = torch.full_like(p, float((i + 1) * (rank + 1)))
    parallel.allreduce_grads(net)
    ok = ok and all(torch.allclose(p.grad, torch.full_like(p, float((i + 1) * 1.5))) for i, p in enumerate(net.parameters()))
    # after the gathered path .grad are views of ONE flat buffer: the next collective runs in place on it
    flat = parallel.flat_grad_view(net)
    ok = ok and flat is not None and flat.numel() == sum(p.numel() for p in net.parameters())
    ptr = flat.data_ptr()
    flat.fill_(float(rank + 1))
    extra = parallel.allreduce_grads(net, average=False, extra=torch.tensor([2.0, float(rank)]))
    ok = ok and parallel.flat_grad_view(net).data_ptr() == ptr and bool((flat == 3.0).all())
    ok = ok and all(bool((p.grad == 3.0).all()) for p in net.parameters()) and extra.tolist() == [4.0, 1.0]
    net.gene_multipliers.grad = torch.zeros_like(net.gene_multipliers)      # no longer one buffer
    ok = ok and parallel.flat_grad_view(net) is None
    w0 = net.net_sums.linear_out.weight.detach().clone()
    gathered = [torch.zeros_like(w0) for _ in range(world)]
    dist.all_gather(gathered, w0)
    same = all(torch.equal(g, gathered[0]) for g in gathered)
    if rank == 0:
        with open(out, "w") as fh:
            fh.write("%d %d %.1f" % (int(ok), int(same), float(total)))
    dist.destroy_process_group()


def test_allreduce_and_broadcast_world2(tmp_path):
    out = os.path.join(str(tmp_path), "res.txt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    ok, same, total = open(out).read().split()
    assert ok == "1" and same == "1" and float(total) == 5.0
