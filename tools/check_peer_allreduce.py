"""2 / 4 / 8 GPUs (torchrun): the peer-memory gradient all-reduce (phx_peer_allreduce, csrc/phx_peer.cu) against an NCCL
all-reduce of the same vectors -- values, bit-identity across the ranks, run-to-run determinism -- and the time of both on
the 35.8 MB flat gradient of the breast model.  Prints PASS / FAIL on rank 0.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/check_peer_allreduce.py"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import phoenix_b200 as pb  # noqa: E402,F401
from phoenix_b200 import parallel  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dev = torch.device("cuda", torch.cuda.current_device())
    dist.init_process_group("nccl", device_id=dev)
    G, H = 11165, 200
    P = 4 * G * H + 2 * H + G          # 8 943 565 floats: not a multiple of 4 -> exercises the tail path
    pr = parallel.PeerAllReduce(P, device=dev, nvls=False)
    has_mc = pr._mc != 0
    ok = True
    outs = []
    for rep in range(5 if has_mc else 3):
        pr.nvls = rep >= 3          # the last two rounds: sums formed inside the NVSwitch (multimem)
        torch.manual_seed(100 * rep + rank)
        x = torch.randn(P, device=dev)
        ref = x.clone()
        dist.all_reduce(ref)
        pr.buffer[:P].copy_(x)
        out = pr.reduce(scale=0.5 if rep == 2 else 1.0).clone()
        if rep == 2:
            ref *= 0.5
        err = float((out - ref).abs().max() / ref.abs().max())
        ok = ok and err < 1e-6
        gathered = [torch.empty_like(out) for _ in range(world)]
        dist.all_gather(gathered, out)
        ok = ok and all(torch.equal(g, gathered[0]) for g in gathered)      # bit-identical on every rank
        outs.append((x, out))
    pr.nvls = False
    x, out = outs[0]
    pr.buffer[:P].copy_(x)
    ok = ok and torch.equal(pr.reduce().clone(), out)                        # deterministic
    if has_mc:
        pr.nvls = True
        x, out = outs[3]
        pr.buffer[:P].copy_(x)
        ok = ok and torch.equal(pr.reduce().clone(), out)
        pr.nvls = False
    # timing: device time of one collective, max over ranks
    def timed(fn, n=20):
        for _ in range(3):
            fn()
        dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n):
            fn()
        b.record()
        torch.cuda.synchronize()
        t = torch.tensor([a.elapsed_time(b) / n], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)
    y = torch.randn(P, device=dev)
    t_nccl = timed(lambda: dist.all_reduce(y))
    t_peer = timed(lambda: pr.reduce())
    t_nvls = float("nan")
    if has_mc:
        pr.nvls = True
        t_nvls = timed(lambda: pr.reduce())
    flag = torch.tensor([int(ok)], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("peer all-reduce, %d ranks, %d floats (%.1f MB): max rel. error vs NCCL < 1e-6, bit-identical across ranks, "
              "deterministic: %s | device time per call: peer kernel %.1f us, NVLS (multimem) kernel %.1f us, NCCL all_reduce "
              "%.1f us  ->  %s" % (world, P, P * 4 / 1e6, bool(int(flag)), 1e3 * t_peer, 1e3 * t_nvls, 1e3 * t_nccl,
                                   "PASS" if int(flag) else "FAIL"),
              flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
