"""BASELINE config 5 (synthetic scaling sweep): 20 000 genes x 4 096 trajectories in ONE batched odeint call,
rk4 and dopri5, forward solve and forward + adjoint, on one B200 (streaming engine + tcgen05 contractions).

    python tools/sweep_c5.py [--genes 20000] [--neurons 200] [--rows 4096] [--t1 0.1] [--cpu-rows 8]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/sweep_c5.py ...

Under torchrun the rows are sharded over the ranks (weak scaling: --rows is PER RANK), weights are broadcast from rank
0, the adjoint leg ends with ONE NCCL sum-allreduce of the flat gradient, times are the max over ranks.

Prints one JSON line per (method, leg): gene-steps/s = B * G * RHS evaluations / device time.  With --cpu-rows N the
oracle (torch CPU port of the reference) is timed on N of the rows for the "vs host CPU" column.
"""
import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import phoenix_b200 as pb  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--genes", type=int, default=20000)
    ap.add_argument("--neurons", type=int, default=200)
    ap.add_argument("--rows", type=int, default=4096)
    ap.add_argument("--cpu-rows", type=int, default=0)
    ap.add_argument("--strong", action="store_true", help="--rows is the TOTAL batch, split evenly over the ranks")
    ap.add_argument("--global-norm", action="store_true",
                    help="dopri5 forward: all-reduce the error-norm sums over the ranks every step (one controller for the "
                         "whole batch, exactly like the unsharded call)")
    ap.add_argument("--t1", type=float, default=0.1, help="the call integrates t = [0, T1] (SURVEY 8d grid: 0.01, 0.1, 1)")
    ap.add_argument("--rtol", type=float, default=1e-7)
    ap.add_argument("--atol", type=float, default=1e-9)
    a = ap.parse_args()
    G, H, B = a.genes, a.neurons, a.rows
    import torch.distributed as dist
    from phoenix_b200 import parallel
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    torch.manual_seed(5)
    net = pb.ODENet("cuda", G, neurons=H)
    parallel.broadcast_parameters(net)
    if a.strong:
        assert B % world == 0, "--strong needs a batch divisible by the number of ranks"
        B = B // world
    torch.manual_seed(100 + rank)
    y0 = torch.rand(B, G, device="cuda")
    pb.set_sync_errors(True)
    if a.global_norm and world > 1:
        parallel.enable_global_norm()
    for method, t, kw in (("rk4", torch.tensor([0.0, a.t1]), {}),
                          ("dopri5", torch.tensor([0.0, a.t1]), {"rtol": a.rtol, "atol": a.atol})):
        for leg in ("forward", "forward+adjoint"):
            def run():
                if leg == "forward":
                    with torch.no_grad():
                        pb.odeint(net, y0, t, method=method, **kw)
                    n = pb.last_status()["n_rhs"]
                    return n
                for p in net.parameters():
                    p.grad = None
                yg = y0.detach().requires_grad_(True)
                y = pb.odeint_adjoint(net, yg, t, method=method, **kw)
                nf = pb.last_status()["n_rhs"]
                (y[1] ** 2).mean().backward()
                nb = pb.last_status()["n_rhs"]
                if world > 1:
                    parallel.allreduce_grads(net, average=False)
                return nf + nb
            run()
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            evals = run()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            work = float(B * G * evals)
            if world > 1:
                st = torch.tensor([ms, work], dtype=torch.float64, device="cuda")
                mx = st.clone()
                dist.all_reduce(mx, op=dist.ReduceOp.MAX)
                dist.all_reduce(st, op=dist.ReduceOp.SUM)
                ms, work = float(mx[0]), float(st[1])
            if rank == 0:
                print(json.dumps({"n_gpus": world, "G": G, "H": H, "rows_per_gpu": B, "method": method, "leg": leg,
                                  "scaling": "strong" if a.strong else "weak", "rtol": kw.get("rtol"), "t1": a.t1,
                                  "global_norm": bool(a.global_norm and world > 1 and method == "dopri5"),
                                  "rhs_evals": evals, "ms": ms, "gene_steps_per_s": work / (ms * 1e-3)}), flush=True)
    if a.cpu_rows and rank == 0 and world == 1:
        from oracle import phoenix_oracle as O
        torch.set_num_threads(os.cpu_count() or 1)
        w = O.Weights(*[p.detach().cpu() for p in net.parameters()]) if hasattr(O, "Weights") else None
        yc = y0[:a.cpu_rows].cpu()
        O.odeint(w, yc, torch.tensor([0.0, a.t1]), method="rk4")
        t0 = time.perf_counter()
        _, log = O.odeint(w, yc, torch.tensor([0.0, a.t1]), method="rk4")
        dt = time.perf_counter() - t0
        print(json.dumps({"cpu_port": True, "rows": a.cpu_rows, "method": "rk4", "leg": "forward", "s": dt,
                          "gene_steps_per_s": a.cpu_rows * G * 4 / dt, "cores": os.cpu_count()}), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
