"""2+ GPUs (torchrun): a batched dopri5 forward solve with its rows sharded over the ranks and the exact-global-norm mode
(parallel.enable_global_norm: one all-reduce of the error-norm sums per step) must take the SAME attempted steps as the
unsharded call and produce the same rows.  Prints PASS / FAIL on rank 0.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/check_global_norm.py"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import phoenix_b200 as pb  # noqa: E402
from phoenix_b200 import parallel  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    G, H, B = 3551, 120, 64 * world
    torch.manual_seed(3)
    net = pb.ODENet("cuda", G, neurons=H)
    parallel.broadcast_parameters(net)
    y0 = torch.rand(B, G, generator=torch.Generator().manual_seed(4)).cuda()
    t = torch.tensor([0.0, 0.4, 1.0], dtype=torch.float64)
    pb.set_step_logging(True)
    lo, hi = parallel.shard_range(B, rank, world)
    with torch.no_grad():
        full = pb.odeint(net, y0, t, method="dopri5")            # every rank: the unsharded call
        log_full = pb.last_step_log()
        mine_local = pb.odeint(net, y0[lo:hi], t, method="dopri5")   # per-shard controller
        log_local = pb.last_step_log()
        parallel.enable_global_norm()
        mine = pb.odeint(net, y0[lo:hi], t, method="dopri5")
        log_glob = pb.last_step_log()
        parallel.disable_global_norm()
    same_steps = len(log_glob) == len(log_full) and all(
        a[2] == b[2] and abs(a[1] - b[1]) <= 1e-6 * abs(b[1]) for a, b in zip(log_glob, log_full))
    err = float((mine - full[:, lo:hi]).norm() / full[:, lo:hi].norm())
    err_local = float((mine_local - full[:, lo:hi]).norm() / full[:, lo:hi].norm())
    ok = torch.tensor([int(same_steps and err < 2e-6)], device="cuda")
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("global-norm mode: steps %d (unsharded %d, per-shard %d), identical sequence: %s, rel. error of the rows "
              "%.2e (per-shard controllers: %.2e, first dt differs: %s)  ->  %s" % (
                  len(log_glob), len(log_full), len(log_local), same_steps, err, err_local,
                  log_local[1][1] != log_full[1][1] if len(log_local) > 1 and len(log_full) > 1 else None,
                  "PASS" if int(ok) else "FAIL"), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
