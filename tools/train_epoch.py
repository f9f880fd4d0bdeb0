"""BASELINE metric 2: train epoch time -- the `while not epoch_done` loop of the reference (train_insilico.py:309-337)
with its `training_step` (train_insilico.py:124-140) restated verbatim over phoenix_b200: per-sample
odeint_adjoint solves, data loss, prior-constrained loss on a 10 000-row batch through `prior_only_forward`
(train_insilico.py:134,209-210), `composed_loss.backward()` and the six-group Adam step (train_insilico.py:245-253).
Synthetic data of the shapes in SURVEY.md 8(d).

    python tools/train_epoch.py [--config sim690|yeast|breast] [--epochs 3] [--many]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/train_epoch.py ...

Under torchrun every rank takes its share of the samples of each step and of the prior rows; ONE allreduce of the flat
gradient per optimiser step.  Prints one JSON line: seconds per epoch (max over ranks), split into the three phases.
"""
import argparse
import json
import os
import sys
import time

import torch
import torch.optim as optim

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import phoenix_b200 as pb  # noqa: E402
from phoenix_b200 import parallel  # noqa: E402

CONFIGS = {
    # name: (genes, neurons, batch, steps/epoch, dt, method, lambda, prior rows, prior scale)
    "sim350": (350, 40, 4, 138, 2.0, "dopri5", 0.99, 10000, 1.0),
    "sim690": (690, 40, 4, 138, 2.0, "dopri5", 0.99, 10000, 1.0),
    "yeast": (3551, 120, 4, 6, 5.0, "dopri5", 0.8, 10000, 4.0),
    "breast": (11165, 200, 17, 10, 0.0051, "dopri5", 0.99, 10000, 1.0),
}


def run_epochs(config, epochs=3, many=True, instrument=True, fused_prior=True, fused_adam=True, peer=True):
    """Time `epochs` epochs of the reference's training loop on this rank's share (call under torchrun for N > 1; the
    process group must already exist).  Returns the dict that main() prints (rank 0) or None."""
    import torch.distributed as dist
    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank() if world > 1 else 0
    dev = torch.device("cuda", torch.cuda.current_device())
    G, H, batch, steps, dt, method, lam, K, pscale = CONFIGS[config]
    torch.manual_seed(0)
    odenet = pb.ODENet(dev, G, neurons=H)
    parallel.broadcast_parameters(odenet)
    # explicit opt-in to lazy solver-error checking: the host runs ahead of the GPU instead of synchronising after every
    # solve (the default, which raises the reference's AssertionError at the call site); checked once per epoch below
    pb.set_sync_errors(False)
    opt = optim.Adam([
        {'params': odenet.net_sums.linear_out.weight}, {'params': odenet.net_sums.linear_out.bias},
        {'params': odenet.net_prods.linear_out.weight}, {'params': odenet.net_prods.linear_out.bias},
        {'params': odenet.net_alpha_combine.linear_out.weight},
        {'params': odenet.gene_multipliers, 'lr': 5 * 1e-3}], lr=1e-3, weight_decay=0.0,
        # the reference constructs optim.Adam(...) with defaults (train_insilico.py:245-253: the for-each implementation,
        # 42 launches per step over the six groups); fused=True is the same update as 6 launches
        **({"fused": True} if fused_adam else {}))
    if world > 1 and peer:   # gradient sum over NVLink peer memory (one kernel per rank) instead of an NCCL all-reduce
        peer = parallel.try_enable_peer_allreduce(odenet) is not None
    gen = torch.Generator(device=dev).manual_seed(1000 + rank)
    lo, hi = parallel.shard_range(batch, rank, world)
    klo, khi = parallel.shard_range(K, rank, world)
    # prior: sparse +-0.5 matrix at 3 % density; prior_grad = batch_for_prior @ prior_mat once at start-up
    prior_mat = (torch.rand(G, G, device=dev, generator=gen) < 0.03).float() * 0.5
    batch_for_prior = (torch.rand(khi - klo, 1, G, device=dev, generator=gen) - 0.5) * pscale
    prior_grad = pb.prior_grad_from_matrix(batch_for_prior, prior_mat)   # sparse product (train_insilico.py:209)
    del prior_mat
    data = torch.rand(steps, hi - lo, 1, G, device=dev, generator=gen)
    target = torch.rand(steps, hi - lo, 1, G, device=dev, generator=gen)
    tau = torch.rand(steps, hi - lo, generator=torch.Generator().manual_seed(5 + rank))
    tt = torch.stack([tau, tau + dt], dim=2)   # [steps, n, 2] on the host like datahandler.py:107
    ph = {"solve": 0.0, "prior": 0.0, "backward+opt": 0.0}

    def training_step(i, timed):
        t0 = time.perf_counter()
        opt.zero_grad()
        b, t, tg = data[i], tt[i], target[i]
        if b.shape[0] == 0:
            loss_data = torch.zeros((), device=dev)
        elif many:
            predictions = pb.odeint_adjoint_many(odenet, b, t, method=method)[:, 1]
            loss_data = torch.sum((predictions - tg) ** 2) / (batch * G)
        else:
            predictions = torch.zeros(b.shape, device=dev)
            for index, (time_, batch_point) in enumerate(zip(t, b)):
                predictions[index, :, :] = pb.odeint_adjoint(odenet, batch_point, time_, method=method)[1]
            loss_data = torch.sum((predictions - tg) ** 2) / (batch * G)
        if timed:
            torch.cuda.synchronize(); t1 = time.perf_counter()
        if fused_prior:   # the two lines of train_insilico.py:134-135 as one fused operation (phoenix_b200.prior_loss)
            loss_prior = pb.prior_loss(odenet, batch_for_prior, prior_grad) * ((khi - klo) / K)
        else:
            pred_grad = odenet.prior_only_forward(t, batch_for_prior)
            loss_prior = torch.sum((pred_grad - prior_grad) ** 2) / (K * G)
        composed_loss = lam * loss_data + (1 - lam) * loss_prior
        if timed:
            torch.cuda.synchronize(); t2 = time.perf_counter()
        composed_loss.backward()
        if world > 1:
            parallel.allreduce_grads(odenet, average=False)   # losses are normalised by the GLOBAL batch / K above
        opt.step()
        if timed:
            torch.cuda.synchronize(); t3 = time.perf_counter()
            ph["solve"] += t1 - t0; ph["prior"] += t2 - t1; ph["backward+opt"] += t3 - t2
        return loss_data, loss_prior

    for i in range(min(3, steps)):
        training_step(i, False)
    torch.cuda.synchronize()
    times = []
    for ep in range(epochs):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(steps):
            ld, lp = training_step(i, False)
        torch.cuda.synchronize()
        times.append(time.perf_counter() - t0)
        pb.check_errors()
    if instrument:
        for i in range(steps):      # one more, instrumented epoch for the phase split (synchronises between phases)
            training_step(i, True)
    pb.check_errors()
    best = min(times)
    if world > 1:
        tbest = torch.tensor([best], dtype=torch.float64, device=dev)
        dist.all_reduce(tbest, op=dist.ReduceOp.MAX)
        best = float(tbest)
    if rank != 0:
        return None
    tot = sum(ph.values()) or 1.0
    return {"metric": "train epoch time", "config": config, "genes": G, "neurons": H, "n_gpus": world,
            "batch_size": batch, "steps_per_epoch": steps, "method": method, "prior_rows": K,
            "sample_loop": "odeint_adjoint_many" if many else "per-sample odeint_adjoint",
            "prior_term": "phoenix_b200.prior_loss (fused)" if fused_prior else "prior_only_forward + torch ops",
            "optimizer": "torch.optim.Adam(fused=True)" if fused_adam else "torch.optim.Adam (for-each, the default)",
            "grad_allreduce": None if world == 1 else ("phx_peer_allreduce" if peer else "NCCL"),
            "epoch_s": best, "ms_per_step": 1e3 * best / steps,
            "phase_share": {k: v / tot for k, v in ph.items()} if instrument else None,
            "loss_data": float(ld), "loss_prior": float(lp)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="breast", choices=sorted(CONFIGS))
    ap.add_argument("--epochs", type=int, default=3)
    ap.add_argument("--many", action="store_true", help="sample loop inside the library (odeint_adjoint_many)")
    ap.add_argument("--unfused-prior", action="store_true", help="the reference's two prior-loss lines as they are")
    ap.add_argument("--foreach-adam", action="store_true", help="torch.optim.Adam with its defaults, as the reference builds it")
    ap.add_argument("--nccl", action="store_true", help="NCCL all-reduce of the gradients instead of the peer-memory kernel")
    a = ap.parse_args()
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    dev = torch.device("cuda", torch.cuda.current_device())
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    out = run_epochs(a.config, a.epochs, a.many, fused_prior=not a.unfused_prior, fused_adam=not a.foreach_adam,
                     peer=not a.nccl)
    if out is not None:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
