"""How fair a stand-in is the CPU port for the reference?  `bench.py --impl reference` and `cpu_baseline` time
oracle/phoenix_oracle.py because the reference (pure Python, no setup.py) cannot travel to the GPU box.  This script runs
in the BUILD container, where /root/reference exists: the UNMODIFIED reference modules (odenet.ODENet + the vendored
torchdiffeq, driven exactly as train_insilico.py:128-138 does: odeint_adjoint per sample, MSE, backward) and the port on
the same samples of the bench workload (11 165 genes x 200 neurons, dopri5, dt = 0.0051), same thread count, one after
the other.  Prints one JSON line: seconds per sample and gene-steps/s of both, and their ratio.

    python tools/port_vs_reference.py [--samples 3] [--threads N]
"""
import argparse
import json
import os
import sys
import time

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/ode_net/code"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--samples", type=int, default=3)
    ap.add_argument("--threads", type=int, default=os.cpu_count() or 1)
    ap.add_argument("--repeats", type=int, default=5)
    a = ap.parse_args()
    if not os.path.isdir(REF):
        raise SystemExit("needs the reference checkout at %s (build container only)" % REF)
    sys.path.insert(0, REPO)
    import bench                                        # the workload generator and the port's timing loop
    from oracle import phoenix_oracle as O
    torch.set_num_threads(a.threads)
    w, y0, target, t = bench.workload(2000, "cpu")
    G = bench.G

    # ---- the port, exactly as bench.py times it (one sample per call here; the calls are interleaved with the
    # reference's below and the best of --repeats is kept: the build container's cores are shared and noisy) ----
    bench.cpu_sample(w, y0, target, t, 1)
    # ---- the unmodified reference ----
    sys.path.insert(0, REF)
    for k in [k for k in sys.modules if k == "torchdiffeq" or k.startswith("torchdiffeq.") or k == "odenet"]:
        del sys.modules[k]
    from torchdiffeq import odeint_adjoint              # noqa: E402  (reference, vendored 0.1.1)
    from odenet import ODENet                           # noqa: E402  (reference)
    net = ODENet("cpu", G, neurons=bench.H)
    net.float()
    with torch.no_grad():
        for p, src in zip(net.parameters(), w):
            p.copy_(src.reshape(p.shape))
    nfe = [0]
    fwd = net.forward

    def counted(tt, y):
        nfe[0] += 1
        return fwd(tt, y)
    net.forward = counted

    def ref_sample(i):
        net.zero_grad()
        pred = odeint_adjoint(net, y0[i], t[i], method=bench.METHOD)[1]
        loss = torch.mean((pred - target[i]) ** 2)
        loss.backward()
        return float(loss)

    ref_sample(0)
    best_r, best_p, evals_r, evals_p = [], [], 0, 0
    for i in range(a.samples):
        br = bp = None
        for _ in range(a.repeats):
            nfe[0] = 0
            t0 = time.perf_counter()
            ref_sample(i)
            dr = time.perf_counter() - t0
            er = nfe[0]
            ep, dp = bench.cpu_sample(w, y0[i:i + 1], target[i:i + 1], t[i:i + 1], 1)
            br = dr if br is None else min(br, dr)
            bp = dp if bp is None else min(bp, dp)
        best_r.append(br)
        best_p.append(bp)
        evals_r += er      # the reference's backward evaluates func once per augmented-dynamics call: one RHS
        evals_p += ep      # evaluation of the metric each, like the port's count
    s_r, s_p = sum(best_r), sum(best_p)
    # parity of what was timed: the port's gradients against the reference's on the last sample
    net.zero_grad()
    i = a.samples - 1
    pred = odeint_adjoint(net, y0[i], t[i], method=bench.METHOD)[1]
    torch.mean((pred - target[i]) ** 2).backward()
    wp = O.Weights(*w)
    y, _ = O.odeint(wp, y0[i], t[i], method=bench.METHOD)
    gy = torch.zeros_like(y)
    gy[1] = 2.0 * (y[1] - target[i]) / target[i].numel()
    _, grads, _ = O.adjoint_backward(wp, t[i], y, gy, method=bench.METHOD)
    rel = max(float((g.reshape(-1) - p.grad.reshape(-1)).norm() / p.grad.norm()) for g, p in zip(grads, net.parameters())
              if float(p.grad.norm()) > 0)
    print(json.dumps({
        "workload": "breast 11165 x 200, dopri5, dt=0.0051, %d samples (fwd + adjoint), best of %d interleaved repeats per "
                    "sample" % (a.samples, a.repeats), "threads": a.threads,
        "reference": {"s_per_sample": s_r / a.samples, "rhs_evals": evals_r, "gene_steps_per_s": G * evals_r / s_r},
        "port": {"s_per_sample": s_p / a.samples, "rhs_evals": evals_p, "gene_steps_per_s": G * evals_p / s_p},
        "port_over_reference_speed": (G * evals_p / s_p) / (G * evals_r / s_r),
        "max_grad_rel_l2_port_vs_reference": rel, "where": "build container CPU, not the GPU box's host"}))


if __name__ == "__main__":
    main()
