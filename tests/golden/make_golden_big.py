"""Generate the BASELINE-config goldens (C3 yeast, C4 breast) on the reference's REAL data by running the UNMODIFIED
reference.  Build container only (needs /root/reference):  python tests/golden/make_golden_big.py

What is run is the data loss half of ``training_step`` (train_insilico.py:124-138) exactly as the scripts issue it:
``readcsv`` (csvreader.py:12-53) loads the shipped expression CSV, a `single`-type batch is a set of (y_i, y_{i+1})
pairs of one trajectory with their two time stamps (datahandler.py:95-120), every sample is its own
``odeint_adjoint(odenet, y0[1,G], t[2], method)`` call, ``loss = mean((pred - target)^2)`` and ONE ``backward()``
accumulates the six gradients over the samples.

  * C3  pramila_3551genes_1sample_24T.csv, H = 120 (config_yeast.cfg:4,6), batch 4: pairs starting at time points
        0, 7, 20 (the one dt = 10 gap, 100 -> 110) and 22; dopri5 and rk4; init-distribution weights
  * C4  desmedt_11165genes_1TESTsample_8middleT.csv, dt = 0.0051; pairs 0, 3, 6; H = 200 (headline) and H = 40
        (config_breast.cfg:4); dopri5

The weights are ``oracle.make_weights(G, H, seed, dense=False)`` (the reference's init distribution, odenet.py:61-75;
torch-CPU RNG => identical on the GPU box, verified through the stored checksums) copied into the reference ``ODENet``.
To keep the fixtures small the six gradients (up to 35.8 MB each case) are stored as a strided subsample + float64
norm + float64 sum per tensor; everything of size G (y, adj_y0, targets) is stored in full.
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402  (imports the reference, installs the step-log hook)

from csvreader import readcsv  # noqa: E402  (reference)

REF = "/root/reference"
YEAST = REF + "/pramila_yeast_data/clean_data/pramila_3551genes_1sample_24T.csv"
BREAST = REF + "/breast_cancer_data/clean_data/desmedt_11165genes_1TESTsample_8middleT.csv"
SUB_MAX = 16384


def subsample(t):
    flat = t.reshape(-1)
    stride = max(1, flat.numel() // SUB_MAX)
    return stride, flat[::stride].clone()


def checksum(w):
    return np.array([float(p.double().sum()) for p in w.as_list()] + [float(p.double().abs().sum()) for p in w.as_list()])


def run_batch(net, y0s, ts, targets, method, threads):
    torch.set_num_threads(threads)
    net.zero_grad()
    mg._LOG.clear()
    y0s = y0s.clone().requires_grad_(True)
    preds, flogs = [], []
    for i in range(y0s.shape[0]):
        preds.append(mg.odeint_adjoint(net, y0s[i], ts[i], method=method)[1])
        flogs.append([r for r in mg._LOG if r[0] == r[0]])
        mg._LOG.clear()
    pred = torch.stack(preds)
    loss = torch.mean((pred - targets) ** 2)
    loss.backward()
    blog = per_sample_blogs(list(mg._LOG), ts) if method == "dopri5" else [[] for _ in range(len(ts))]
    mg._LOG.clear()
    return pred.detach(), loss.detach(), y0s.grad.detach().clone(), [p.grad.detach().clone() for p in net.parameters()], \
        flogs, blog


_orig_before = mg.rk_common.RKAdaptiveStepsizeODESolver._before_integrate


def _marked_before(self, t):
    mg._LOG.append((float("nan"), float(t[0]), float("nan")))     # marker: a new adaptive solve starts at t[0]
    return _orig_before(self, t)


mg.rk_common.RKAdaptiveStepsizeODESolver._before_integrate = _marked_before


def split_sweeps(log):
    """[(t_start, [rows])] for every adaptive solve recorded in `log` (markers written by _marked_before)."""
    out = []
    for row in log:
        if row[0] != row[0]:
            out.append((row[1], []))
        else:
            out[-1][1].append(row)
    return out


def per_sample_blogs(blog, ts):
    """Backward sweeps in SAMPLE order: sweep k starts at -t1 of its sample (autograd runs them in reverse order)."""
    sweeps = split_sweeps(blog)
    assert len(sweeps) == len(ts), (len(sweeps), len(ts))
    order = list(range(len(ts)))[::-1]
    out = [None] * len(ts)
    for (t_start, rows), i in zip(sweeps, order):
        assert abs(t_start + float(ts[i][1])) < 1e-6, (t_start, ts[i])
        out[i] = rows
    return out


def case(name, csv, H, pairs, method, seed):
    data_np, data_pt, t_np, t_pt, G, ntraj, _, _ = readcsv(csv, "cpu", noise_to_add=0, scale_expression=1)
    traj, tt = data_pt[0], t_pt[0]
    y0s = torch.stack([traj[i] for i in pairs])             # [N, 1, G]
    targets = torch.stack([traj[i + 1] for i in pairs])
    ts = torch.stack([torch.stack([tt[i], tt[i + 1]]) for i in pairs])   # [N, 2] float32 like datahandler.py:112
    w = mg.make_weights(G, H, seed, dense=False)
    net = mg.ref_net(w)
    pred, loss, ady, grads, flogs, blog = run_batch(net, y0s, ts, targets, method, 1)
    stable = 1
    extra = {}
    if method == "dopri5":
        pred8, loss8, ady8, grads8, flogs8, blog8 = run_batch(net, y0s, ts, targets, method, 8)
        stable = int(flogs == flogs8 and blog == blog8)
        extra["self_y_rel"] = np.float64(((pred - pred8).norm() / pred.norm()).item())
        extra["self_grad_rel"] = np.array([((a - b).norm() / (a.norm() + 1e-30)).item() for a, b in zip(grads, grads8)])
    torch.set_num_threads(8)
    out = {"y0": y0s.numpy(), "t": ts.numpy(), "target": targets.numpy(), "pred": pred.numpy(), "loss": loss.numpy(),
           "adj_y0": ady.numpy(), "wsum": checksum(w), "stable": np.int64(stable), "pairs": np.array(pairs),
           }
    for i, (fl, bl) in enumerate(zip(flogs, blog)):
        out["flog%d" % i] = np.array(fl, dtype=np.float64).reshape(-1, 3)
        out["blog%d" % i] = np.array(bl, dtype=np.float64).reshape(-1, 3)
    for i, g in enumerate(grads):
        stride, sub = subsample(g)
        out["grad%d_sub" % i] = sub.numpy()
        out["grad%d_stride" % i] = np.int64(stride)
        out["grad%d_norm" % i] = np.float64(g.double().norm().item())
        out["grad%d_sum" % i] = np.float64(g.double().sum().item())
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out, **extra)
    return {"name": name, "kind": "batch", "G": G, "H": H, "N": len(pairs), "method": method, "seed": seed,
            "stable": stable, "fwd_steps": [len(f) for f in flogs], "bwd_steps": [len(b) for b in blog]}


def main():
    man = []
    man.append(case("c3_yeast_real_dopri5", YEAST, 120, [0, 7, 20, 22], "dopri5", 3001))
    man.append(case("c3_yeast_real_rk4", YEAST, 120, [0, 7, 20, 22], "rk4", 3001))
    man.append(case("c4_breast_real_h200_dopri5", BREAST, 200, [0, 3, 6], "dopri5", 4001))
    man.append(case("c4_breast_real_h40_dopri5", BREAST, 40, [0, 3, 6], "dopri5", 4002))
    with open(os.path.join(HERE, "manifest_big.json"), "w") as fh:
        json.dump(man, fh, indent=1)
    for m in man:
        print(m)


if __name__ == "__main__":
    main()
