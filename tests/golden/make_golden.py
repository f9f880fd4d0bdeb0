"""Generate the golden vectors under tests/golden/ by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference):  python tests/golden/make_golden.py

The reference (QuackenbushLab/phoenix) ships no tests or golden vectors for this path, so
parity is pinned on outputs of the reference's own modules: ``odenet.ODENet`` and the vendored ``torchdiffeq``
(``/root/reference/ode_net/code``), imported here via sys.path and driven exactly as the training scripts do
(``odeint_adjoint(odenet, y0, t, method=...)`` then ``loss.backward()``, train_insilico.py:124-140).

Each case is written as ``<name>.npz`` holding the inputs (six weights, y0, t, target) and the reference outputs
(f, autograd VJPs, y(t), loss, adj_y0, six gradients, attempted-step logs ``(t0, dt, accepted)`` for the forward and
the adjoint solve).  dopri5 cases are run with 1 and with 8 CPU threads; when the two step logs differ the case is
marked ``stable=0`` — the reference is then not reproducible against ITSELF at the default rtol=1e-7 (fp32 noise
decides accept/reject), and parity tests fall back to value tolerances for that case.
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, "/root/reference/ode_net/code")
sys.path.insert(0, REPO)

from torchdiffeq import odeint, odeint_adjoint  # noqa: E402  (reference, vendored 0.1.1)
from torchdiffeq._impl import rk_common  # noqa: E402
from odenet import ODENet  # noqa: E402  (reference)

from oracle.phoenix_oracle import make_weights  # noqa: E402  (only the deterministic input generator)

_LOG = []
_orig_step = rk_common.RKAdaptiveStepsizeODESolver._adaptive_step


def _logged_step(self, st):
    new = _orig_step(self, st)
    _LOG.append((float(st.t1), float(st.dt), float(bool(new.t1 != st.t1))))
    return new


rk_common.RKAdaptiveStepsizeODESolver._adaptive_step = _logged_step


def ref_net(w):
    net = ODENet("cpu", w.G, neurons=w.H)
    net.float()
    with torch.no_grad():
        net.gene_multipliers.copy_(w.gene_multipliers)
        net.net_prods.linear_out.weight.copy_(w.Wp)
        net.net_prods.linear_out.bias.copy_(w.bp)
        net.net_sums.linear_out.weight.copy_(w.Ws)
        net.net_sums.linear_out.bias.copy_(w.bs)
        net.net_alpha_combine.linear_out.weight.copy_(w.Wa)
    names = [n for n, _ in net.named_parameters()]
    assert names == ["gene_multipliers", "net_prods.linear_out.weight", "net_prods.linear_out.bias",
                     "net_sums.linear_out.weight", "net_sums.linear_out.bias",
                     "net_alpha_combine.linear_out.weight"], names
    return net


def wdict(w):
    return {"w_m": w.gene_multipliers.numpy(), "w_Wp": w.Wp.numpy(), "w_bp": w.bp.numpy(), "w_Ws": w.Ws.numpy(),
            "w_bs": w.bs.numpy(), "w_Wa": w.Wa.numpy()}


def case_rhs(name, G, H, B, seed, dense, lo, hi, neg):
    w = make_weights(G, H, seed, dense=dense, neg_mult_frac=neg)
    net = ref_net(w)
    gen = torch.Generator().manual_seed(seed + 7)
    y = (torch.rand(B, 1, G, generator=gen) * (hi - lo) + lo).requires_grad_(True)
    g = torch.randn(B, 1, G, generator=gen)
    out = {}
    for decay in (1, 0):
        net.zero_grad()
        if y.grad is not None:
            y.grad = None
        f = net.forward(None, y) if decay else net.prior_only_forward(None, y)
        f.backward(g)
        tag = "decay" if decay else "prior"
        out["f_" + tag] = f.detach().numpy()
        out["ybar_" + tag] = y.grad.detach().numpy().copy()
        for i, p in enumerate(net.parameters()):
            out["pbar%d_%s" % (i, tag)] = (torch.zeros_like(p) if p.grad is None else p.grad).numpy().copy()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), y=y.detach().numpy(), g=g.numpy(), **wdict(w), **out)
    return {"name": name, "kind": "rhs", "G": G, "H": H, "B": B}


def run_solve(net, y0, t, target, method, threads, adjoint=True, rtol=1e-7, atol=1e-9):
    torch.set_num_threads(threads)
    net.zero_grad()
    _LOG.clear()
    if not adjoint:  # forward-only callers (validation / influence scan run under no_grad)
        with torch.no_grad():
            y = odeint(net, y0, t, rtol=rtol, atol=atol, method=method)
        flog = list(_LOG)
        _LOG.clear()
        loss = torch.mean((y[1:] - target) ** 2)
        return y.clone(), loss, torch.zeros_like(y0), [torch.zeros_like(p) for p in net.parameters()], flog, []
    y0 = y0.clone().requires_grad_(True)
    y = odeint_adjoint(net, y0, t, rtol=rtol, atol=atol, method=method)
    flog = list(_LOG)
    _LOG.clear()
    loss = torch.mean((y[1:] - target) ** 2)
    loss.backward()
    blog = list(_LOG)
    _LOG.clear()
    return (y.detach().clone(), loss.detach().clone(), y0.grad.detach().clone(),
            [p.grad.detach().clone() for p in net.parameters()], flog, blog)


def case_solve(name, G, H, B, seed, dense, method, times, t_dtype, lo=0.0, hi=1.0, neg=0.0, squeeze=False,
               adjoint=True, rtol=1e-7, atol=1e-9):
    """``squeeze=False``: y0 is [B,1,G] (how batches reach odeint, datahandler.py:87-120); True: [1,G] per-sample."""
    w = make_weights(G, H, seed, dense=dense, neg_mult_frac=neg)
    net = ref_net(w)
    gen = torch.Generator().manual_seed(seed + 11)
    shape = (1, G) if squeeze else (B, 1, G)
    y0 = torch.rand(*shape, generator=gen) * (hi - lo) + lo
    t = torch.tensor(times, dtype=t_dtype)
    target = torch.rand(len(times) - 1, *shape, generator=gen)
    y, loss, ady, grads, flog, blog = run_solve(net, y0, t, target, method, 1, adjoint, rtol, atol)
    stable = 1
    extra = {}
    if method == "dopri5":
        y8, loss8, ady8, grads8, flog8, blog8 = run_solve(net, y0, t, target, method, 8, adjoint, rtol, atol)
        stable = int(flog == flog8 and blog == blog8)
        # the reference's own 1-thread vs 8-thread discrepancy = its noise floor for this case
        extra["self_y_rel"] = np.float64(((y - y8).norm() / y.norm()).item())
        extra["self_grad_rel"] = np.array([((a - b).norm() / (a.norm() + 1e-30)).item()
                                           for a, b in zip(grads, grads8)])
        extra["flog8"] = np.array(flog8, dtype=np.float64).reshape(-1, 3)
        extra["blog8"] = np.array(blog8, dtype=np.float64).reshape(-1, 3)
    torch.set_num_threads(8)
    np.savez_compressed(
        os.path.join(HERE, name + ".npz"), y0=y0.numpy(), t=t.numpy(), target=target.numpy(), y=y.numpy(),
        loss=loss.numpy(), adj_y0=ady.numpy(), flog=np.array(flog, dtype=np.float64).reshape(-1, 3),
        blog=np.array(blog, dtype=np.float64).reshape(-1, 3), stable=np.int64(stable),
        has_adjoint=np.int64(int(adjoint)), rtol=np.float64(rtol), atol=np.float64(atol),
        **{"grad%d" % i: g.numpy() for i, g in enumerate(grads)}, **wdict(w), **extra)
    return {"name": name, "kind": "solve", "G": G, "H": H, "B": B, "method": method, "T": len(times),
            "stable": stable, "adjoint": int(adjoint), "rtol": rtol, "atol": atol, "fwd_steps": len(flog), "bwd_steps": len(blog)}


def main():
    torch.manual_seed(0)
    man = []
    # --- RHS / VJP (SURVEY 8a rows a1-a4, a15 formulas) ---
    man.append(case_rhs("rhs_g37_h5_b3", 37, 5, 3, 101, True, -0.5, 1.5, 0.3))
    man.append(case_rhs("rhs_g350_h40_b1", 350, 40, 1, 102, False, 0.0, 1.0, 0.0))
    man.append(case_rhs("rhs_g350_h40_b4_dense", 350, 40, 4, 103, True, -0.5, 1.5, 0.2))
    man.append(case_rhs("rhs_g129_h33_b17", 129, 33, 17, 104, True, 0.0, 1.0, 0.1))
    # --- fixed-step solves (a7) ---
    for method in ("euler", "rk4", "midpoint"):
        man.append(case_solve("solve_%s_g37_h5_b1" % method, 37, 5, 1, 201, True, method, [0.0, 0.7], torch.float32,
                              squeeze=True))
        man.append(case_solve("solve_%s_g350_h40_b1" % method, 350, 40, 1, 202, False, method, [0.0, 2.0],
                              torch.float32, squeeze=True))
    # NB the reference's fixed-grid ADJOINT raises for float64 t (zeros_like(t) promotes the flat augmented state
    # to double, adjoint.py:119 + misc.py:153), so the float64-t fixed-grid case is forward-only.
    man.append(case_solve("solve_rk4_g129_h33_b5_t4", 129, 33, 5, 203, True, "rk4", [0.0, 0.3, 0.5, 1.1],
                          torch.float32, neg=0.1))
    man.append(case_solve("solve_rk4_g129_h33_b5_t4_f64_fwd", 129, 33, 5, 205, True, "rk4", [0.0, 0.3, 0.5, 1.1],
                          torch.float64, neg=0.1, adjoint=False))
    man.append(case_solve("solve_dopri5_g97_h12_b60_t10_fwd", 97, 12, 60, 206, True, "dopri5",
                          list(np.arange(0, 1, 0.1)), torch.float64, lo=-0.5, hi=0.5, adjoint=False))
    man.append(case_solve("solve_euler_g129_h33_b5_t4", 129, 33, 5, 204, True, "euler", [0.0, 0.3, 0.5, 1.1],
                          torch.float32, neg=0.1))
    # --- dopri5 (a8-a12) + adjoint (a13-a16) ---
    man.append(case_solve("solve_dopri5_g37_h5_b1", 37, 5, 1, 301, True, "dopri5", [0.0, 0.7], torch.float32,
                          squeeze=True))
    man.append(case_solve("solve_dopri5_g350_h40_b1_sparse", 350, 40, 1, 302, False, "dopri5", [0.0, 2.0],
                          torch.float32, squeeze=True))
    man.append(case_solve("solve_dopri5_g350_h40_b1_dense", 350, 40, 1, 303, True, "dopri5", [2.0, 3.0],
                          torch.float32, squeeze=True))
    man.append(case_solve("solve_dopri5_g129_h33_b5_t4", 129, 33, 5, 304, True, "dopri5", [0.0, 0.3, 0.5, 1.1],
                          torch.float64, neg=0.1))
    man.append(case_solve("solve_dopri5_g64_h16_b3_neg", 64, 16, 3, 305, True, "dopri5", [0.0, 0.1, 0.2, 0.3, 0.4],
                          torch.float64, lo=-0.5, hi=0.5))
    man.append(case_solve("solve_dopri5_g690_h40_b1", 690, 40, 1, 306, False, "dopri5", [0.0, 2.0], torch.float32,
                          squeeze=True))
    # --- loose tolerances: the error estimate is far above fp32 rounding noise, so the accepted/rejected step
    # sequence is reproducible and pins the controller semantics (a9-a11) independently of summation order ---
    man.append(case_solve("solve_dopri5_loose_g350_h40_b1", 350, 40, 1, 401, True, "dopri5", [0.0, 6.0],
                          torch.float32, squeeze=True, rtol=1e-3, atol=1e-5))
    man.append(case_solve("solve_dopri5_loose_g129_h33_b3_t4", 129, 33, 3, 402, True, "dopri5",
                          [0.0, 1.5, 2.0, 7.0], torch.float64, neg=0.1, rtol=1e-3, atol=1e-5))
    man.append(case_solve("solve_dopri5_loose_g690_h40_b1", 690, 40, 1, 403, True, "dopri5", [0.0, 9.0],
                          torch.float32, squeeze=True, rtol=1e-4, atol=1e-6))
    man.append(case_solve("solve_dopri5_loose_g97_h12_b60_t10_fwd", 97, 12, 60, 404, True, "dopri5",
                          list(np.arange(0, 10, 1.0)), torch.float64, lo=-0.5, hi=0.5, adjoint=False, rtol=1e-3,
                          atol=1e-5))
    with open(os.path.join(HERE, "manifest.json"), "w") as fh:
        json.dump(man, fh, indent=1)
    for m in man:
        print(m)


if __name__ == "__main__":
    main()
