"""GPU check + timing of the tcgen05 batched RHS against a float64 torch evaluation of odenet.py:85-91.

    python tools/tc_check.py [--shapes G,H,B ...] [--reps N]

For each shape: relative L2 error of f = ODENet.forward(y) in the three precision modes against float64, and the
device time per call (CUDA events, L2 flushed between calls is unnecessary: operands exceed L2 at the large shapes).
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import phoenix_b200 as pb  # noqa: E402


def ref64(net, y, decay=True):
    Ws = net.net_sums.linear_out.weight.double()
    bs = net.net_sums.linear_out.bias.double()
    Wp = net.net_prods.linear_out.weight.double()
    bp = net.net_prods.linear_out.bias.double()
    Wa = net.net_alpha_combine.linear_out.weight.double()
    m = net.gene_multipliers.double()
    y = y.double()
    z = y - 0.5
    s = z / (1 + z.abs())
    l = torch.log1p(s)
    S = s @ Ws.t() + bs
    P = torch.exp(l @ Wp.t() + bp)
    J = torch.cat([S, P], -1) @ Wa.t()
    return torch.relu(m) * (J - y) if decay else J


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shapes", nargs="*", default=["350,40,256", "1001,100,300", "3551,120,1024", "11165,200,1024",
                                                    "20000,200,4096"])
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--modes", nargs="*", default=["fp32", "3xtf32", "tf32"])
    ap.add_argument("--vjp", action="store_true", help="also check / time forward + backward (state and parameter cotangents)")
    a = ap.parse_args()
    for sh in a.shapes:
        G, H, B = (int(v) for v in sh.split(","))
        torch.manual_seed(G + H + B)
        net = pb.ODENet("cuda", G, neurons=H)
        with torch.no_grad():
            for p in net.parameters():
                if p.dim() == 2 and p.shape[0] != 1:
                    p.copy_(torch.randn_like(p) * 0.05)
        y = torch.rand(B, G, device="cuda")
        with torch.no_grad():
            r = ref64(net, y)
        flops = 8.0 * B * G * H
        if a.vjp:
            gcot = torch.randn(B, G, device="cuda")
            y64 = y.double().requires_grad_(True)
            ps = [p_ for p_ in net.parameters()]
            for p_ in ps:
                p_.grad = None

            class N64:  # float64 view of the parameters with autograd
                pass
            n64 = N64()
            leaves = [p_.detach().double().requires_grad_(True) for p_ in ps]
            m64, Wp64, bp64, Ws64, bs64, Wa64 = leaves
            z = y64 - 0.5
            s_ = z / (1 + z.abs())
            l_ = torch.log1p(s_)
            J64 = torch.cat([s_ @ Ws64.t() + bs64, torch.exp(l_ @ Wp64.t() + bp64)], -1) @ Wa64.t()
            r64 = torch.relu(m64) * (J64 - y64)
            r64.backward(gcot.double())
            ref_ybar = y64.grad
            ref_grads = [t.grad for t in leaves]
            del J64, r64, z, s_, l_
        for mode in a.modes:
            pb.set_precision(mode)
            with torch.no_grad():
                f = net(None, y)
                torch.cuda.synchronize()
                err = float((f.double() - r).norm() / r.norm())
                mx = float((f.double() - r).abs().max())
                ev = [torch.cuda.Event(enable_timing=True) for _ in range(a.reps + 1)]
                ev[0].record()
                for i in range(a.reps):
                    net(None, y)
                    ev[i + 1].record()
                torch.cuda.synchronize()
                ms = min(ev[i].elapsed_time(ev[i + 1]) for i in range(a.reps))
            rec = {"G": G, "H": H, "B": B, "mode": mode, "rel_l2": err, "max_abs": mx, "ms": ms,
                   "fp32_equiv_tflops": flops / ms / 1e9}
            if a.vjp:
                def fb():
                    for p_ in net.parameters():
                        p_.grad = None
                    yg = y.clone().requires_grad_(True)
                    net(None, yg).backward(gcot)
                    return yg
                yg = fb()
                torch.cuda.synchronize()
                rel = lambda x, r_: float((x.double() - r_).norm() / r_.norm())
                rec["ybar_rel_l2"] = rel(yg.grad, ref_ybar)
                rec["grad_rel_l2"] = [rel(p_.grad, r_) for p_, r_ in zip(net.parameters(), ref_grads)]
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(a.reps):
                    fb()
                e1.record()
                torch.cuda.synchronize()
                rec["fwd_bwd_ms"] = e0.elapsed_time(e1) / a.reps
            print(json.dumps(rec), flush=True)
            if os.environ.get("PHX_TC_PROF"):
                from phoenix_b200 import _lib
                _lib.load().phx_tc_prof_dump()
                sys.stdout.flush()
        pb.set_precision("3xtf32")


if __name__ == "__main__":
    main()
