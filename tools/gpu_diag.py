"""Run every golden case through the CUDA path on cuda:0 and print the discrepancies (no asserts): a one-shot
diagnostic for gpurun.  Output goes to stdout and gpurun_out/diag.txt."""
import os
import sys
import time
import traceback

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))

import phoenix_b200 as pb  # noqa: E402
from golden_util import compare_logs, load, manifest, rel_l2  # noqa: E402

OUT = []


def say(*a):
    s = " ".join(str(x) for x in a)
    print(s, flush=True)
    OUT.append(s)


def net_from(d, dev="cuda"):
    G, H = d["w_Ws"].shape[1], d["w_Ws"].shape[0]
    net = pb.ODENet(dev, G, neurons=H)
    with torch.no_grad():
        net.gene_multipliers.copy_(torch.from_numpy(d["w_m"]))
        net.net_prods.linear_out.weight.copy_(torch.from_numpy(d["w_Wp"]))
        net.net_prods.linear_out.bias.copy_(torch.from_numpy(d["w_bp"]))
        net.net_sums.linear_out.weight.copy_(torch.from_numpy(d["w_Ws"]))
        net.net_sums.linear_out.bias.copy_(torch.from_numpy(d["w_bs"]))
        net.net_alpha_combine.linear_out.weight.copy_(torch.from_numpy(d["w_Wa"]))
    return net


def main():
    pb.set_sync_errors(True)
    pb.set_step_logging(True)
    say("device", torch.cuda.get_device_name(0))
    for m in manifest("rhs"):
        try:
            d = load(m["name"])
            net = net_from(d)
            for tag, fn in (("decay", net.forward), ("prior", net.prior_only_forward)):
                net.zero_grad()
                y = torch.from_numpy(d["y"]).cuda().requires_grad_(True)
                g = torch.from_numpy(d["g"]).cuda()
                f = fn(None, y)
                f.backward(g)
                errs = [rel_l2(f.detach().cpu(), d["f_" + tag]), rel_l2(y.grad.cpu(), d["ybar_" + tag])]
                for i, p in enumerate(net.parameters()):
                    ref = d["pbar%d_%s" % (i, tag)]
                    got = p.grad.cpu() if p.grad is not None else torch.zeros_like(torch.from_numpy(ref))
                    errs.append(rel_l2(got, ref))
                say("RHS", m["name"], tag, " ".join("%.1e" % e for e in errs))
        except Exception:
            say("RHS", m["name"], "EXC", traceback.format_exc())
    for m in manifest("solve"):
        try:
            d = load(m["name"])
            net = net_from(d)
            y0 = torch.from_numpy(d["y0"]).cuda().requires_grad_(True)
            t = torch.from_numpy(d["t"])
            target = torch.from_numpy(d["target"]).cuda()
            t0 = time.time()
            if m["adjoint"]:
                y = pb.odeint_adjoint(net, y0, t, rtol=float(d["rtol"]), atol=float(d["atol"]), method=m["method"])
            else:
                with torch.no_grad():
                    y = pb.odeint(net, y0, t, rtol=float(d["rtol"]), atol=float(d["atol"]), method=m["method"])
            flog = pb.last_step_log()
            st = pb.last_status()
            msg = ["SOLVE", m["name"], "y %.1e" % rel_l2(y.detach().cpu(), d["y"]), "st", st]
            if m["method"] == "dopri5":
                msg += ["flog:", compare_logs(flog, d["flog"])[1], "stable=%d" % int(d["stable"]),
                        "dts mine", " ".join("%.4g" % r[1] for r in flog[:6]), "ref",
                        " ".join("%.4g" % r[1] for r in d["flog"][:6])]
            say(*msg)
            if m["adjoint"]:
                loss = torch.mean((y[1:] - target) ** 2)
                loss.backward()
                blog = pb.last_step_log()
                st = pb.last_status()
                errs = [rel_l2(y0.grad.cpu(), d["adj_y0"])] + [rel_l2(p.grad.cpu(), d["grad%d" % i])
                                                               for i, p in enumerate(net.parameters())]
                msg = ["  ADJ", "loss %.3e/%.3e" % (loss.item(), float(d["loss"])),
                       " ".join("%.1e" % e for e in errs), "st", st]
                if m["method"] == "dopri5":
                    msg += ["blog:", compare_logs(blog, d["blog"], 1e-4)[1]]
                    msg += ["ref-self-noise %.1e" % float(max(d["self_grad_rel"]))]
                say(*msg)
            say("   time %.3fs" % (time.time() - t0))
        except Exception:
            say("SOLVE", m["name"], "EXC", traceback.format_exc().splitlines()[-1])
    os.makedirs(os.path.join(REPO, "gpurun_out"), exist_ok=True)
    with open(os.path.join(REPO, "gpurun_out", "diag.txt"), "w") as fh:
        fh.write("\n".join(OUT) + "\n")


if __name__ == "__main__":
    main()
