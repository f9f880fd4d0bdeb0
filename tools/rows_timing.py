"""Device time of the rows kernels (phx_solve_forward_rows / phx_solve_adjoint_rows through odeint_adjoint_many) against the
number of samples N of a step: shows what a pass of 4 lock-step rows and a partially filled last pass cost.
    python tools/rows_timing.py [G H dt] """
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import phoenix_b200 as pb  # noqa: E402


def main():
    G = int(sys.argv[1]) if len(sys.argv) > 1 else 11165
    H = int(sys.argv[2]) if len(sys.argv) > 2 else 200
    dt = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0051
    net = pb.ODENet("cuda:0", G, neurons=H)
    pb.set_sync_errors(False)
    for N in (1, 2, 3, 4, 5, 8, 12, 16, 17, 20):
        y0 = torch.rand(N, 1, G, device="cuda")
        t = torch.tensor([[0.0, dt]] * N)
        fw, bw = [], []
        for rep in range(8):
            y0g = y0.clone().requires_grad_(True)
            a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            a.record()
            y = pb.odeint_adjoint_many(net, y0g, t, method="dopri5")
            loss = (y[:, 1] ** 2).mean()
            torch.cuda.synchronize()
            b.record()
            loss.backward()
            c.record()
            torch.cuda.synchronize()
            if rep >= 3:
                fw.append(a.elapsed_time(b))
                bw.append(b.elapsed_time(c))
        print("N=%2d  forward %.1f us  backward (adjoint + unpack) %.1f us  per sample %.1f us" % (
            N, 1e3 * min(fw), 1e3 * min(bw), 1e3 * (min(fw) + min(bw)) / N), flush=True)
    pb.check_errors()


if __name__ == "__main__" and not (len(sys.argv) > 1 and sys.argv[1] == "single"):
    main()


def single_call_engines():
    """Device-limited time of ONE-sample calls (the literal loop of train_insilico.py:128-130): 20 calls enqueued back to back,
    rows kernels against the one-problem resident kernels."""
    from phoenix_b200 import engine
    G, H, dt = 11165, 200, 0.0051
    net = pb.ODENet("cuda:0", G, neurons=H)
    pb.set_sync_errors(False)
    t = torch.tensor([0.0, dt])
    y0 = torch.rand(1, G, device="cuda")
    res = {}
    for eng in ("rows", "resident"):
        engine.FORCE_ENGINE = eng
        outs = None
        for rep in range(3):
            ys = [y0.clone().requires_grad_(True) for _ in range(20)]
            a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            torch.cuda.synchronize()
            a.record()
            outs = [pb.odeint_adjoint(net, y, t, method="dopri5") for y in ys]
            b.record()
            loss = sum((o[1] ** 2).mean() for o in outs)
            torch.cuda.synchronize()
            b2 = torch.cuda.Event(enable_timing=True)
            b2.record()
            loss.backward()
            c.record()
            torch.cuda.synchronize()
        res[eng] = (outs[0].detach().clone(), net.net_sums.linear_out.weight.grad.clone())
        net.zero_grad()
        print("engine %-8s one-sample calls: forward %.1f us / call, backward %.1f us / call" % (
            eng, 1e3 * a.elapsed_time(b) / 20, 1e3 * b2.elapsed_time(c) / 20), flush=True)
    engine.FORCE_ENGINE = None
    y_r, g_r = res["rows"]
    y_s, g_s = res["resident"]
    print("rows vs resident: y identical %s (rel %.2e), grad identical %s (rel %.2e)" % (
        torch.equal(y_r, y_s), float((y_r - y_s).norm() / y_s.norm()), torch.equal(g_r, g_s),
        float((g_r - g_s).norm() / g_s.norm())))
    pb.check_errors()


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "single":
    single_call_engines()
