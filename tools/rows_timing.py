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


if __name__ == "__main__":
    main()
