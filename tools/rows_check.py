"""Quick GPU check of the rows kernels against the CPU oracle (development aid): python tools/rows_check.py [quick]"""
import os
import sys
import time

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import phoenix_b200 as pb  # noqa: E402
from oracle import phoenix_oracle as O  # noqa: E402


def rel(a, b):
    a, b = a.double().cpu().reshape(-1), b.double().cpu().reshape(-1)
    return float((a - b).norm() / (b.norm() + 1e-300))


def make_net(w):
    net = pb.ODENet("cuda", w.G, neurons=w.H)
    with torch.no_grad():
        for p, src in zip(net.parameters(), w.as_list()):
            p.copy_(src)
    return net


def case(G, H, N, method, dt, seed, dense=True, T=2):
    w = O.make_weights(G, H, seed, dense=dense)
    net = make_net(w)
    gen = torch.Generator().manual_seed(seed + 1)
    y0 = torch.rand(N, 1, G, generator=gen)
    tau = torch.rand(N, generator=gen)
    t = torch.stack([tau + dt * i * (1 + 0.3 * i) for i in range(T)], dim=1)
    target = torch.rand(N, 1, G, generator=gen)
    t0 = time.time()
    ya = y0.cuda().requires_grad_(True)
    many = pb.odeint_adjoint_many(net, ya, t, method=method)
    loss = torch.mean((many[:, -1] - target.cuda()) ** 2)
    loss.backward()
    torch.cuda.synchronize()
    el = time.time() - t0
    # oracle, sample by sample
    ey, ea, gsum = 0.0, 0.0, None
    for i in range(N):
        yr, _ = O.odeint(w, y0[i], t[i], method=method)
        gy = torch.zeros_like(yr)
        gy[-1] = 2.0 * (yr[-1] - target[i]) / target.numel()
        ady, grads, _ = O.adjoint_backward(w, t[i], yr, gy, method=method)
        ey = max(ey, rel(many[i].detach(), yr))
        ea = max(ea, rel(ya.grad[i], ady))
        gsum = grads if gsum is None else [a + b for a, b in zip(gsum, grads)]
    eg = [rel(p.grad, g) for p, g in zip(net.parameters(), gsum)]
    # bit identity against single calls
    net.zero_grad()
    yb = y0.cuda().requires_grad_(True)
    one = torch.stack([pb.odeint_adjoint(net, yb[i], t[i], method=method) for i in range(N)])
    torch.mean((one[:, -1] - target.cuda()) ** 2).backward()
    same_y = bool(torch.equal(one, many))
    same_a = bool(torch.equal(yb.grad, ya.grad))
    print("G=%5d H=%3d N=%2d T=%d %-7s dt=%-7g  y %.1e  adj_y0 %.1e  grads %s  | single==many: y %s adj %s  (%.2fs)" % (
        G, H, N, T, method, dt, ey, ea, " ".join("%.1e" % x for x in eg), same_y, same_a, el), flush=True)
    pb.check_errors()


def main():
    pb.set_sync_errors(True)
    quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
    case(37, 5, 1, "euler", 0.7, 1)
    case(37, 5, 3, "rk4", 0.7, 2)
    case(37, 5, 3, "euler", 0.4, 21, T=3)
    case(129, 33, 2, "midpoint", 0.3, 22, T=3)
    case(129, 33, 5, "rk4", 0.3, 3, T=3)
    case(129, 33, 2, "midpoint", 0.3, 4)
    case(37, 5, 2, "dopri5", 0.7, 5)
    case(350, 40, 4, "dopri5", 1.0, 6)
    case(350, 40, 6, "dopri5", 0.5, 7, T=3)
    if quick:
        return
    case(690, 40, 5, "dopri5", 1.0, 8, dense=False)
    case(1001, 100, 3, "rk4", 0.2, 9)
    case(3551, 120, 5, "rk4", 0.5, 10)
    case(3551, 120, 4, "dopri5", 0.5, 11)
    case(11165, 40, 3, "dopri5", 0.0051, 12, dense=False)
    case(11165, 200, 5, "rk4", 0.0051, 13, dense=False)
    case(11165, 200, 6, "dopri5", 0.0051, 14, dense=False)
    case(11165, 200, 2, "dopri5", 0.3, 15, dense=False)


if __name__ == "__main__":
    main()
