"""cProfile of the per-sample Python path (odeint_adjoint + backward, one call per sample as train_insilico.py:128-130)
at the breast shape: where the host time of the reference-facing loop goes."""
import cProfile
import os
import pstats
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import phoenix_b200 as pb  # noqa: E402

G, H, N = 11165, 200, 17
net = pb.ODENet("cuda", G, neurons=H)
pb.set_sync_errors(False)     # lazy checks: the host runs ahead, so the profile shows pure host work
y0 = torch.rand(N, 1, G, device="cuda")
tgt = torch.rand(N, 1, G, device="cuda")
tau = torch.rand(N)
t = torch.stack([tau, tau + 0.0051], dim=1)


def step():
    net.zero_grad(set_to_none=True)
    preds = [pb.odeint_adjoint(net, y0[i], t[i], method="dopri5")[1] for i in range(N)]
    loss = torch.mean((torch.stack(preds) - tgt) ** 2)
    loss.backward()


for _ in range(3):
    step()
torch.cuda.synchronize()
# wall time per step when the host runs free (GPU time + host gaps) against the host's own time per step (enqueue only)
import time  # noqa: E402
t0 = time.perf_counter()
for _ in range(20):
    step()
t_enq = time.perf_counter() - t0
torch.cuda.synchronize()
t_all = time.perf_counter() - t0
print("per step: host enqueue %.3f ms, wall to completion %.3f ms" % (t_enq / 20 * 1e3, t_all / 20 * 1e3))
pr = cProfile.Profile()
pr.enable()
for _ in range(20):
    step()
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(28)
st.sort_stats("tottime").print_stats(22)
pb.check_errors()
