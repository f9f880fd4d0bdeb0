"""In-kernel phase timing of the resident solver kernels on cuda:0 (phx_ctx_set_profile): where the cycles of one
forward solve / one adjoint sweep go, as seen by CTA 0.  Usage: python tools/phase_profile.py [G H method dt reps]"""
import ctypes
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import phoenix_b200 as pb  # noqa: E402
from phoenix_b200 import _lib  # noqa: E402

NAMES = ["setup", "phaseA", "allred1", "finalize", "phaseB", "allred2", "gsp", "phaseC", "epilogue", "combine",
         "pp_d01", "pp_d2", "pp_step", "pp_interp", "pp_copy", "norms", "ctrl", "TOTAL"]


def main():
    G = int(sys.argv[1]) if len(sys.argv) > 1 else 11165
    H = int(sys.argv[2]) if len(sys.argv) > 2 else 200
    method = sys.argv[3] if len(sys.argv) > 3 else "dopri5"
    dt = float(sys.argv[4]) if len(sys.argv) > 4 else 0.0051
    reps = int(sys.argv[5]) if len(sys.argv) > 5 else 5
    N = int(sys.argv[6]) if len(sys.argv) > 6 else 1     # samples per call (rows kernels: 4 in lock-step per pass)
    lib = _lib.load()
    ctx = _lib.ctx(0)
    net = pb.ODENet("cuda:0", G, neurons=H)
    y0 = torch.rand(N, 1, G, device="cuda")
    t = torch.tensor([[0.0, dt]] * N)
    n = lib.phx_profile_slots()
    prof = torch.zeros(n, dtype=torch.int64, device="cuda")
    mhz = torch.cuda.clock_rate() / 1e3 if hasattr(torch.cuda, "clock_rate") else 1900.0

    def show(tag, evs):
        torch.cuda.synchronize()
        v = prof.tolist()
        ms = sum(a.elapsed_time(b) for a, b in evs) / len(evs)
        tot = v[len(NAMES) - 1] / reps
        print("%s  event time %.1f us/launch, CTA0 total %.0f cycles/launch (%.0f MHz nominal)" % (tag, ms * 1e3, tot, mhz))
        for name, c in zip(NAMES, v):
            if c:
                print("   %-10s %12.0f cyc  %5.1f%%  ~%.1f us" % (name, c / reps, 100.0 * c / reps / tot, c / reps / tot * ms * 1e3))
        prof.zero_()

    for _ in range(3):
        y0g = y0.clone().requires_grad_(True)
        y = pb.odeint_adjoint_many(net, y0g, t, method=method)
        (y[:, 1] ** 2).mean().backward()
    torch.cuda.synchronize()
    _lib.check(lib.phx_ctx_set_profile(ctx, ctypes.c_void_p(prof.data_ptr())), "set_profile")
    evs = []
    ys = []
    for _ in range(reps):
        y0g = y0.clone().requires_grad_(True)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        y = pb.odeint_adjoint_many(net, y0g, t, method=method)
        b.record()
        evs.append((a, b))
        ys.append((y, y0g))
    show("FORWARD  G=%d H=%d %s dt=%g N=%d" % (G, H, method, dt, N), evs)
    print("   status", pb.last_status())
    evs = []
    for y, y0g in ys:
        loss = (y[:, 1] ** 2).mean()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        loss.backward()
        b.record()
        evs.append((a, b))
    show("ADJOINT  G=%d H=%d %s dt=%g N=%d" % (G, H, method, dt, N), evs)
    print("   status", pb.last_status())
    lib.phx_ctx_set_profile(ctx, None)


if __name__ == "__main__":
    main()
