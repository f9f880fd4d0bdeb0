"""Micro-benchmarks of the resident-kernel building blocks on cuda:0 (all-reduce, scalar sums, one weight pass)."""
import ctypes
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import phoenix_b200 as pb  # noqa: E402
from phoenix_b200 import _lib, engine  # noqa: E402


def main():
    mb = os.path.join(os.path.dirname(os.path.abspath(__file__)), "experiments", "libphx_microbench.so")
    if not os.path.exists(mb):   # diagnostics live outside the product library
        import subprocess
        subprocess.check_call([os.path.join(os.path.dirname(mb), "build_microbench.sh")])
    ctypes.CDLL(_lib.LIB_PATH, mode=ctypes.RTLD_GLOBAL)
    lib = ctypes.CDLL(mb)
    _lib.load()
    ctx = _lib.ctx(0)
    fn = lib.phx_microbench
    fn.argtypes = [ctypes.c_void_p] + [ctypes.c_int] * 6 + [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t,
                                                            ctypes.c_void_p, ctypes.c_void_p]
    fn.restype = ctypes.c_int
    for G, H in ((11165, 200), (3551, 120), (690, 40), (350, 40)):
        net = pb.ODENet("cuda:0", G, neurons=H)
        packed, _, _, _ = engine.packed_weights(net)
        nb = _lib.load().phx_solve_workspace_bytes(ctx, G, H, 1, 2, 0)
        ws = torch.zeros(nb, dtype=torch.uint8, device="cuda")
        out = torch.zeros(200 * 1024, device="cuda")
        K2 = 2 * ((H + 3) // 4 * 4)
        sp = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        for what, name, n in ((0, "allreduce_f n=K2", K2), (1, "grid_sum_d nd=4", 4), (2, "pass over W1", 0), (3, "fused pass B", 0), (4, "fused pass B+A", 0)):
            iters = 200
            res = []
            for rep in range(3):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                rc = fn(ctx, G, H, 1, what, n, iters, ctypes.c_void_p(packed.data_ptr()),
                        ctypes.c_void_p(ws.data_ptr()), ws.numel(), ctypes.c_void_p(out.data_ptr()), sp)
                b.record()
                torch.cuda.synchronize()
                assert rc > 0, _lib.last_error()
                res.append(a.elapsed_time(b) * 1e3 / iters)
            print("G=%5d H=%3d nCTA=%3d  %-20s %.2f us/op (runs: %s)  out0=%.3f" %
                  (G, H, rc, name, min(res), " ".join("%.2f" % r for r in res), float(out[0])), flush=True)


if __name__ == "__main__":
    main()
