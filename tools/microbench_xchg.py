"""Where the cycles of the two-phase all-reduce go (instrumented copy in tools/experiments/phx_microbench.cu)."""
import ctypes
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from phoenix_b200 import _lib  # noqa: E402


def main():
    mb = os.path.join(os.path.dirname(os.path.abspath(__file__)), "experiments", "libphx_microbench.so")
    if not os.path.exists(mb):   # diagnostics live outside the product library
        import subprocess
        subprocess.check_call([os.path.join(os.path.dirname(mb), "build_microbench.sh")])
    ctypes.CDLL(_lib.LIB_PATH, mode=ctypes.RTLD_GLOBAL)
    lib = ctypes.CDLL(mb)
    L = _lib.load()
    ctx = _lib.ctx(0)
    fn = lib.phx_microbench_xchg
    fn.argtypes = [ctypes.c_void_p] + [ctypes.c_int] * 5 + [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    fn.restype = ctypes.c_int
    G, H = 11165, 200
    nb = L.phx_solve_workspace_bytes(ctx, G, H, 1, 2, 0)
    ws = torch.zeros(nb, dtype=torch.uint8, device="cuda")
    out = torch.zeros(160 * 8, dtype=torch.int64, device="cuda")
    sp = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    iters = 500
    names = ["puts issued", "partials seen", "result posted", "results seen", "after barrier"]
    for variant in (1, 4, 1, 4):
        for n in (400, 800):
            out.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            nc = fn(ctx, G, H, variant, n, iters, ctypes.c_void_p(ws.data_ptr()), ctypes.c_void_p(out.data_ptr()), sp)
            b.record()
            torch.cuda.synchronize()
            assert nc > 0, _lib.last_error()
            v = out.view(160, 8)[:nc, :5].double() / iters
            print("variant %d n=%d: %.2f us/op" % (variant, n, a.elapsed_time(b) * 1e3 / iters))
            for k, nm in enumerate(names):
                col = v[:, k]
                print("   %-14s mean %6.0f  min %6.0f  max %6.0f  (CTA0 %6.0f, CTA%d %6.0f)" %
                      (nm, col.mean(), col.min(), col.max(), col[0], nc - 1, col[nc - 1]))


if __name__ == "__main__":
    main()
