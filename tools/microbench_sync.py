"""Raw inter-CTA signalling costs on cuda:0: tagged-slot ping-pong between two CTAs and one-hop flag barriers."""
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
    fn = lib.phx_microbench_sync
    fn.argtypes = [ctypes.c_int] * 4 + [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    fn.restype = ctypes.c_int
    slots = torch.zeros(1 << 16, dtype=torch.int64, device="cuda")
    out = torch.zeros(256, dtype=torch.int64, device="cuda")
    sp = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    iters = 2000
    for peer in (1, 2, 37, 74, 100, 147):
        slots.zero_()
        assert fn(0, 148, peer, iters, slots.data_ptr(), out.data_ptr(), sp) == 0
        torch.cuda.synchronize()
        print("pingpong CTA0 <-> CTA%-3d  %.0f cycles per round trip (2 hops)" % (peer, out[0].item() / iters))
    for what, name in ((1, "barrier, 1 line per CTA"), (2, "barrier, dense flags")):
        for n in (22, 44, 74, 148):
            slots.zero_()
            assert fn(what, n, 0, iters, slots.data_ptr(), out.data_ptr(), sp) == 0
            torch.cuda.synchronize()
            v = out[:n].double() / iters
            print("%-26s nCTA=%3d  %.0f cycles per barrier (min %.0f max %.0f over CTAs)" %
                  (name, n, v.mean().item(), v.min().item(), v.max().item()))


if __name__ == "__main__":
    main()
