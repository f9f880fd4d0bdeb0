"""A/B of the streaming engine's fused stage algebra (csrc/phx_stream.cu `fuse`, PHX_STREAM_FUSE) on forward solves:
the same odeint call with the RK stage algebra in the epilogue of the joint contraction (1) and as stand-alone
elementwise kernels (0), over shapes from one wave of joint tiles to many.

    python tools/stream_fuse_ab.py [--shapes 11165x200x60,3551x120x256,...]

Prints one JSON line per (shape, method): tiles of the joint contraction and ms fused / unfused (best of 3, CUDA events).
Result on B200 (profiles/r04d_stream_fuse_ab.txt): bit-identical, and the stand-alone kernels win at every shape -- the
library default (PHX_STREAM_FUSE unset) is therefore "unfused"; that run still printed the single-wave rule it was testing
as "auto"."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import phoenix_b200 as pb  # noqa: E402


def timed(fn, reps=3):
    best = None
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        best = ms if best is None else min(best, ms)
    return best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shapes", default="11165x200x60,3551x120x256,3551x120x1024,3551x120x4096,11165x200x1024,"
                                        "20000x200x4096")
    a = ap.parse_args()
    pb.set_sync_errors(True)
    pb.engine.FORCE_ENGINE = "stream"
    for shape in a.shapes.split(","):
        G, H, B = [int(x) for x in shape.split("x")]
        torch.manual_seed(5)
        net = pb.ODENet("cuda", G, neurons=H)
        y0 = torch.rand(B, G, device="cuda")
        tiles = ((G + 127) // 128) * ((B + 255) // 256)
        for method, t, kw in (("rk4", torch.tensor([0.0, 0.1]), {}),
                              ("dopri5", torch.linspace(0.0, 0.1, 10, dtype=torch.float64), {})):
            out, ys = {}, {}
            for fuse in ("1", "0"):
                os.environ["PHX_STREAM_FUSE"] = fuse

                def run():
                    with torch.no_grad():
                        ys[fuse] = pb.odeint(net, y0, t, method=method, **kw)
                run()
                out[fuse] = timed(run)
            os.environ.pop("PHX_STREAM_FUSE", None)
            print(json.dumps({"G": G, "H": H, "rows": B, "method": method, "output_times": len(t), "joint_tiles": tiles,
                              "waves": tiles / 148.0, "ms_fused": out["1"], "ms_unfused": out["0"],
                              "default": "unfused",
                              "bit_identical": bool(torch.equal(ys["1"], ys["0"]))}), flush=True)
        del net, y0
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
