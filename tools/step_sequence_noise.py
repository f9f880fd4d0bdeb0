"""How well defined is the reference's dopri5 step sequence at its default rtol = 1e-7 / atol = 1e-9 in fp32?

Build container only (imports the UNMODIFIED reference from /root/reference).  For each default-tolerance dopri5
golden the reference is run (a) as is, (b) on the SAME ODE with the genes relabelled by a random permutation
(y0, the columns of Ws / Wp, the rows of Wa and gene_multipliers permuted consistently: mathematically the identical
initial-value problem, only the summation order of the G-long dot products changes), and the attempted-step logs
(t0, dt, accepted) are compared.  If a relabelling of the genes changes the sequence, the sequence is a property of
the summation order (fp32 rounding of the error estimate), not of the algorithm, and no implementation with a
different reduction tree (threads, SIMD width, GPU) can be expected to reproduce it beyond the first divergence.

    python tools/step_sequence_noise.py > profiles/r02_step_sequence_noise.txt
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "tests", "golden"))
sys.path.insert(0, os.path.join(HERE, "..", "tests"))
import make_golden as mg  # noqa: E402  (imports the reference + installs the step-log hook)
from golden_util import compare_logs, load, manifest, weights_of  # noqa: E402
from oracle.phoenix_oracle import Weights  # noqa: E402


def permuted(w, perm):
    return Weights(w.gene_multipliers[:, perm].contiguous(), w.Wp[:, perm].contiguous(), w.bp.clone(),
                   w.Ws[:, perm].contiguous(), w.bs.clone(), w.Wa[perm, :].contiguous())


def implied_ratio(log, i):
    """error ratio of attempt i recovered from the controller law dt' = dt * 0.9 / ratio^(1/5) (misc.py:94-103); None
    when the factor was clamped or there is no next attempt"""
    if i + 1 >= len(log):
        return None
    f = log[i + 1][1] / log[i][1]
    if f >= 10.0 - 1e-9 or abs(f - 0.2) < 1e-9 or (log[i][2] and abs(f - 1.0) < 1e-12):
        return None
    return (0.9 / f) ** 5


def fmt(row):
    return "(t0=%.9g dt=%.9g %s)" % (row[0], row[1], "acc" if row[2] else "REJ")


def main():
    torch.set_num_threads(1)
    print("# reference vs reference-with-genes-relabelled (same ODE), default rtol 1e-7 / atol 1e-9, 1 thread")
    print("%-36s %-4s %5s %5s %6s  %s" % ("golden", "leg", "n_ref", "n_prm", "prefix", "first divergence (ref | relabelled), implied error ratios"))
    for m in manifest("solve"):
        if m["method"] != "dopri5" or m["rtol"] > 1e-6 or not m["adjoint"]:
            continue
        d = load(m["name"])
        w = weights_of(d)
        y0, t, target = torch.from_numpy(d["y0"]), torch.from_numpy(d["t"]), torch.from_numpy(d["target"])
        perm = torch.randperm(w.G, generator=torch.Generator().manual_seed(5))
        _, _, _, _, f0, b0 = mg.run_solve(mg.ref_net(w), y0, t, target, "dopri5", 1)
        _, _, _, _, f1, b1 = mg.run_solve(mg.ref_net(permuted(w, perm)), y0[..., perm].contiguous(), t,
                                          target[..., perm].contiguous(), "dopri5", 1)
        for leg, a, b in (("fwd", f0, f1), ("bwd", b0, b1)):
            n, msg = compare_logs(b, a, 1e-6)
            extra = ""
            if msg != "identical" and n < min(len(a), len(b)):
                ra, rb = implied_ratio(a, n - 1) if n else None, implied_ratio(b, n - 1) if n else None
                extra = "attempt %d: %s | %s ; ratio of the attempt before: %s | %s" % (
                    n, fmt(a[n]), fmt(b[n]),
                    "%.2e" % ra if ra else "clamped", "%.2e" % rb if rb else "clamped")
            pat = "same accept/reject pattern" if [r[2] for r in a] == [r[2] for r in b] else "accept/reject pattern differs"
            print("%-36s %-4s %5d %5d %6d  %s ; %s" % (m["name"], leg, len(a), len(b), n, msg if not extra else extra, pat))


if __name__ == "__main__":
    main()
