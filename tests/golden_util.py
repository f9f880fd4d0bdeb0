"""Helpers shared by the parity tests: load the committed golden vectors (tests/golden/*.npz, produced from the
reference by tests/golden/make_golden.py) and compare tensors."""
import json
import os

import numpy as np
import torch

from oracle.phoenix_oracle import Weights

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def manifest(kind=None):
    with open(os.path.join(GOLDEN_DIR, "manifest.json")) as fh:
        man = json.load(fh)
    return [m for m in man if kind is None or m["kind"] == kind]


def load(name):
    return dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))


def weights_of(d):
    t = lambda k: torch.from_numpy(d[k]).clone()
    return Weights(t("w_m"), t("w_Wp"), t("w_bp"), t("w_Ws"), t("w_bs"), t("w_Wa"))


def rel_l2(a, b):
    a = torch.as_tensor(a, dtype=torch.float64).reshape(-1)
    b = torch.as_tensor(b, dtype=torch.float64).reshape(-1)
    den = b.norm().item()
    return (a - b).norm().item() / (den if den > 0 else 1.0)


def compare_logs(mine, ref, dt_rtol=1e-6):
    """Return (n_equal_prefix, message).  Entries are (t0, dt, accepted)."""
    n = 0
    for a, b in zip(mine, ref):
        same = (bool(a[2]) == bool(b[2]) and abs(a[1] - b[1]) <= dt_rtol * abs(b[1])
                and abs(a[0] - b[0]) <= dt_rtol * max(abs(b[0]), abs(b[1])))
        if not same:
            return n, "first divergence at attempt %d: mine=%s ref=%s" % (n, tuple(a), tuple(b))
        n += 1
    if len(mine) != len(ref):
        return n, "length differs: mine=%d ref=%d" % (len(mine), len(ref))
    return n, "identical"


def assert_logs_close(mine, ref, dt_rtol, what="", first_rtol=1e-4):
    """Step-log parity as far as it is well defined.

    At the reference default rtol=1e-7 the dopri5 error estimate is fp32 rounding noise (the embedded error of a smooth
    problem is far below 1e-7*|y|), so the step-size sequence after the first step depends on the summation order of
    the contractions: the reference does not reproduce its own sequence between 1 and 8 CPU threads (golden field
    `stable`).  What is well defined and asserted here:
      * the first step (Hairer heuristic, misc.py:47-86) to `first_rtol`;
      * the step sizes of the first three attempts to `dt_rtol` (callers pass 2e-2 where the controller works above the
        noise floor and the reference is self-reproducible, 0.25 at the noise floor: a 2x difference in a noise-level
        error ratio moves dt by 2^(1/5) = 15 %);
      * the number of accepted steps within max(1, 15 %) and of attempts (accepted + rejected) within max(2, 25 %)
        above the noise floor, max(3, 40 %) at it: when dt sits at the accuracy limit the error ratio hovers around 1
        and whether an attempt is rejected is decided by rounding noise of the same size as the tolerance.
    """
    mine = [tuple(x) for x in mine]
    ref = [tuple(x) for x in ref]
    assert len(mine) > 0 and len(ref) > 0, what
    assert abs(mine[0][1] - ref[0][1]) <= first_rtol * abs(ref[0][1]), (what, "first dt", mine[0], ref[0])
    n, msg = compare_logs(mine, ref, dt_rtol)
    for a, b in list(zip(mine, ref))[:3]:
        assert abs(a[1] - b[1]) <= dt_rtol * abs(b[1]), (what, msg)
    acc_m, acc_r = sum(1 for x in mine if x[2]), sum(1 for x in ref if x[2])
    assert abs(acc_m - acc_r) <= max(1, int(0.15 * acc_r)), (what, "accepted", acc_m, acc_r, msg)
    att_tol = max(2, int(0.25 * len(ref))) if dt_rtol < 0.1 else max(3, int(0.4 * len(ref)))
    assert abs(len(mine) - len(ref)) <= att_tol, (what, "attempts", len(mine), len(ref), msg)
    return msg
