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


def assert_logs_close(mine, ref, dt_rtol, what=""):
    """Step-log parity as far as it is well defined.  The reference's accept/reject sequence is decided by an error
    estimate that is at fp32 rounding level for the default rtol=1e-7 (it differs between 1 and 8 CPU threads of the
    reference itself, see tests/golden/*.npz `stable`), so exact equality is only asserted by the callers for cases
    the reference reproduces; in general we require: same first step (Hairer heuristic, well conditioned), a common
    prefix of at least 3 attempts within dt_rtol, and the same number of attempts within max(2, 15%)."""
    mine = [tuple(x) for x in mine]
    ref = [tuple(x) for x in ref]
    assert len(mine) > 0 and len(ref) > 0, what
    assert abs(mine[0][1] - ref[0][1]) <= 1e-5 * abs(ref[0][1]), (what, "first dt", mine[0], ref[0])
    n, msg = compare_logs(mine, ref, dt_rtol)
    assert n >= min(3, len(ref)), (what, msg)
    assert abs(len(mine) - len(ref)) <= max(2, int(0.15 * len(ref))), (what, msg)
    return msg
