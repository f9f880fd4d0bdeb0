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
