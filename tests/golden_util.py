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


# ---- BASELINE-config goldens on the reference's real data (tests/golden/make_golden_big.py) ------------------------------
def manifest_big():
    with open(os.path.join(GOLDEN_DIR, "manifest_big.json")) as fh:
        return json.load(fh)


def big_case_inputs(m, d):
    """(weights, y0 [N,1,G], t [N,2], target [N,1,G]) of a `batch` golden; the weights are regenerated from the seed and
    verified against the stored checksums."""
    from oracle.phoenix_oracle import make_weights
    w = make_weights(m["G"], m["H"], m["seed"], dense=False)
    chk = np.array([float(p.double().sum()) for p in w.as_list()] + [float(p.double().abs().sum()) for p in w.as_list()])
    assert np.allclose(chk, d["wsum"], rtol=1e-12, atol=0), "regenerated weights differ from the golden's"
    return w, torch.from_numpy(d["y0"]), torch.from_numpy(d["t"]), torch.from_numpy(d["target"])


def check_big_case(m, d, pred, loss, adj_y0, grads, ytol, gtol):
    """Compare outputs with a `batch` golden: full tensors of size G, strided subsample + norm of the six gradients."""
    assert rel_l2(pred, d["pred"]) < ytol, ("pred", rel_l2(pred, d["pred"]))
    assert abs(float(loss) - float(d["loss"])) <= 10 * ytol * abs(float(d["loss"]))
    assert rel_l2(adj_y0, d["adj_y0"]) < gtol, ("adj_y0", rel_l2(adj_y0, d["adj_y0"]))
    for i, g in enumerate(grads):
        g = torch.as_tensor(g).detach().cpu()
        stride = int(d["grad%d_stride" % i])
        sub = g.reshape(-1)[::stride]
        ref_norm = float(d["grad%d_norm" % i])
        # the subsample is compared against the FULL tensor's scale (a strided sample of a sparse gradient may be tiny)
        scale = ref_norm * (sub.numel() / g.numel()) ** 0.5
        err = (sub.double() - torch.from_numpy(d["grad%d_sub" % i]).double()).norm().item()
        assert err <= gtol * max(scale, 1e-30), (m["name"], "grad", i, err / max(scale, 1e-30))
        assert abs(g.double().norm().item() - ref_norm) <= gtol * ref_norm, (m["name"], "grad norm", i)
