"""Golden for the prior-constrained loss path from the UNMODIFIED reference.  Build container only:
    python tests/golden/make_golden_prior.py

What runs is train_insilico.py:64-68 (`read_prior_matrix` on the SHIPPED edge_prior_matrix_G690_noise_0.0.csv),
:208-209 (`batch_for_prior`, `prior_grad = torch.matmul(batch_for_prior, prior_mat)`) and :134-138 (`pred_grad =
odenet.prior_only_forward(t, batch_for_prior)`, `loss_prior = torch.mean((pred_grad - prior_grad) ** 2)`, backward)
with the reference's own modules.  `train_insilico` is imported with matplotlib stubbed (not installed here; it is only
used for plots).  256 prior rows instead of 10 000 keep the fixture small."""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402  (puts the reference on sys.path)

for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.lines", "matplotlib.patches", "matplotlib.font_manager",
             "mpl_toolkits", "mpl_toolkits.axes_grid1"):
    sys.modules.setdefault(name, types.ModuleType(name))
sys.modules["matplotlib.lines"].Line2D = object
sys.modules["mpl_toolkits.axes_grid1"].make_axes_locatable = object
import train_insilico as ti  # noqa: E402  (reference)

PRIOR = "/root/reference/ground_truth_simulator/clean_data/edge_prior_matrix_G690_noise_0.0.csv"


def main():
    G, H, K = 690, 40, 256
    prior_mat = ti.read_prior_matrix(PRIOR, sparse=False, num_genes=G)
    torch.manual_seed(11)
    batch_for_prior = torch.rand(K, 1, G) - 0.5
    prior_grad = torch.matmul(batch_for_prior, prior_mat)
    w = mg.make_weights(G, H, 6001, dense=False)
    net = mg.ref_net(w)
    net.zero_grad()
    pred_grad = net.prior_only_forward(None, batch_for_prior)
    loss_prior = torch.mean((pred_grad - prior_grad) ** 2)
    loss_prior.backward()
    nz = prior_mat.nonzero()
    out = {"batch_for_prior": batch_for_prior.numpy(), "prior_grad": prior_grad.numpy(),
           "prior_rows": nz[:, 0].numpy().astype(np.int32), "prior_cols": nz[:, 1].numpy().astype(np.int32),
           "prior_vals": prior_mat[nz[:, 0], nz[:, 1]].numpy(), "loss_prior": loss_prior.detach().numpy(),
           "seed": np.int64(6001), "G": np.int64(G), "H": np.int64(H)}
    for i, p in enumerate(net.parameters()):
        out["grad%d" % i] = (torch.zeros_like(p) if p.grad is None else p.grad).numpy()
    np.savez_compressed(os.path.join(HERE, "prior_g690_h40_k256.npz"), **out)
    print("nnz", nz.shape[0], "loss", float(loss_prior), {k: v.shape for k, v in out.items() if hasattr(v, "shape")})


if __name__ == "__main__":
    main()
