"""Start-up input pipeline of the reference, same file formats, vectorised parsing.

* ``readcsv``           csvreader.py:12-53: expression CSV -- row 0 = ``G,ntraj``; then per trajectory G rows of T values and
                        one row of T time stamps; blanks are NaN (ragged trajectories).  The reference parses every cell
                        with a Python ``float()`` in nested loops and draws ``np.random.normal(0, noise)`` per cell; here
                        the block is parsed by numpy in one call.  Same return tuple, same dtypes and shapes; with
                        ``noise_to_add = 0`` the arrays are identical to the reference's (tests/test_io.py).
* ``read_prior_matrix`` train_insilico.py:64-73: dense G x G CSV or 1-based (row, col, value) triplets -> dense float32.
* checkpoints            ``ODENet.save / load`` (odenet.py:100-133) keep the reference's four-file pickle format.
"""
import numpy as np
import torch


def _rows(fp):
    with open(fp, "r") as fh:
        return [line.rstrip("\n").rstrip("\r").split(",") for line in fh if line.strip("\r\n") != ""]


def readcsv(fp, device, noise_to_add, scale_expression):
    print("Reading from file {}".format(fp))
    print("Adding requested noise of {}".format(noise_to_add))
    print("Scaling gene-expression values by {} fold".format(scale_expression))
    rows = _rows(fp)
    dim, ntraj = int(float(rows[0][0])), int(float(rows[0][1]))
    data = rows[1:]
    data_np, data_pt, data_np_0noise, data_pt_0noise, t_np, t_pt = [], [], [], [], [], []
    for traj in range(ntraj):
        block = data[traj * (dim + 1):(traj + 1) * (dim + 1)]
        length = len(block[0])
        arr = np.array([[float(c) if c != "" else np.nan for c in r] for r in block], dtype=np.float64)  # [dim+1, T]
        expr = arr[:dim]
        t_row = arr[dim]
        t_np.append(np.array(t_row))
        t_pt.append(torch.tensor([float(v) for v in t_row]).to(device))   # Python floats -> float32, like the reference
        noisy = expr + np.random.normal(0, noise_to_add, size=expr.shape) if noise_to_add else expr
        traj_data = np.zeros((length, 1, dim), dtype=np.float32)
        traj_data[:, 0, :] = (scale_expression * noisy).T
        traj_0 = np.zeros((length, 1, dim), dtype=np.float32)
        traj_0[:, 0, :] = (scale_expression * expr).T
        data_np.append(traj_data)
        data_np_0noise.append(traj_0)
        data_pt.append(torch.tensor(traj_data).to(device))
        data_pt_0noise.append(torch.tensor(traj_0).to(device))
    return data_np, data_pt, t_np, t_pt, dim, ntraj, data_np_0noise, data_pt_0noise


def read_prior_matrix(prior_mat_file_loc, sparse=False, num_genes=11165):
    """Dense float32 [G, G] prior matrix from a dense CSV (``sparse=False``) or from 1-based triplets (``sparse=True``).
    Feed it to ``phoenix_b200.prior_grad_from_matrix`` (the product with the 10 000 prior rows as a sparse product)."""
    mat = np.loadtxt(prior_mat_file_loc, delimiter=",", dtype=np.float64, ndmin=2)
    if not sparse:
        return torch.from_numpy(mat).float()
    out = torch.zeros(num_genes, num_genes, dtype=torch.float64)
    idx = (torch.from_numpy(mat[:, 0].astype(np.int64) - 1), torch.from_numpy(mat[:, 1].astype(np.int64) - 1))
    out.index_put_(idx, torch.from_numpy(mat[:, 2]), accumulate=True)    # duplicates add, like sparse_coo -> to_dense
    return out.float()
