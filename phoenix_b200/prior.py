"""The prior-constrained loss term of the reference's ``training_step`` (train_insilico.py:134-137) and its one-time set-up
(train_insilico.py:208-209) on the GPU.

The reference computes, every optimiser step,

    pred_grad  = odenet.prior_only_forward(t, batch_for_prior)          # [10000, 1, G]
    loss_prior = torch.mean((pred_grad - prior_grad) ** 2)

and back-propagates through it.  ``prior_loss`` is the same term as ONE fused operation (opt-in; the unfused lines above
keep working unchanged): the tcgen05 joint contraction compares with ``prior_grad`` in its epilogue, so the 10 000 x G
joint, the difference, its square and the cotangent 2 (J - prior_grad) / N are never separate passes over memory, and
the backward reuses the branch vector [S|P] left in the workspace.  ``prior_grad_from_matrix`` is
``torch.matmul(batch_for_prior, prior_mat)`` for the 0.5-3 % dense prior matrices PHOENIX uses, as a sparse product.
"""
import ctypes

import torch

from . import _lib, engine


def prior_grad_from_matrix(batch_for_prior, prior_mat):
    """``torch.matmul(batch_for_prior, prior_mat)`` (train_insilico.py:209) with ``prior_mat`` [G, G] (dense or sparse
    torch tensor, any device) taken as a sparse matrix.  ``batch_for_prior``: CUDA float32 ``[..., G]``."""
    x = batch_for_prior.detach().to(torch.float32).contiguous()
    dev = engine._device_index(x)
    G = x.shape[-1]
    B = x.numel() // G
    pm = prior_mat.to_dense() if prior_mat.is_sparse else prior_mat
    if tuple(pm.shape) != (G, G):
        raise RuntimeError("prior_mat must be [%d, %d] (got %s)" % (G, G, tuple(pm.shape)))
    pm = pm.to(device=x.device, dtype=torch.float32)
    cr = pm.t().contiguous().nonzero()          # (column, row) pairs, sorted by column then row
    val = pm.t()[cr[:, 0], cr[:, 1]].contiguous()
    counts = torch.bincount(cr[:, 0], minlength=G)
    colptr = torch.zeros(G + 1, dtype=torch.int32, device=x.device)
    colptr[1:] = torch.cumsum(counts, 0).to(torch.int32)
    rowidx = cr[:, 1].to(torch.int32).contiguous()
    out = torch.empty_like(x)
    lib = _lib.load()
    with torch.cuda.device(dev):
        _lib.check(lib.phx_prior_setup(_lib.ctx(dev), G, B, engine._ptr(x), engine._ptr(colptr), engine._ptr(rowidx),
                                       engine._ptr(val), engine._ptr(out), engine._stream_ptr(dev)), "prior_setup")
    return out


_planes_for = engine.hill_planes_for
_hill_cache = engine.hill_cache


class _PriorLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, net, x, prior_grad, *params):
        packed, G, H, dev = engine.packed_weights(net)
        x2 = x.detach().to(torch.float32).contiguous()
        pg = prior_grad.detach().to(torch.float32).contiguous()
        if x2.shape != pg.shape or x2.shape[-1] != G:
            raise RuntimeError("batch_for_prior and prior_grad must both be [..., %d]" % G)
        B = x2.numel() // G
        lib = _lib.load()
        with torch.cuda.device(dev):
            ws = engine._workspace(dev, lib.phx_rhs_workspace_bytes(G, H, B), "rhs")
            gcot = torch.empty_like(x2)
            loss = torch.empty(1, dtype=torch.float32, device=x2.device)
            planes = engine.hill_planes_for(x2, x, dev) if B >= lib.phx_tc_min_rows() else None
            with _hill_cache(dev, x2, planes):
                _lib.check(lib.phx_prior_loss(_lib.ctx(dev), G, H, B, engine._ptr(packed), engine._ptr(x2),
                                              engine._ptr(pg), ctypes.c_float(2.0 / (B * G)), engine._ptr(gcot),
                                              engine._ptr(loss), engine._ptr(ws), ws.numel(), engine._stream_ptr(dev)),
                           "prior_loss")
        with engine._state.lock:
            engine._last_rhs[ws.data_ptr()] = engine._rhs_signature(net, packed, x2, B)
        ctx.net, ctx.x2, ctx.gcot, ctx.planes, ctx.dev = net, x2, gcot, planes, dev
        return loss.reshape(())

    @staticmethod
    def backward(ctx, grad_out):
        with _hill_cache(ctx.dev, ctx.x2, ctx.planes):
            _, flat = engine.rhs_vjp(ctx.net, ctx.x2, ctx.gcot, False, need_ybar=False, need_grads=True, flat=True)
        flat *= grad_out          # one pass over the flat [P] vector; the six gradients stay views of ONE buffer
        G = ctx.x2.shape[-1]
        grads = engine.split_flat_grads(flat, G, (flat.numel() - G) // (4 * G + 2))
        out = [None, None, None]
        for i, need in enumerate(ctx.needs_input_grad[3:]):
            # gene_multipliers do not enter prior_only_forward (odenet.py:93-98): the reference leaves that .grad untouched
            out.append(grads[i] if (need and i > 0) else None)
        return tuple(out)


def prior_loss(odenet, batch_for_prior, prior_grad):
    """``torch.mean((odenet.prior_only_forward(t, batch_for_prior) - prior_grad) ** 2)`` (train_insilico.py:134-135) as
    one fused, differentiable operation (gradients flow to the ODENet parameters; ``batch_for_prior`` and ``prior_grad``
    are constants, as in the reference)."""
    return _PriorLoss.apply(odenet, batch_for_prior, prior_grad, *engine.net_params(odenet))
