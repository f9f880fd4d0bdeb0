"""Multi-GPU plumbing for the hot path: one process per GPU, the independent odeint problems
of a training step (train_insilico.py:128-130) are sharded across ranks, the six weight tensors are replicated, and
the ONLY exchange is one sum-allreduce of the flat parameter-gradient vector per optimiser step (NCCL over
NVLink 5 / NVSwitch on the B200 box, gloo in the CPU tests).  The reference has no distributed code at all."""
import torch
import torch.distributed as dist

from . import engine


def shard_range(n_items, rank, world_size):
    """Contiguous, balanced slice [lo, hi) of n_items independent samples for this rank."""
    base, rem = divmod(n_items, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def flatten_grads(net):
    """The six .grad tensors concatenated in the reference parameter order (zeros where a grad is missing)."""
    parts = []
    for p in engine.net_params(net):
        g = p.grad if p.grad is not None else torch.zeros_like(p)
        parts.append(g.reshape(-1))
    return torch.cat(parts)


def flat_grad_view(net):
    """The six ``.grad`` tensors as ONE flat ``[P]`` tensor WITHOUT a copy, or None.  The adjoint kernels write the
    parameter cotangents into one flat buffer (reference order) and hand autograd six views of it; when a parameter has
    no gradient yet autograd keeps that view as ``.grad``, so after ``backward()`` the six gradients normally still sit
    back to back in one storage -- which is what lets the collective below run in place."""
    params = engine.net_params(net)
    g0 = params[0].grad
    if g0 is None:
        return None
    st, off = g0.untyped_storage(), g0.storage_offset()
    for p in params:
        g = p.grad
        if (g is None or g.dtype != g0.dtype or not g.is_contiguous() or g.storage_offset() != off
                or g.untyped_storage().data_ptr() != st.data_ptr()):
            return None
        off += g.numel()
    return torch.empty(0, dtype=g0.dtype, device=g0.device).set_(st, g0.storage_offset(), (off - g0.storage_offset(),))


def adopt_flat_grads_(net, flat):
    """Make the six ``.grad`` tensors views of ``flat`` (no copy)."""
    o = 0
    for p in engine.net_params(net):
        n = p.numel()
        p.grad = flat[o:o + n].view_as(p)
        o += n


def unflatten_grads_(net, flat):
    o = 0
    for p in engine.net_params(net):
        n = p.numel()
        if p.grad is None:
            p.grad = flat[o:o + n].view_as(p).clone()
        else:
            p.grad.copy_(flat[o:o + n].view_as(p))
        o += n


def allreduce_grads(net, group=None, average=True, extra=None):
    """Combine the parameter gradients (and optional extra scalars such as the loss) of all ranks.

    ``average=True`` (default) divides the sum by the world size: the reference's losses are ``torch.mean`` over the
    LOCAL samples (train_insilico.py:132,135), so with the samples of a step sharded over the ranks the averaged gradient
    is the single-GPU gradient (exactly when the shards are equal, up to the shard-size weighting otherwise) and the
    learning rate keeps its meaning as N grows.  A caller that normalises its losses by the GLOBAL batch size itself
    (tools/train_epoch.py) passes ``average=False`` to get the plain sum.

    One collective over the flat ``[P]`` gradient, IN PLACE when the six gradients are views of one buffer
    (``flat_grad_view``: the normal case after a backward through this package -- no gather / scatter passes over the
    35.8 MB vector); otherwise they are gathered once and ``.grad`` becomes views of the reduced buffer (no copy back).
    ``extra`` rides in a second, tiny collective in the first case and at the tail of the vector in the second."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return extra
    world = dist.get_world_size(group)
    flat = flat_grad_view(net)
    if flat is not None:
        work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group, async_op=extra is not None)
        out = None
        if extra is not None:
            out = extra.reshape(-1).to(flat.dtype).clone()
            dist.all_reduce(out, op=dist.ReduceOp.SUM, group=group)
            work.wait()
            out = out.view_as(extra)
        if average:
            flat /= world
            if out is not None:
                out /= world
        return out
    flat = flatten_grads(net)
    n = flat.numel()
    if extra is not None:
        flat = torch.cat([flat, extra.reshape(-1).to(flat.dtype)])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    if average:
        flat /= world
    adopt_flat_grads_(net, flat[:n])
    return flat[n:].view_as(extra) if extra is not None else None


def broadcast_parameters(net, src=0, group=None):
    """Replicate rank `src`'s six weight tensors on every rank (start of training / after loading a checkpoint)."""
    if not dist.is_available() or not dist.is_initialized():
        return
    with torch.no_grad():
        for p in engine.net_params(net):
            dist.broadcast(p, src=src, group=group)
    engine.invalidate(net)   # the collective writes the storages without bumping the version counters


# ---- exact global error norm for a batched dopri5 solve sharded over ranks (SURVEY 8e caveat) -------------------------------
_SUM_HOOK_T = None
_sum_hooks = {}


class _DeviceDoubles:
    """A raw device pointer to n float64 values as a CUDA-array-interface object (torch.as_tensor wraps it, no copy)."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (ptr, False), "version": 2}


def enable_global_norm(device=None, group=None):
    """A single batched ``odeint`` call uses ONE error norm over all its rows (torchdiffeq/_impl/misc.py:10-11).  When the
    rows of such a call are split over ranks, call this once per rank (same number of rows on every rank): every norm
    evaluation of the streaming dopri5 forward solve then all-reduces its partial sums (3 doubles) over `group`, so all
    ranks take identical accept / reject decisions -- the same step sequence as the unsharded call, up to summation
    order.  Without it each rank steps by the norm of its own shard.  ``disable_global_norm`` switches it off."""
    import ctypes
    global _SUM_HOOK_T
    from . import _lib
    if not dist.is_available() or not dist.is_initialized():
        return
    dev = torch.cuda.current_device() if device is None else torch.device(device).index
    if _SUM_HOOK_T is None:
        _SUM_HOOK_T = ctypes.CFUNCTYPE(None, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p)

    def hook(ptr, n, stream, user):
        t = torch.as_tensor(_DeviceDoubles(ptr, n), device=torch.device("cuda", dev))
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)

    cb = _SUM_HOOK_T(hook)
    _sum_hooks[dev] = cb          # keep the trampoline alive
    _lib.check(_lib.load().phx_ctx_set_global_norm(_lib.ctx(dev), ctypes.cast(cb, ctypes.c_void_p), None,
                                                   dist.get_world_size(group)), "set_global_norm")


def disable_global_norm(device=None):
    from . import _lib
    dev = torch.cuda.current_device() if device is None else torch.device(device).index
    _lib.check(_lib.load().phx_ctx_set_global_norm(_lib.ctx(dev), None, None, 1), "set_global_norm")
    _sum_hooks.pop(dev, None)
