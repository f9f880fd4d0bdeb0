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
    """Combine the parameter gradients (and optional extra scalars such as the loss) of all ranks in ONE collective.

    ``average=True`` (default) divides the sum by the world size: the reference's losses are ``torch.mean`` over the
    LOCAL samples (train_insilico.py:132,135), so with the samples of a step sharded over the ranks the averaged gradient
    is the single-GPU gradient (exactly when the shards are equal, up to the shard-size weighting otherwise) and the
    learning rate keeps its meaning as N grows.  A caller that normalises its losses by the GLOBAL batch size itself
    (tools/train_epoch.py) passes ``average=False`` to get the plain sum."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return extra
    flat = flatten_grads(net)
    if extra is not None:
        flat = torch.cat([flat, extra.reshape(-1).to(flat.dtype)])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    if average:
        flat /= dist.get_world_size(group)
    n_extra = 0 if extra is None else extra.numel()
    unflatten_grads_(net, flat[:flat.numel() - n_extra] if n_extra else flat)
    return flat[flat.numel() - n_extra:].view_as(extra) if n_extra else None


def broadcast_parameters(net, src=0, group=None):
    """Replicate rank `src`'s six weight tensors on every rank (start of training / after loading a checkpoint)."""
    if not dist.is_available() or not dist.is_initialized():
        return
    with torch.no_grad():
        for p in engine.net_params(net):
            dist.broadcast(p, src=src, group=group)
    engine.invalidate(net)   # the collective writes the storages without bumping the version counters
