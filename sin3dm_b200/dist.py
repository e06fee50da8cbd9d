"""Multi-GPU sampling: independent samples shard across ranks.

One process per GPU (torchrun).  The only collectives are ONE broadcast of the flattened checkpoint at
start-up (NCCL over NVLink/NVSwitch) and, optionally, ONE all-gather of the finished latents.  Nothing is
exchanged inside the step loop: GroupNorm is per sample and the per-step noise is keyed by
(seed, global sample index, step), so a sample's trajectory is bit-identical at 1, 2, 4 or 8 GPUs.
The reference has no live distributed path (src/utils/dist_util.py:19-42 only sets CUDA_VISIBLE_DEVICES).
"""
import torch
import torch.distributed as dist


def shard_range(n_samples, world_size, rank):
    """Contiguous block partition: -> (first global sample index, count) for this rank."""
    base, extra = divmod(n_samples, world_size)
    count = base + (1 if rank < extra else 0)
    start = rank * base + min(rank, extra)
    return start, count


def broadcast_parameters(module, src=0):
    """One collective for the whole checkpoint: flatten -> broadcast -> scatter back (≈28 MB for the default UNet)."""
    params = [p for p in module.state_dict().values() if torch.is_tensor(p)]
    if not params:
        return 0
    flat = torch.cat([p.detach().reshape(-1).float() for p in params])
    dist.broadcast(flat, src=src)
    off = 0
    with torch.no_grad():
        for p in params:
            n = p.numel()
            p.copy_(flat[off:off + n].view_as(p))
            off += n
    return flat.numel()


def all_reduce_gradients(flat_grad, average=True):
    """Data-parallel training: ONE all-reduce of the flat gradient buffer per step (``FusedAdamWEMA.grad``, ~28 MB for the default
    UNet: the backward writes every gradient into that one buffer, so there is nothing to bucket).  NCCL over NVLink / NVSwitch on
    the GPU box, gloo in the CPU tests.  The reference trains on one device (src/utils/dist_util.py:19-42); this is the exchange
    step its commented-out ``sync_params`` / DDP lineage (guided-diffusion) would have had."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return flat_grad
    dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM)
    if average:
        flat_grad.div_(dist.get_world_size())
    return flat_grad


def sample_sharded(sample_fn, n_samples, sample_shape, batch_size=None, gather=True, device=None):
    """Runs ``sample_fn(shape, sample_base)`` for this rank's block of the ``n_samples`` global samples.

    ``sample_fn`` is typically ``lambda shape, base: diffusion.p_sample_loop(model, shape, model_kwargs=...,
    seed=seed, sample_base=base)``.  Returns the local samples, or with ``gather`` all samples in global order on
    every rank (one all_gather of equal-sized, zero-padded blocks).
    """
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    if device is None and torch.cuda.is_available():
        device = torch.device("cuda", torch.cuda.current_device())      # a rank that owns no sample still joins the gather
    start, count = shard_range(n_samples, world, rank)
    bs = batch_size or max(count, 1)
    outs = []
    for off in range(0, count, bs):
        n = min(bs, count - off)
        outs.append(sample_fn([n, *sample_shape], start + off))
    if outs:
        local = torch.cat(outs)
    else:
        local = torch.empty(0, *sample_shape, device=device)
    if not gather or world == 1:
        return local
    cap = -(-n_samples // world)
    pad = torch.zeros(cap, *sample_shape, device=local.device, dtype=local.dtype)
    pad[:count] = local
    blocks = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(blocks, pad)
    return torch.cat([blocks[r][: shard_range(n_samples, world, r)[1]] for r in range(world)])
