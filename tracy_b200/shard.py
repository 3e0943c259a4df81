"""Multi-GPU sharding of a batch of independent pairs.

Every (trace, window) pair is independent, so a batch is split into contiguous index ranges, one per rank (one
process per GPU); there is no data-path collective. The only exchanges are the ones the north star names: one
broadcast of the shared reference bytes before, one gather of the int32 scores after. Both go through
torch.distributed (NCCL over NVLink on the GPU box, gloo in the CPU tests)."""
import numpy as np


def partition(n, world, rank):
    """Contiguous [lo, hi) of ceil-balanced size for uniform shapes."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def partition_by_cost(len1, len2, world):
    """Ragged shapes: contiguous ranges balanced by sum(m_i * n_i) via a prefix sum. Returns world+1 boundaries."""
    cost = np.asarray(len1, np.int64) * np.asarray(len2, np.int64)
    total = int(cost.sum())
    pre = np.concatenate([[0], np.cumsum(cost)])
    bounds = [0]
    for r in range(1, world):
        target = total * r / world
        k = int(np.searchsorted(pre, target, side="left"))
        bounds.append(min(max(k, bounds[-1]), len(cost)))
    bounds.append(len(cost))
    return bounds


def broadcast_reference(ref_tensor, src=0, group=None):
    """One broadcast of the shared reference bytes (a uint8 tensor allocated with the same size on every rank)."""
    import torch.distributed as dist
    dist.broadcast(ref_tensor, src=src, group=group)
    return ref_tensor


def gather_scores(local_scores, counts, group=None):
    """All-gather of per-rank int32 score tensors of (possibly) unequal length -> one tensor in batch order.
    `counts` are the per-rank lengths (from partition()); ranks pad to max(counts) so a single collective suffices."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    mx = max(counts)
    buf = torch.zeros(mx, dtype=local_scores.dtype, device=local_scores.device)
    buf[: local_scores.numel()] = local_scores
    out = torch.empty(world * mx, dtype=local_scores.dtype, device=local_scores.device)
    dist.all_gather_into_tensor(out, buf, group=group)
    return torch.cat([out[r * mx: r * mx + counts[r]] for r in range(world)])


def broadcast_genome_and_index(ctx, text_tensor, src=0, group=None):
    """The broadcast the north star names, with its consumer: rank `src` holds the reference text (uint8 tensor on the rank's
    GPU; the other ranks pass a tensor of the same size), one NCCL broadcast over NVLink ships it, and every rank builds its
    own anchoring index from the device copy (17 ms per 64 Mbp -- cheaper than shipping the 17 B/char index itself).
    Returns the rank's KmerIndex."""
    import torch
    broadcast_reference(text_tensor, src=src, group=group)
    # dist.broadcast only enqueues (NCCL's stream on the GPU box); the index build copies the text on the context's own
    # non-blocking stream, which is ordered against neither NCCL's nor torch's: wait for the text to have landed first.
    if text_tensor.is_cuda:
        torch.cuda.synchronize(text_tensor.device)
    return ctx.build_index_device(text_tensor.data_ptr(), text_tensor.numel())
