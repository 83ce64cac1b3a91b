"""Data-parallel plumbing of the train step: clip sharding arithmetic and the single flat gradient all-reduce
(SURVEY §8e, C3/C7).  One process per GPU; ``torch.distributed`` (NCCL on GPUs, gloo in the CPU tests)."""
import torch
import torch.distributed as dist


def per_rank_batch(total_batch, world_size):
    """``BATCH_SIZE // world`` with the remainder dropped (trainer.py:75,78)."""
    return total_batch // world_size


def shard_slice(total_batch, world_size, rank):
    """Clips of the global batch owned by ``rank`` (equal contiguous shards, like DistributedSampler + drop_last)."""
    b = per_rank_batch(total_batch, world_size)
    return slice(rank * b, (rank + 1) * b)


def pack_flat(tensors, out=None):
    """Concatenate tensors into one flat fp32 buffer (allocated when ``out`` is None). Returns (flat, offsets)."""
    n = sum(t.numel() for t in tensors)
    if out is None:
        out = torch.zeros(n, dtype=torch.float32, device=tensors[0].device)
    offs, o = [], 0
    for t in tensors:
        out[o:o + t.numel()].copy_(t.reshape(-1))
        offs.append(o)
        o += t.numel()
    return out, offs


def allreduce_flat_(flat, group=None):
    """ONE all-reduce(sum) over the flat gradient buffer per step; the 1/W average is folded into Adam's grad_scale."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    return flat


def grad_scale(group=None):
    if dist.is_available() and dist.is_initialized():
        return 1.0 / dist.get_world_size(group)
    return 1.0
