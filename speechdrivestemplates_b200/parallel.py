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


def bucket_plan(named_sizes, marks, total, tail=None):
    """Cut the flat gradient buffer (parameters in ``named_sizes`` order, [(name, numel)]) into buckets in the order a backward
    pass completes them.  ``marks`` are parameter names in backward order (last layer's side first): bucket i spans from
    marks[i] up to the start of the previous bucket (the first one up to ``tail``, default ``total``); a final unmarked bucket
    covers what is left at the front.  Returns [(mark or None, start, end)] -- disjoint, covering [0, tail)."""
    offs, off = {}, 0
    for name, n in named_sizes:
        offs[name] = off
        off += n
    end = total if tail is None else tail
    out = []
    for m in marks:
        start = offs[m]
        if not 0 <= start < end:
            raise ValueError("bucket marks must walk the parameter list backwards: %s" % (marks,))
        out.append((m, start, end))
        end = start
    if end > 0:
        out.append((None, 0, end))
    return out


def gather_rows(rows, idx, group=None, out=None, async_op=False):
    """The clip-code gradient of a step touches only the B rows of this rank's clips.  Instead of all-reducing the dense
    (N, D) table, every rank contributes its (B, D) rows + (B) indices: returns (rows_all (W*B, D), idx_all (W*B), works).
    Scattering rows_all by idx_all (duplicates accumulate) gives the SUM over ranks of the dense gradients."""
    world = dist.get_world_size(group)
    rows_all, idx_all = out if out is not None else (rows.new_empty((world * rows.shape[0],) + tuple(rows.shape[1:])),
                                                     idx.new_empty(world * idx.shape[0]))
    w1 = dist.all_gather_into_tensor(rows_all, rows.contiguous(), group=group, async_op=async_op)
    w2 = dist.all_gather_into_tensor(idx_all, idx.contiguous(), group=group, async_op=async_op)
    return rows_all, idx_all, [w for w in (w1, w2) if w is not None]
