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


class PeerExchange:
    """The step's gradient exchange over peer memory (SDT_COMM=p2p): the flat gradient buffer lives in a symmetric allocation
    (``torch.distributed._symmetric_memory``: every rank can address every rank's copy, and on NVSwitch systems the NVLS multicast
    address of all copies), ``sdt_p2p_allreduce`` (csrc/p2p.cu) reduces it in place between two cross-GPU barriers.  torch supplies
    allocation, rendezvous and the barrier (plumbing); the data path is the library's kernel -- no NCCL call in the step.

    ``flat``: the symmetric fp32 buffer (use it as the trainer's gradient buffer); ``allreduce(scal)`` sums ``flat`` over the ranks
    in place and, if ``scal`` (a local f64 vector) is given, replaces it by its sum over the ranks."""

    def __init__(self, numel, device, group, scal_n=16, use_multicast=None):
        import os
        import torch.distributed._symmetric_memory as symm
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        if self.world > 16:
            raise RuntimeError("PeerExchange: at most 16 ranks (one NVSwitch node)")
        numel = int(numel) + ((-int(numel)) % 4)
        self.flat = symm.empty(numel, dtype=torch.float32, device=device)
        self.flat.zero_()
        self._h = symm.rendezvous(self.flat, group)
        self.scal_src = symm.empty(int(scal_n), dtype=torch.float64, device=device)
        self.scal_src.zero_()
        self._hs = symm.rendezvous(self.scal_src, group)
        self.scal_n = int(scal_n)

        def addrs(h, t):
            base = [int(p) for p in h.buffer_ptrs]
            off = int(t.data_ptr()) - base[self.rank]            # the tensor's offset inside the symmetric allocation
            return [b + off for b in base], off
        self._ptrs, off = addrs(self._h, self.flat)
        self._sptrs, _ = addrs(self._hs, self.scal_src)
        if use_multicast is None:           # measured: NVLS multimem wins from 3 ranks up (8 GPUs: 3.12 vs 3.17 ms), peer loads at 2 (3.07 vs 3.12)
            env = os.environ.get("SDT_P2P_MULTICAST")
            use_multicast = (env != "0") if env is not None else self.world >= 3
        mc = int(self._h.multicast_ptr or 0) if use_multicast else 0            # 0 where the system has no NVLS multicast
        self.multicast = mc + off if mc else 0
        import ctypes as C
        self._c_ptrs = (C.c_uint64 * self.world)(*self._ptrs)
        self._c_sptrs = (C.c_uint64 * self.world)(*self._sptrs)
        torch.cuda.synchronize(device)
        dist.barrier(group=group)

    def allreduce(self, scal=None):
        import ctypes as C
        from . import ops
        from ._lib import call
        if scal is not None:
            self.scal_src[:scal.numel()].copy_(scal)
        self._h.barrier(channel=0)                    # every rank's gradients (and scalars) are final and visible
        call("sdt_p2p_allreduce", self._c_ptrs, C.c_uint64(self.multicast), self.rank, self.world, self.flat.numel(),
             self._c_sptrs, C.c_void_p(scal.data_ptr()) if scal is not None else None, scal.numel() if scal is not None else 0, ops._stream())
        self._h.barrier(channel=1)                    # every shard has been written everywhere (and every scalar block read)
