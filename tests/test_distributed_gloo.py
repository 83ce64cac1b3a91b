"""world_size-2 gloo test (CPU) of the data-parallel host logic: equal clip shards + ONE flat all-reduce(sum) with the
1/W scale == the full-batch gradient of the per-sample-norm generator.  Rank gradients come from the CPU
oracle; identity activations (slope 1.0) keep the comparison free of LeakyReLU mask flips."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _grads(cfg, sd, audio, code, gt):
    from oracle import sdt_oracle as O
    sd = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    mel = O.mel_spectrogram(audio)
    pred = O.generator_forward(mel, 64, code, sd, cfg, True, "netG.")
    loss = torch.abs(pred - gt).mean()
    names = list(sd)
    return names, torch.autograd.grad(loss, [sd[k] for k in names])


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    from oracle import sdt_oracle as O
    from speechdrivestemplates_b200 import parallel
    cfg = O.make_cfg("voice2pose_sdt_bp", g_leaky=1.0)
    torch.manual_seed(0)
    sd = O.init_generator(cfg)
    g = torch.Generator().manual_seed(9)
    total = 2 * world
    audio = 0.1 * torch.randn(total, 68266, generator=g)
    code = 0.1 * torch.randn(total, 32, generator=g)
    gt = torch.randn(total, 64, 2, 121, generator=g)
    sl = parallel.shard_slice(total, world, rank)
    assert parallel.per_rank_batch(total + 1, world) == 2          # remainder dropped (trainer.py:75,78)
    names, gr = _grads(cfg, sd, audio[sl], code[sl], gt[sl])
    flat, offs = parallel.pack_flat(list(gr))
    parallel.allreduce_flat_(flat)
    flat *= parallel.grad_scale()
    if rank == 0:
        _n, full = _grads(cfg, sd, audio, code, gt)
        worst = 0.0
        for o, t in zip(offs, full):
            ref = t.reshape(-1).numpy()
            got = flat[o:o + t.numel()].numpy()
            worst = max(worst, float(np.abs(got - ref).max() / (np.sqrt((ref ** 2).mean()) + 1e-30)))
        np.save(os.path.join(out_dir, "worst.npy"), np.asarray([worst]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_flat_allreduce_equals_full_batch(tmp_path):
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    worst = float(np.load(os.path.join(str(tmp_path), "worst.npy"))[0])
    assert worst < 5e-3, worst          # fp32 noise floor of the early-layer weight gradients ~1e-3 of rms


def _worker_rows(rank, world, port, out_dir):
    """Sparse clip-code exchange == dense all-reduce; bucketed all-reduce == flat all-reduce (same sums, disjoint cover)."""
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from speechdrivestemplates_b200 import parallel
    g = torch.Generator().manual_seed(100 + rank)
    N, D, B = 50, 32, 6
    idx = torch.randint(0, 8, (B,), generator=g)                 # few distinct rows: duplicates inside and across ranks
    rows = torch.randn(B, D, generator=g)
    dense = torch.zeros(N, D).index_add_(0, idx, rows)
    dist.all_reduce(dense)
    rows_all, idx_all, works = parallel.gather_rows(rows, idx)
    assert works == [] and rows_all.shape == (world * B, D) and idx_all.shape == (world * B,)
    assert torch.equal(rows_all[rank * B:(rank + 1) * B], rows) and torch.equal(idx_all[rank * B:(rank + 1) * B], idx)
    sparse = torch.zeros(N, D).index_add_(0, idx_all, rows_all)
    ok_rows = bool(torch.allclose(sparse, dense, rtol=0, atol=1e-6))
    sizes = [("a.weight", 40), ("b.weight", 24), ("c.weight", 64), ("d.weight", 8), ("e.bias", 4)]
    total = sum(n for _, n in sizes) + 20                        # + a tail that stays off the wire (the dense code table)
    plan = parallel.bucket_plan(sizes, ["d.weight", "b.weight"], total, tail=140)
    assert plan == [("d.weight", 128, 140), ("b.weight", 40, 128), (None, 0, 40)], plan
    flat = torch.randn(total, generator=g)
    ref = flat.clone()
    dist.all_reduce(ref)
    for _m, lo, hi in plan:
        dist.all_reduce(flat[lo:hi])
    ok_buckets = bool(torch.equal(flat[:140], ref[:140])) and not bool(torch.equal(flat[140:], ref[140:]))
    if rank == 0:
        np.save(os.path.join(out_dir, "ok.npy"), np.asarray([ok_rows, ok_buckets]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_code_row_exchange_and_buckets(tmp_path):
    port = 31500 + os.getpid() % 2000
    mp.spawn(_worker_rows, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    ok = np.load(os.path.join(str(tmp_path), "ok.npy"))
    assert ok.all(), ok


def test_bucket_plan_rejects_forward_order():
    import pytest
    sys.path.insert(0, ROOT)
    from speechdrivestemplates_b200 import parallel
    with pytest.raises(ValueError):
        parallel.bucket_plan([("a", 8), ("b", 8)], ["a", "b"], 16)
