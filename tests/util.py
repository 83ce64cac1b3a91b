"""Shared helpers for the test-suite: golden fixture access + digest helpers (mirror tests/golden/make_golden.py)."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
N_SAMPLES = 32


def golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def sample_idx(numel, n=N_SAMPLES):
    out, x = [], 12345 + numel
    for _ in range(min(n, numel)):
        x = (1103515245 * x + 12345) % (2 ** 31)
        out.append(x % numel)
    return np.asarray(out, np.int64)


def tensor_digest(t):
    t = torch.as_tensor(t).detach().double().flatten().cpu()
    return np.asarray([t.sum().item(), (t * t).sum().item(), t.abs().sum().item()], np.float64)


def samples_of(t):
    t = torch.as_tensor(t).detach().flatten().cpu()
    return t[torch.from_numpy(sample_idx(t.numel()))].numpy()


def oliver_stat(parted=True):
    g = golden("speaker_stat_oliver")
    tag = "parted" if parted else "global"
    return {"mean": g[tag + "_mean"], "std": g[tag + "_std"], "scale_factor": float(g[tag + "_scale_factor"])}


def rel_err(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))
