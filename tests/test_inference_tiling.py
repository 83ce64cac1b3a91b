"""Host arithmetic of the time-tiled long-audio forward (speechdrivestemplates_b200/inference.py), on CPU: for any utterance length and
chunk size, at every encoder level and from every admissible base level,

* the columns the tiles OWN partition the level's full width (every column's statistics and stored copy come from exactly one tile);
* the owned columns lie inside what the tile computes locally, and inside the part of it that is EXACT, i.e. whose receptive field
  never touches the zero padding a tile edge fakes at an interior cut (that is what the 64-column halo is for)."""
import random

import pytest

from speechdrivestemplates_b200 import inference
from speechdrivestemplates_b200.engine import ENC2D


def _out_w(w, k, s, p):
    return (w + 2 * p - k) // s + 1


def _levels(T):
    widths, strides, w, st = [], [], T, 1
    for (_n, _co, _ci, _kh, kw, s, p) in ENC2D:
        w = _out_w(w, kw, s, p)
        st *= s
        widths.append(w)
        strides.append(st)
    return widths, strides


@pytest.mark.parametrize("seed", range(6))
def test_owned_columns_partition_every_level_and_stay_exact(seed):
    rng = random.Random(seed)
    for _ in range(60):
        T = rng.randint(300, 9000)
        chunk_cols = rng.choice([64, 128, 256, 424, 856, 1704]) // 8 * 8
        if chunk_cols + 2 * inference.HALO >= T:
            continue
        widths, strides = _levels(T)
        tiles = inference.plan_tiles(T, chunk_cols)
        assert tiles[0][0] == 0 and tiles[-1][1] == T and all(a[1] == b[0] for a, b in zip(tiles, tiles[1:]))
        assert all(t[0] % 8 == 0 for t in tiles) and all(t[1] % 8 == 0 for t in tiles[:-1])
        for l in range(8):
            # ownership partitions [0, W_l)
            edge = 0
            for t in tiles:
                o0, o1, _off = inference.owned_window(t, T, strides[l], widths[l])
                assert o0 == edge and o1 >= o0
                edge = o1
            assert edge == widths[l]
            # from every base level below l (-1 = the mel): local extent and exact region cover the owned columns
            for base in range(-1, l):
                for t in tiles:
                    a0, a1, hl, hr = t
                    c0, c1 = inference.base_window(t, T, 1 if base < 0 else strides[base], T if base < 0 else widths[base])
                    w = c1 - c0
                    lo, hi = 0, w                       # exact local columns at the current level
                    cut_l, cut_r = hl > 0, a1 + hr < T   # interior cuts fake a zero padding there; true ends (also a halo clipped at T) are the real padding
                    for j in range(base + 1, l + 1):
                        _n, _co, _ci, _kh, kw, s, p = ENC2D[j]
                        nw = _out_w(w, kw, s, p)
                        nlo = -(-(lo + p) // s) if cut_l else 0
                        nhi = (hi - kw + p) // s + 1 if cut_r else nw
                        lo, hi, w = max(nlo, 0), min(nhi, nw), nw
                    o0, o1, off = inference.owned_window(t, T, strides[l], widths[l])
                    if o1 > o0:
                        assert 0 <= o0 - off and o1 - off <= w, (T, chunk_cols, l, base, t)
                        assert lo <= o0 - off and o1 - off <= hi, (T, chunk_cols, l, base, t, lo, hi, o0 - off, o1 - off)
                    # the tile's local grid is the global grid shifted by `off` (strided layers stay phase-aligned)
                    assert (a0 - hl) % strides[l] == 0


def test_one_tile_when_the_chunk_covers_the_utterance():
    assert inference.plan_tiles(500, 1000) == [(0, 500, 0, 0)]
    tiles = inference.plan_tiles(1000, 808)            # the 192-column rest is below a quarter chunk: one tile takes it all
    assert tiles == [(0, 1000, 0, 0)]
    tiles = inference.plan_tiles(1000, 400)
    assert tiles == [(0, 400, 0, 64), (400, 800, 64, 64), (800, 1000, 64, 0)]
