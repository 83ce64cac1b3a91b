"""Long-audio inference (demo path): wav -> pose stream through the mel front end and the generator in eval mode.

Reference: ``Trainer.demo`` (core/pipelines/trainer.py:459-484) loops ``Voice2Pose.demo_step`` (voice2pose.py:386-410), which runs
``Voice2PoseModel.forward(return_loss=False)`` on the WHOLE utterance in one fully-convolutional forward; ``DATASET.MAX_DEMO_LENGTH``
(24 s by default) bounds it because every intermediate map of the 2-D encoder is alive at once (4.75 GB for 10 minutes here).

``StreamingGenerator`` gives the same result with a fraction of the memory: the 2-D encoder -- 97 % of the
FLOPs and nearly all of the memory -- runs in time tiles with receptive-field halos, and the InstanceNorm2d statistics, which span
the whole utterance, are made exact by sweeping the layers: sweep l recomputes the layers below l on each tile with their
already-final statistics and accumulates layer l's sums over the tile's OWNED columns (deterministic fixed-order partials, one
``sdt_norm_finalize`` per layer).  The first block (1 -> 64 channels, the largest map) needs no sweep: its statistics follow in
closed form from the mel (csrc/first_layer.cu), and a tile's activated map is one pass.  A sweep starts from the nearest STORED level below l: the mel, or the full-length raw map of a
layer in ``store_layers`` (default 3 and 5: maps at 1/4 and 1/8 of the mel resolution, 0.26 + 0.13 MB per second of audio), which
cuts the recomputation from 4.4x the one-shot FLOPs (``store_layers=()``: memory independent of the length apart from the mel and
the final 1/8-resolution map) to 1.8x.  The last sweep writes the encoder's final map; the resize to ``num_frames`` and the 1-D
UNet + decoder then run once over the full length (9 MB per layer for 10 minutes).  ``chunk_frames=0`` keeps the one-shot forward.
"""
import torch

from . import _lib, ops, pipeline
from .engine import ENC2D, ENC_PREFIX
from .networks import SequenceGeneratorCNN
from .ops import ConvGeom

MEL_PER_FRAME = 16000.0 / 15.0 / 160.0          # mel columns per video frame (hop 160 at 16 kHz, 15 fps)
HALO = 64                                       # mel columns; > the encoder's receptive radius (42), multiple of its stride 8


def plan_tiles(T, chunk_cols, halo=HALO):
    """Time tiles of a T-column mel: [(a0, a1, hl, hr)] -- owned mel columns [a0, a1), loaded columns [a0 - hl, a1 + hr).  Every interior
    boundary is a multiple of chunk_cols (itself a multiple of the encoder's total stride 8); the last tile takes a short rest."""
    tiles, a0 = [], 0
    while a0 < T:
        a1 = min(T, a0 + chunk_cols)
        if T - a1 < chunk_cols // 4:          # do not leave a sliver: the last tile takes the rest
            a1 = T
        tiles.append((a0, a1, halo if a0 > 0 else 0, halo if a1 < T else 0))
        a0 = a1
    return tiles


def base_window(tile, T, stride_base, width_base):
    """Columns [c0, c1) of a stored level (cumulative stride stride_base, full width width_base; the mel: stride 1, width T) that a
    tile loads: the level's image of the tile's loaded mel columns."""
    a0, a1, hl, hr = tile
    c0 = (a0 - hl) // stride_base
    c1 = width_base if a1 + hr >= T else (a1 + hr) // stride_base
    return c0, c1


def owned_window(tile, T, stride_l, width_l):
    """(o0, o1, off): the tile OWNS output columns [o0, o1) of a level with cumulative stride stride_l and full width width_l (its
    statistics and its stored copy come from these only); off = global column of the tile's local column 0 at that level."""
    a0, a1, hl, _hr = tile
    o0 = min(a0 // stride_l, width_l)
    o1 = width_l if a1 >= T else min(a1 // stride_l, width_l)
    return o0, o1, (a0 - hl) // stride_l


class StreamingGenerator:
    """cfg: a Voice2Pose config (generator with NORM='IN' for the tiled path).  ``netG`` may be passed in (e.g. a trained
    ``Voice2PoseModel.netG``); otherwise a freshly initialised generator is built under the caller's torch seed."""

    def __init__(self, cfg, device, conv_math=None, chunk_frames=0, netG=None, mel=None, store_layers=(3, 5), graph=False):
        self.cfg = cfg
        self.device = torch.device(device)
        own = netG is None
        self.netG = (SequenceGeneratorCNN(cfg).to(self.device).eval() if own else netG)        # a caller's generator keeps its mode / device
        if own or self.netG.conv_math != (None if conv_math is None else int(conv_math)):
            self.netG.set_conv_math(conv_math)
        self.mel = (mel if mel is not None else pipeline.MelSpectrogram()).to(self.device)
        self.math = ops.resolve_math(conv_math)
        self.chunk_frames = int(chunk_frames)
        self.store_layers = tuple(store_layers)
        # graph=True: the time-tiled forward of a given utterance length is captured into a CUDA graph on first use and replayed
        # afterwards (its ~1,100 small launches are host-bound when issued one by one); a new length costs one extra eager pass
        self.use_graph = bool(graph)
        self._graph_cache = {}
        self.last_chunks = 1
        self.launches_per_call = 0
        self._bufs = {}

    # ---- public calls ---------------------------------------------------------------------------------
    @torch.no_grad()
    def __call__(self, audio_host, num_frames, code):
        """audio_host (1, L) f32 host tensor (pinned recommended), code (1, D) or None -> (1, num_frames, 2, K) host tensor."""
        a = audio_host.to(self.device, non_blocking=True)
        c = code.to(self.device) if code is not None else None
        return self.forward_device(a, num_frames, c).cpu()

    @torch.no_grad()
    def forward_device(self, audio, num_frames, code):
        if self.netG.training:
            raise RuntimeError("StreamingGenerator runs the generator in eval mode (call .eval() on the model first)")
        n0 = _lib.launch_count
        mel = self.mel(audio)
        T = mel.shape[-1]
        chunk_cols = int(round(self.chunk_frames * MEL_PER_FRAME / 8.0)) * 8 if self.chunk_frames > 0 else 0
        if 0 < chunk_cols < HALO:
            chunk_cols = HALO                       # a tile never reaches further back than its left neighbour
        if chunk_cols <= 0 or chunk_cols + 2 * HALO >= T or self.netG.norm_kind != "IN":
            self.last_chunks = 1
            out = self.netG(mel, num_frames, code)
        elif self.use_graph:
            out = self._tiled_graphed(audio, mel, num_frames, code, chunk_cols)
        else:
            out = self._tiled(mel, num_frames, code, chunk_cols)
        self.launches_per_call = max(_lib.launch_count - n0, getattr(self, "_graph_launches", 0))
        return out

    def _tiled_graphed(self, audio, mel, num_frames, code, chunk_cols):
        key = (tuple(audio.shape), int(num_frames), int(chunk_cols), code is not None, self.store_layers)
        ent = self._graph_cache.get(key)
        if ent is None:
            a = audio.clone()
            c = code.clone() if code is not None else None
            mel_buf = torch.empty_like(mel)

            def run():
                return self._tiled(self.mel(a, out=mel_buf), num_frames, c, chunk_cols)
            n0 = _lib.launch_count
            run()                                   # eager pass: buffers, weight-operand tables
            self._graph_launches = _lib.launch_count - n0
            torch.cuda.synchronize(self.device)
            g = torch.cuda.CUDAGraph()
            side = torch.cuda.Stream(device=self.device)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                with torch.cuda.graph(g, stream=side):
                    out = run()
            torch.cuda.current_stream().wait_stream(side)
            if len(self._graph_cache) >= 4:         # a handful of utterance lengths at most
                self._graph_cache.pop(next(iter(self._graph_cache)))
            ent = self._graph_cache[key] = (g, a, c, out)
        g, a, c, out = ent
        a.copy_(audio)
        if c is not None:
            c.copy_(code)
        g.replay()
        return out.clone()

    # ---- tiled 2-D encoder ----------------------------------------------------------------------------
    def _buf(self, name, shape, dtype=torch.float32):
        """A view of a named flat device buffer that only ever grows (tiles differ in width by their halos)."""
        n = 1
        for d in shape:
            n *= int(d)
        t = self._bufs.get(name)
        if t is None or t.numel() < n or t.dtype != dtype:
            t = torch.empty(max(n, 1), device=self.device, dtype=dtype)
            self._bufs[name] = t
        return t[:n].view(tuple(int(d) for d in shape))

    def _tiled(self, mel, num_frames, code, chunk_cols):
        eng = self.netG.engine()
        B, _, T = mel.shape
        slope = eng.slope
        tf32 = self.math >= 1
        params = {n: p.detach() for n, p in self.netG.named_parameters()}
        params = eng.prepare(B, T, num_frames, params)           # one weight-operand refresh for the whole call
        geoms = eng.enc_geoms
        hw = eng.enc_hw                                           # hw[l] = input size of layer l, hw[l + 1] = its output size
        stride, s_acc = [], 1
        for g in geoms:
            s_acc *= g.sw
            stride.append(s_acc)                                  # cumulative stride of layer l's OUTPUT in mel columns
        tiles = plan_tiles(T, chunk_cols)
        self.last_chunks = len(tiles)
        stored = {}                               # level -> full-length RAW map (B, H, W, C); level -1 = the mel
        stats = []                                # per layer: (scale, shift), each (B, C)
        keep = (set(self.store_layers) | {7}) - {0}
        rows_per_part = 64
        # first block (1 -> 64 channels, the largest map): its InstanceNorm statistics follow in closed form from the input (54 tap
        # moments of the whole mel, csrc/first_layer.cu), so it needs no sweep, and a tile's activated map is ONE pass from the mel tile
        w0 = params[ENC_PREFIX + ENC2D[0][0] + ".conv.weight"]
        sc0, sh0, _m = ops.first_layer_stats(mel, w0, out=(self._buf("scale0", (B, 64)), self._buf("shift0", (B, 64)),
                                                           self._buf("mom0", (B, 54), torch.float64)),
                                             scratch=self._buf("mom_partial0", (B, ops.first_layer_units(80, T), 54), torch.float64))
        stats.append((sc0, sh0))
        for l in range(1, 8):
            co = ENC2D[l][1]
            oh, ow_full = hw[l + 1]
            base = max([s for s in stored if s < l], default=-1)
            parts = [-(-(oh * (-(-(t[1] - t[0]) // stride[l]) + 1)) // rows_per_part) for t in tiles]
            partial = self._buf("partial%d" % l, (B, sum(parts), 2, co))
            if l in keep:
                stored[l] = self._buf("stored%d" % l, (B, oh, ow_full, co))
            part_off = 0
            for ti, (a0, a1, hl, hr) in enumerate(tiles):
                # ---- the tile of the base level (mel columns, or the stored raw map of `base` normalised + activated)
                if base < 0:
                    c0, c1 = base_window(tiles[ti], T, 1, T)
                    H, W = 80, c1 - c0
                    mt = self._buf("tile_mel", (B, 80, W))
                    mt.copy_(mel[:, :, c0:c1])
                    src = ops.first_layer_act(mt, w0, sc0, sh0, slope, out=self._buf("act0", (B, 80, W, 64)), tf32=tf32)
                    first = 1                           # layers first .. l run as convolutions on the tile
                else:
                    c0, c1 = base_window(tiles[ti], T, stride[base], hw[base + 1][1])
                    H, W = hw[base + 1][0], c1 - c0
                    src = self._buf("tile_base%d" % base, (B, H, W, ENC2D[base][1]))
                    src.copy_(stored[base][:, :, c0:c1])
                    ops.scale_shift_act(src, stats[base][0], stats[base][1], ENC2D[base][1], slope, out=src, tf32=tf32)
                    first = base + 1
                # ---- layers first .. l on the tile
                for j in range(first, l + 1):
                    g = geoms[j]
                    name = ENC_PREFIX + ENC2D[j][0]
                    oh_j, ow_j = g.out_hw(H, W)
                    raw = self._buf("raw%d" % j, (B, oh_j, ow_j, g.cout))
                    wt, wt_nk = eng.wprep.fwd[name]
                    ops.conv_gemm(ops.fwd_desc(g, src, wt, raw, B, H, W, per_image=True, wt_nk=wt_nk, math=self.math))
                    if j < l:
                        act = self._buf("act%d" % j, (B, oh_j, ow_j, g.cout))
                        ops.scale_shift_act(raw, stats[j][0], stats[j][1], g.cout, slope, out=act, tf32=tf32)
                        src, H, W = act, oh_j, ow_j
                # ---- layer l's statistics over the OWNED output columns (global [o0, o1)) + the stored copy
                o0, o1, off = owned_window(tiles[ti], T, stride[l], ow_full)
                ops.chan_stats(raw, o0 - off, o1 - off, rows_per_part, partial[:, part_off:part_off + parts[ti]])
                part_off += parts[ti]
                if l in keep and o1 > o0:
                    stored[l][:, :, o0:o1].copy_(raw[:, :, o0 - off:o1 - off])
            sc, sh = self._buf("scale%d" % l, (B, co)), self._buf("shift%d" % l, (B, co))
            ops.norm_finalize(partial.view(-1, 2, co), B, co, oh * ow_full, None, None, None,
                              out=(sc, sh, self._buf("mean%d" % l, (B, co)), self._buf("rstd%d" % l, (B, co))))
            stats.append((sc, sh))
        # resize to num_frames + code, then the 1-D stack over the full length (9 MB per layer for 10 minutes)
        D = eng.code_dim
        x0 = self._buf("x0", (B, num_frames, 256 + D))
        ops.enc_to_seq_fwd(stored[7], stats[7][0], stats[7][1], 256, slope, code if D > 0 else None, num_frames, out=x0, tf32=tf32)
        pred = eng.forward(None, num_frames, code, params, False, self.netG._buffers_dict(), from_x0=x0)
        return pred.view(B, num_frames, 2, -1).clone()
