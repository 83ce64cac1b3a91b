"""Execution engines of the hot path: they own the HBM workspaces (channels-last activations, statistics,
gradient buffers) and sequence the C-ABI kernels for forward and backward.  No autograd inside; the nn.Module
drop-ins (networks.py) and the fused train step (pipeline.py) sit on top.

Layer tables follow the reference: AudioEncoder generator.py:15-30, UNet_1D :53-68, decoder :98-104,
PoseSeqEncoder autoencoder.py:17-25.  HBM layout: 2-D maps (B,H,W,C), sequences (B,L,C); what is stored for a
2-D block is its RAW convolution output plus per-(b,c) [IN] or per-c [BN] (scale, shift); the next consumer's loader
normalises + activates on the fly.  The 1-D generator stack stores activated tensors (they are tiny).
"""
import torch

from . import ops
from .ops import ConvGeom

# (state-dict suffix, Cout, Cin, kh, kw, stride, pad) -- generator.py:15-30
ENC2D = [
    ("0.0", 64, 1, 3, 3, 1, 1), ("0.1", 64, 64, 4, 4, 2, 1),
    ("1.0", 128, 64, 3, 3, 1, 1), ("1.1", 128, 128, 4, 4, 2, 1),
    ("2.0", 256, 128, 3, 3, 1, 1), ("2.1", 256, 256, 4, 4, 2, 1),
    ("3.0", 256, 256, 3, 3, 1, 1), ("3.1", 256, 256, 6, 3, 1, 0),
]
ENC_PREFIX = "audio_encoder.specgram_encoder_2d."
UNET_E = ["e0", "e1", "e2", "e3", "e4", "e5", "e6"]
UNET_D = ["d5", "d4", "d3", "d2", "d1"]


class Arena:
    """Named persistent device buffers (allocated once through PyTorch's caching allocator)."""

    def __init__(self, device):
        self.device = device
        self.bufs = {}

    def get(self, name, shape, dtype=torch.float32, zero=False):
        shape = tuple(int(s) for s in shape)
        t = self.bufs.get(name)
        if t is None or tuple(t.shape) != shape or t.dtype != dtype:
            t = (torch.zeros if zero else torch.empty)(shape, device=self.device, dtype=dtype)
            self.bufs[name] = t
        return t

    def nbytes(self):
        return sum(t.numel() * t.element_size() for t in self.bufs.values())


class WeightPrep:
    """All GEMM weight operands of a network, refreshed by ONE kernel launch per step (sdt_weight_prep_batch).

    The parameters keep the reference layout (Cout,Cin,kh,kw); the kernels want (K,N) [FFMA] or K-major (N,K)
    [tcgen05] copies, per stride-parity class for the data gradient.  Pointers are static across steps, so the item
    table is built and uploaded once and rebuilt only if a parameter moved or the math mode changed.
    """

    def __init__(self, arena, math):
        self.arena = arena
        self.math = math
        self.key = None
        self.fwd = {}       # name -> (wt or None, wt_nk or None)
        self.dgrad = {}     # name -> [(cls, wt or None, wt_nk or None)]

    def ensure(self, layers, params, with_dgrad):
        """layers: [(name, weight key, geom, (H, W) of the layer input, needs_dgrad[, tc_ok])].

        tc_ok=False keeps a layer on the FFMA operands even in math mode 1 (the tcgen05 kernel stages one image's
        scale/shift per CTA, so layers whose loader transform spans a batch-wide statistic stay on the FFMA kernel)."""
        import numpy as np
        tc = self.math >= 1
        key = (tc, with_dgrad, tuple(tuple(l[6:]) for l in layers)) + tuple(params[l[1]].data_ptr() for l in layers)
        if key == self.key:
            return
        self.key = key
        A = self.arena
        dt = np.dtype([("w", "<u8"), ("out", "<u8"), ("i", "<i4", (12,))])
        items, max_elems = [], 1
        self.fwd, self.dgrad = {}, {}

        def add(w, out, g, mode, ky0, kx0, th, tw, cin_pad=0):
            nonlocal max_elems
            kstep = g.sw if mode in (1, 3) else 1          # forward operands take every tap; dgrad classes every s-th
            items.append((w.data_ptr(), out.data_ptr(), (g.cout, g.cin, g.kh, g.kw, mode, ky0, kx0, kstep, th, tw, cin_pad, 0)))
            max_elems = max(max_elems, th * tw * max(g.cin, cin_pad) * g.cout)

        for layer in layers:
            name, wkey, g, (H, W), needs_dgrad = layer[:5]
            tc_ok = layer[5] if len(layer) > 5 else True
            cin_pad = layer[6] if len(layer) > 6 else 0    # forward-only layers: input channels zero-padded (242 -> 256)
            dgrad_only = len(layer) > 7 and layer[7]       # zero-padded copy of a layer for its data gradient (TMA kernels: N % 64 == 0)
            w = params[wkey]
            if dgrad_only:
                pass
            elif tc and tc_ok and cin_pad and ops.tc_eligible(cin_pad, g.cout):
                assert not needs_dgrad
                nk = A.get("wt_fnk:" + name, (g.cout, g.kh * g.kw * cin_pad))
                add(w, nk, g, 2, 0, 0, g.kh, g.kw, cin_pad)
                self.fwd[name] = (None, nk)
            elif tc and tc_ok and ops.tc_eligible(g.cin, g.cout):
                nk = A.get("wt_fnk:" + name, (g.cout, g.k))
                add(w, nk, g, 2, 0, 0, g.kh, g.kw)
                self.fwd[name] = (None, nk)
            else:
                kn = A.get("wt_f:" + name, (g.k, g.cout))
                add(w, kn, g, 0, 0, 0, g.kh, g.kw)
                self.fwd[name] = (kn, None)
            if with_dgrad and needs_dgrad:
                lst = []
                dtc = tc and tc_ok and (ops.tc_eligible(g.cout, g.cin) or dgrad_only)
                for ci, cls in enumerate(g.dgrad_classes(H, W)):
                    kd = cls["th"] * cls["tw"] * g.cout
                    if dtc:
                        nk = A.get("wt_dnk%d:%s" % (ci, name), (g.cin, kd))
                        add(w, nk, g, 3, cls["ky0"], cls["kx0"], cls["th"], cls["tw"])
                        lst.append((cls, None, nk))
                    else:
                        kn = A.get("wt_d%d:%s" % (ci, name), (kd, g.cin))
                        add(w, kn, g, 1, cls["ky0"], cls["kx0"], cls["th"], cls["tw"])
                        lst.append((cls, kn, None))
                self.dgrad[name] = lst
        arr = np.zeros(len(items), dtype=dt)
        for i, (wp, op, ints) in enumerate(items):
            arr[i] = (wp, op, ints)
        self.table = torch.from_numpy(arr.view(np.uint8).copy()).to(A.device)
        self.n_items, self.max_elems = len(items), max_elems

    def run(self):
        import ctypes as C
        from ._lib import call
        call("sdt_weight_prep_batch", C.c_void_p(self.table.data_ptr()), self.n_items, self.max_elems, ops._stream())


def seq_lengths(num_frames):
    """Lengths of the UNet levels e1..e6 for an input of num_frames (k=4, s=2, p=1 halves with floor)."""
    ls = [num_frames]
    for _ in range(5):
        ls.append((ls[-1] + 2 - 4) // 2 + 1)
    return ls   # [L(e0/e1), L(e2), ..., L(e6)]


class GeneratorEngine:
    """SequenceGeneratorCNN (generator.py:87-117): forward and backward on caller-provided parameter tensors.

    params / grads: dicts keyed by the reference's state-dict names relative to the generator
    (e.g. 'audio_encoder.specgram_encoder_2d.0.0.conv.weight', 'unet.e0.conv.weight', 'decoder.4.bias').
    """

    def __init__(self, norm, leaky, code_dim, n_landmarks, device, math=None):
        if norm not in ("IN", "BN"):
            raise NotImplementedError(norm)                # building_blocks.py:27-28
        self.norm = norm
        self.slope = float(leaky) if isinstance(leaky, float) else (0.2 if leaky else 0.0)   # building_blocks.py:46
        self.code_dim = code_dim or 0
        self.kp2 = n_landmarks * 2
        self.device = device
        self.arena = Arena(device)
        self.enc_geoms = [ConvGeom.conv2d(ci, co, kh, kw, s, p) for (_n, co, ci, kh, kw, s, p) in ENC2D]
        self.shape_key = None
        self.fwd_id = 0
        self.math = ops.resolve_math(math)      # this engine's convolution math mode (every descriptor carries it)
        self.wprep = WeightPrep(self.arena, self.math)
        self.fuse_rownorm = __import__("os").environ.get("SDT_FUSE_ROWNORM", "1") != "0"     # 1-D "IN" blocks: row norm in the conv epilogue
        self.grad_marks = None                  # {layer name: event recorded behind that layer's weight gradient} (multi-GPU buckets)
        self.on_mark = None                     # callback(name, event), called when a mark has been recorded

    # ---- static layer tables -------------------------------------------------------------------
    def seq_layers(self):
        """[(name, geom, input_kind)] of the 1-D stack in forward order. input_kind: 'x0' | 'act:<name>' | 'up:<prev>+<skip>'."""
        c0 = 256 + self.code_dim
        out = [("unet.e0", ConvGeom.conv1d(c0, 256, 3, 1, 1), "x0"),
               ("unet.e1", ConvGeom.conv1d(256, 256, 3, 1, 1), "act:unet.e0")]
        for i in range(2, 7):
            out.append(("unet.e%d" % i, ConvGeom.conv1d(256, 256, 4, 2, 1), "act:unet.e%d" % (i - 1)))
        prev = "unet.e6"
        for j, n in enumerate(UNET_D):
            out.append(("unet." + n, ConvGeom.conv1d(256, 256, 3, 1, 1), "up:%s+unet.e%d" % (prev, 5 - j)))
            prev = "unet." + n
        for i in range(4):
            out.append(("decoder.%d" % i, ConvGeom.conv1d(256, 256, 3, 1, 1), "act:" + prev))
            prev = "decoder.%d" % i
        return out

    def param_shapes(self):
        """Reference parameter/buffer layout (name -> shape); BN affine + running stats only when norm == 'BN'."""
        shapes = {}

        def block(prefix, wshape):
            shapes[prefix + ".conv.weight"] = wshape
            if self.norm == "BN":
                c = wshape[0]
                shapes[prefix + ".norm.weight"] = (c,)
                shapes[prefix + ".norm.bias"] = (c,)

        for (n, co, ci, kh, kw, _s, _p) in ENC2D:
            block(ENC_PREFIX + n, (co, ci, kh, kw))
        for name, g, _k in self.seq_layers():
            block(name, (g.cout, g.cin, g.kw))
        shapes["decoder.4.weight"] = (self.kp2, 256, 1)
        shapes["decoder.4.bias"] = (self.kp2,)
        return shapes

    # ---- helpers ----------------------------------------------------------------------------------
    def _setup(self, B, T, F):
        key = (B, T, F)
        if self.shape_key == key:
            return
        self.shape_key = key
        self.B, self.T, self.F = B, T, F
        hw = [(80, T)]
        for g in self.enc_geoms:
            hw.append(g.out_hw(*hw[-1]))
        self.enc_hw = hw                               # hw[l] = input size of layer l; hw[l+1] = its output size
        ls = seq_lengths(F)
        self.seq_len = {"x0": F}          # e0,e1: F ; e2: ls[1]; ... e6: ls[5]
        self.seq_len["unet.e0"] = F
        self.seq_len["unet.e1"] = F
        for i in range(2, 7):
            self.seq_len["unet.e%d" % i] = ls[i - 1]
        for j, n in enumerate(UNET_D):
            self.seq_len["unet." + n] = self.seq_len["unet.e%d" % (5 - j)]
        for i in range(4):
            self.seq_len["decoder.%d" % i] = F

    def _groups(self):
        return self.B if self.norm == "IN" else 1

    def _bstride(self, c):
        return c if self.norm == "IN" else 0

    def _prep_weight(self, name, w, g):
        """-> (FFMA operand (K,N) or None, tensor-core operand (N,K) or None); filled by the batched prep launch."""
        return self.wprep.fwd[name]

    def _all_layers(self):
        """[(name, geom, input (H, W), needs_dgrad)] of every convolution of the generator."""
        out = []
        for l, (lname, *_r) in enumerate(ENC2D):
            out.append((ENC_PREFIX + lname, ENC_PREFIX + lname + ".conv.weight", self.enc_geoms[l], self.enc_hw[l], l > 0))  # mel needs no gradient
        for name, g, kind in self.seq_layers():
            if kind == "x0":
                lin = self.F
            elif kind.startswith("act:"):
                lin = self.seq_len[kind[4:]]
            else:                                   # up(prev)+skip: resampled to the skip's (== this layer's) length
                lin = self.seq_len[name]
            out.append((name, name + ".conv.weight", g, (1, lin), True))
        if self._x0_pad():
            g0 = self.seq_layers()[0][1]
            out.append(("unet.e0:dpad", "unet.e0.conv.weight_dpad", ConvGeom.conv1d(self._x0_pad(), g0.cout, g0.kw, g0.sw, g0.pw),
                        (1, self.F), True, True, 0, True))
        if self._head_pad():
            out.append(("decoder.4", "decoder.4.weight_pad", ConvGeom.conv1d(256, self._head_pad(), 1, 1, 0), (1, self.F), True))
        else:
            out.append(("decoder.4", "decoder.4.weight", ConvGeom.conv1d(256, self.kp2, 1, 1, 0), (1, self.F), True))
        return out

    def _head_pad(self):
        """The pose head is a 1x1 convolution to 2*K = 242 channels, which no tensor-core tile divides.  In the TMA math modes it runs
        as a 256 -> 256 layer on zero-padded copies (weight rows / bias entries / gradient channels 242..255 are zero, 0.3 MB) so that
        its forward, data gradient and weight gradient use the tcgen05 kernels instead of three ~50 us FFMA launches.  0 = no padding."""
        if self.math >= 2 and self.kp2 % 64 != 0:
            return -(-self.kp2 // 64) * 64
        return 0

    def _x0_pad(self):
        """The first UNet layer reads 256 encoder channels + the clip code (288 with a 32-d code): its data gradient is a GEMM
        with N = 288, which no 64-wide tensor-core tile divides (it ran 74 us on the FFMA kernel, on the critical path).  In the
        TMA math modes the gradient is computed with a zero-padded copy of the weight (input channels 288 -> 320) into a
        320-channel buffer whose tail is zero; the resize / code adjoint reads it as 256 + 64 channels.  0 = no padding."""
        cin = 256 + self.code_dim
        if self.math >= 2 and cin % 64 != 0:
            return -(-cin // 64) * 64
        return 0

    def _padded_head_params(self, params):
        """params + zero-padded copies of decoder.4.weight / .bias (refreshed from the parameters every step)."""
        npad, cpad = self._head_pad(), self._x0_pad()
        if not npad and not cpad:
            return params
        A = self.arena
        out = dict(params)
        if npad:
            wp = A.get("head_w_pad", (npad, 256, 1), zero=True)
            bp = A.get("head_b_pad", (npad,), zero=True)
            wp[:self.kp2].copy_(params["decoder.4.weight"])
            bp[:self.kp2].copy_(params["decoder.4.bias"])
            out["decoder.4.weight_pad"], out["decoder.4.bias_pad"] = wp, bp
        if cpad:
            w0 = params["unet.e0.conv.weight"]
            wp0 = A.get("x0_w_dpad", (w0.shape[0], cpad, w0.shape[2]), zero=True)
            wp0[:, :w0.shape[1]].copy_(w0)
            out["unet.e0.conv.weight_dpad"] = wp0
        return out

    # forward / backward / prepare are attached below (GeneratorEngine.forward = _gen_forward, ...): they are long enough to read
    # better as module-level functions.


def _gen_forward(self, mel, num_frames, code, params, training=True, buffers=None, upto=None, from_x0=None):
    """GeneratorEngine.forward: mel (B,80,T) f32, code (B,D) or None -> pred (B,F,2K), a view of an engine-owned buffer.
    buffers: BN running statistics dict (name -> tensor), updated in training mode, read in eval mode.
    upto='encoder': stop after the 2-D encoder + resize/concat and return the UNet input x0 (B, F, 256 + D) -- lets bench.py
    time the fused mel + encoder forward of the north-star roofline on its own.
    from_x0: the UNet input (B, F, 256 + D) computed by the caller (inference.StreamingGenerator runs the 2-D encoder in time
    tiles): only the 1-D stack runs here; the caller has set the engine up (``prepare``) and refreshed the weight operands."""
    if from_x0 is not None:
        return _gen_forward_seq(self, from_x0, num_frames, params, training, buffers)
    B, _, T = mel.shape
    self._setup(B, T, num_frames)
    A = self.arena
    slope = self.slope
    groups = self._groups()
    self.fwd_id += 1
    self._mel = mel
    self._params = params
    self.materialize = self.math >= 2
    # tensor-core math: every producer of a convolution operand stores TF32-rounded values (include/sdt_b200.h "out_tf32")
    tf32 = self.math >= 1
    # math mode 2 with InstanceNorm and an invertible activation: the 1 -> 64 first block runs as the single-pass special
    # case (csrc/first_layer.cu): no raw map, no separate normalisation pass, closed-form weight gradient
    self.fused_first = self.materialize and self.norm == "IN" and slope > 0.0
    params = self._padded_head_params(params)
    self._params = params
    self.wprep.ensure(self._all_layers(), params, with_dgrad=training)
    # the batched weight-operand refresh (0.12 ms) is not needed by the single-pass first block: run it beside that block on
    # the side stream and join before the first convolution that reads a prepared operand
    wg = getattr(self, "wg_stream", None)
    wprep_ready = None
    if wg is not None and self.fused_first:
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream())
        with torch.cuda.stream(wg):
            wg.wait_event(ev)
            self.wprep.run()
            wprep_ready = torch.cuda.Event()
            wprep_ready.record(wg)
    else:
        self.wprep.run()
    # ---- 2-D encoder: raw conv output + statistics; normalise/activate in the consumer's loader
    src, xf = mel.view(B, 80, T, 1), None
    for l, (lname, co, ci, kh, kw, s, p) in enumerate(ENC2D):
        g = self.enc_geoms[l]
        name = ENC_PREFIX + lname
        H, W = self.enc_hw[l]
        oh, ow = self.enc_hw[l + 1]
        sc = A.get("scale:" + name, (groups, co))
        sh = A.get("shift:" + name, (groups, co))
        if l == 0 and self.fused_first:
            act = A.get("act2d:" + name, (B, oh, ow, co))
            ops.first_layer_fwd(mel, params[name + ".conv.weight"], slope,
                                out=(act, sc, sh, A.get("mom:" + name, (B, 54), torch.float64)),
                                scratch=A.get("mom_partial:" + name, (B, ops.first_layer_units(H, W), 54), torch.float64), tf32=tf32)
            src, xf = act, None
            continue
        if wprep_ready is not None:
            torch.cuda.current_stream().wait_event(wprep_ready)
            wprep_ready = None
        wt, wt_nk = self._prep_weight(name, params[name + ".conv.weight"], g)
        raw = A.get("raw:" + name, (B, oh, ow, co))
        use_batch_stats = self.norm == "IN" or training
        d = ops.fwd_desc(g, src, wt, raw, B, H, W, xf, slope, per_image=True, wt_nk=wt_nk, math=self.math)
        if use_batch_stats:
            partial = A.get("partial:" + name, (ops.row_tiles(d), 2, co))
            d.stat_partial = partial.data_ptr()
            ops.conv_gemm(d)
            mean = A.get("mean:" + name, (groups, co))
            rstd = A.get("rstd:" + name, (groups, co))
            gamma = params.get(name + ".norm.weight")
            beta = params.get(name + ".norm.bias")
            running = None
            if self.norm == "BN":
                running = (buffers[name + ".norm.running_mean"], buffers[name + ".norm.running_var"],
                           buffers[name + ".norm.num_batches_tracked"])
            ops.norm_finalize(partial, groups, co, oh * ow * (B // groups), gamma, beta, running, out=(sc, sh, mean, rstd))
        else:
            ops.conv_gemm(d)
            ops.bn_eval_scale_shift(buffers[name + ".norm.running_mean"], buffers[name + ".norm.running_var"],
                                    params[name + ".norm.weight"], params[name + ".norm.bias"], out=(sc, sh))
        last_raw, last_xf = raw, (sc, sh, self._bstride(co))
        if self.materialize:
            # math mode 2: the TMA-fed tensor-core kernels take plain tensors, so the normalised + activated map is written
            # once (HBM is at ~10 % utilisation; the SM-side operand path is the bottleneck)
            act = A.get("act2d:" + name, (B, oh, ow, co))
            ops.scale_shift_act(raw, sc, sh, self._bstride(co), slope, out=act, tf32=tf32)
            src, xf = act, None
        else:
            src, xf = last_raw, last_xf
    # ---- bilinear resize to F frames + clip-code concat (generator.py:41-42,109-111)
    D = self.code_dim
    h7, w7 = self.enc_hw[8]
    x0 = A.get("x0", (B, num_frames, 256 + D))
    ops.enc_to_seq_fwd(last_raw, last_xf[0], last_xf[1], last_xf[2], slope, code if D > 0 else None, num_frames, out=x0, tf32=tf32)
    if upto == "encoder":
        return x0
    return _gen_forward_seq(self, x0, num_frames, params, training, buffers)


def _gen_prepare(self, B, T, num_frames, params, training=False):
    """Engine set-up + ONE weight-operand refresh for callers that sequence the layers themselves (time-tiled inference)
    -> params with the padded head copies."""
    self._setup(B, T, num_frames)
    self.fwd_id += 1
    self.materialize = self.math >= 2
    self.fused_first = False
    params = self._padded_head_params(params)
    self._params = params
    self.wprep.ensure(self._all_layers(), params, with_dgrad=training)
    self.wprep.run()
    return params


def _gen_forward_seq(self, x0, num_frames, params, training, buffers):
    """1-D stack (UNet_1D generator.py:53-84 + decoder :98-104) on the UNet input x0 (B, F, 256 + D)."""
    A, B, slope = self.arena, x0.shape[0], self.slope
    tf32 = self.math >= 1
    # ---- 1-D stack
    acts = {"x0": x0}
    for name, g, kind in self.seq_layers():
        L_out = self.seq_len[name]
        if kind == "x0":
            xin = x0
        elif kind.startswith("act:"):
            xin = acts[kind[4:]]
        else:
            prev, skip = kind[3:].split("+")
            xin = A.get("xin:" + name, (B, L_out, 256))
            ops.upsample_add_fwd(acts[prev], acts[skip], L_out, out=xin, tf32=tf32)
        L_in = xin.shape[1]
        acts["in:" + name] = xin
        wt, wt_nk = self._prep_weight(name, params[name + ".conv.weight"], g)
        raw = A.get("raw:" + name, (B, L_out, 256))
        d = ops.fwd_desc(g, xin, wt, raw, B, 1, L_in, wt_nk=wt_nk, math=self.math)
        act = A.get("act:" + name, (B, L_out, 256))
        if self.norm == "IN":
            rmean, rrstd = A.get("rmean:" + name, (B * L_out,)), A.get("rrstd:" + name, (B * L_out,))
            if self.fuse_rownorm and ops.conv_rownorm_ok(d):
                # channel LayerNorm + activation in the convolution's epilogue (cluster of four CTAs per row tile): one launch
                ops.conv_gemm(ops.set_rownorm(d, act, rmean, rrstd, slope, tf32))
            else:
                ops.conv_gemm(d)
                ops.rownorm_act_fwd(raw, slope, out=(act, rmean, rrstd), tf32=tf32)
        else:
            sc = A.get("scale:" + name, (1, 256))
            sh = A.get("shift:" + name, (1, 256))
            if training:
                partial = A.get("partial:" + name, (ops.row_tiles(d), 2, 256))
                d.stat_partial = partial.data_ptr()
                ops.conv_gemm(d)
                running = (buffers[name + ".norm.running_mean"], buffers[name + ".norm.running_var"],
                           buffers[name + ".norm.num_batches_tracked"])
                ops.norm_finalize(partial, 1, 256, B * L_out, params[name + ".norm.weight"], params[name + ".norm.bias"], running,
                                  out=(sc, sh, A.get("mean:" + name, (1, 256)), A.get("rstd:" + name, (1, 256))))
            else:
                ops.conv_gemm(d)
                ops.bn_eval_scale_shift(buffers[name + ".norm.running_mean"], buffers[name + ".norm.running_var"],
                                        params[name + ".norm.weight"], params[name + ".norm.bias"], out=(sc, sh))
            ops.scale_shift_act(raw, sc, sh, 0, slope, out=act, tf32=tf32)
        acts[name] = act
    self._acts = acts
    # ---- final 1x1 conv + bias (generator.py:103); channels-last output IS (B,F,2,K) (generator.py:116)
    pred = A.get("pred", (B, num_frames, self.kp2))
    npad = self._head_pad()
    if npad:
        gl = ConvGeom.conv1d(256, npad, 1, 1, 0)
        wt, wt_nk = self._prep_weight("decoder.4", params["decoder.4.weight_pad"], gl)
        pred_pad = A.get("pred_pad", (B, num_frames, npad))
        ops.conv_gemm(ops.fwd_desc(gl, acts["decoder.3"], wt, pred_pad, B, 1, num_frames, bias=params["decoder.4.bias_pad"], wt_nk=wt_nk, math=self.math))
        pred.copy_(pred_pad[..., :self.kp2])
        return pred
    gl = ConvGeom.conv1d(256, self.kp2, 1, 1, 0)
    wt, wt_nk = self._prep_weight("decoder.4", params["decoder.4.weight"], gl)
    ops.conv_gemm(ops.fwd_desc(gl, acts["decoder.3"], wt, pred, B, 1, num_frames, bias=params["decoder.4.bias"], wt_nk=wt_nk, math=self.math))
    return pred


_DIAG_SKIP_WGRAD = bool(__import__("os").environ.get("SDT_DIAG_SKIP_WGRAD"))


def _wgrad(self, g, x, dy, B, H, W, grad_out, xf=None, slope=1.0, post=None, name=None, key=None):
    """Weight gradient of one layer.  Nothing downstream in the backward pass depends on it, so when the engine has a
    `wg_stream` it is enqueued there (after an event marking dy ready) and overlaps the dgrad chain on the main stream.

    `name` in self.grad_marks: an event is recorded right behind this layer's gradient (weight gradients are produced in
    reverse layer order on one stream, so the event also covers every layer after it) -- the fused trainer starts the
    all-reduce of a gradient bucket on it (pipeline.Voice2PoseTrainer._start_buckets).

    `self.defer_reduce` (set by the fused trainers, whose gradient buffers are static): the split-K reduction of the layer is not
    launched here but collected and run as ONE batched launch per gradient bucket / at the join (_wgrad_flush)."""
    if _DIAG_SKIP_WGRAD:              # diagnostic only (wrong gradients): how much of the step do the weight gradients cost?
        return
    wg = getattr(self, "wg_stream", None)
    marks = getattr(self, "grad_marks", None)
    key = key or name
    if wg is not None:
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream())
        with torch.cuda.stream(wg):
            wg.wait_event(ev)
            _wgrad_now(self, g, x, dy, B, H, W, grad_out, xf, slope, post, key)
            if marks is not None and name in marks:
                _wgrad_flush(self)
                marks[name] = torch.cuda.Event()
                marks[name].record(wg)
                if self.on_mark is not None:
                    self.on_mark(name, marks[name])
        self._wg_pending = True
        return
    _wgrad_now(self, g, x, dy, B, H, W, grad_out, xf, slope, post, key)
    if marks is not None and name in marks:
        _wgrad_flush(self)
        marks[name] = torch.cuda.Event()
        marks[name].record(torch.cuda.current_stream())
        if self.on_mark is not None:
            self.on_mark(name, marks[name])


def _wgrad_join(self):
    """Main stream waits for the weight gradients enqueued on wg_stream (the deferred reductions are flushed first)."""
    wg = getattr(self, "wg_stream", None)
    if wg is not None and getattr(self, "_wg_pending", False):
        with torch.cuda.stream(wg):
            _wgrad_flush(self)
        ev = torch.cuda.Event()
        ev.record(wg)
        torch.cuda.current_stream().wait_event(ev)
        self._wg_pending = False
    else:
        _wgrad_flush(self)


def _wgrad_flush(self):
    """One batched launch for the split-K reductions collected since the last flush (on the current stream = the stream the
    weight gradients ran on).  The item tables are cached by their (static) pointers: built during the eager warm-up steps,
    replayed from the CUDA graph afterwards."""
    pending = getattr(self, "_reduce_pending", None)
    if not pending:
        return
    self._reduce_pending = []
    items = [p[0] for p in pending]
    key = tuple((wp.data_ptr(), gr.data_ptr(), sp, n, c, t, acc) for (wp, gr, sp, n, c, t, acc) in items)
    tables = self.__dict__.setdefault("_reduce_tables", {})
    if key not in tables:
        if torch.cuda.is_current_stream_capturing():
            raise RuntimeError("weight-gradient reduction table missing during graph capture (buffers moved since the warm-up step)")
        tables[key] = ops.reduce_item_table(items, self.arena.device)
    table, max_ctas, max_T = tables[key]
    ops.wgrad_reduce_batch(table, len(items), max_ctas, max_T)
    for p in pending:
        if p[1] is not None:
            p[1]()


def _wgrad_now(self, g, x, dy, B, H, W, grad_out, xf=None, slope=1.0, post=None, key=None):
    oh, ow = g.out_hw(H, W)
    splits = ops.wgrad_splits(g, B, oh, ow, math=self.math)
    need = splits * g.cout * g.k
    defer = getattr(self, "defer_reduce", False) and key is not None and g.cin % 32 == 0 and g.kh * g.kw <= 256
    if defer:
        ws = self.arena.get("wgrad_ws:" + key, (need,))            # per layer: it lives until the bucket's batched reduction
    else:
        ws = self.arena.get("wgrad_ws", (max(need, getattr(self, "_ws_elems", 0)),))
        self._ws_elems = ws.numel()
    ops.conv_wgrad(ops.wgrad_desc(g, x, dy, ws, B, H, W, splits, xf, slope, math=self.math))
    if defer:
        self.__dict__.setdefault("_reduce_pending", []).append(((ws, grad_out, splits, g.cout, g.cin, g.kh * g.kw, False), post))
        return
    ops.wgrad_reduce(ws, splits, g, grad_out)
    if post is not None:
        post()


def _dgrad(self, name, g, dy, w, dx, B, H, W, accumulate=False):
    # stride-parity classes with their weight operands, prepared at the start of the step by the batched launch
    ops.conv_gemm_multi([ops.dgrad_desc(g, cls, dy, wt, dx, B, H, W, accumulate, wt_nk=wt_nk, math=self.math)
                         for cls, wt, wt_nk in self.wprep.dgrad[name]])


def _gen_backward(self, g_pred, grads, g_code=None):
    """g_pred (B,F,2K) -> parameter gradients written into grads[name] (reference layout); g_code (B,D) filled.

    NORM='IN' (SDT configs): channel-LayerNorm / InstanceNorm2d backward.  NORM='BN' (voice2pose_s2g): BatchNorm backward
    with batch statistics, which also writes grads[<block>.norm.weight/.bias].
    """
    A, B, F, slope = self.arena, self.B, self.F, self.slope
    bn = self.norm == "BN"
    tf32 = self.math >= 1
    params, acts = self._params, self._acts
    # ---- final conv
    ops.colsum(g_pred, grads["decoder.4.bias"])
    g_act = {}
    g_act["decoder.3"] = A.get("g_act:decoder.3", (B, F, 256))
    npad = self._head_pad()
    if npad:
        gl = ConvGeom.conv1d(256, npad, 1, 1, 0)
        g_pad = A.get("g_pred_pad", (B, F, npad), zero=True)             # channels 242..255 stay zero
        g_pad[..., :self.kp2].copy_(g_pred.view(B, F, self.kp2))
        gw_pad = A.get("head_gw_pad", (npad, 256, 1))
        dst = grads["decoder.4.weight"]
        _wgrad(self, gl, acts["decoder.3"], g_pad, B, 1, F, gw_pad, post=lambda: dst.copy_(gw_pad[:self.kp2]), key="decoder.4")
        _dgrad(self, "decoder.4", gl, g_pad, params["decoder.4.weight_pad"], g_act["decoder.3"], B, 1, F)
    else:
        gl = ConvGeom.conv1d(256, self.kp2, 1, 1, 0)
        _wgrad(self, gl, acts["decoder.3"], g_pred, B, 1, F, grads["decoder.4.weight"], key="decoder.4")
        _dgrad(self, "decoder.4", gl, g_pred, params["decoder.4.weight"], g_act["decoder.3"], B, 1, F)
    # ---- 1-D stack in reverse
    layers = self.seq_layers()
    g_x0 = None
    for name, g, kind in reversed(layers):
        L_out = self.seq_len[name]
        raw = A.get("raw:" + name, (B, L_out, 256))
        if bn:
            tpi = ops.bwd_tiles(L_out, B)
            scratch = (A.get("nb_partial:" + name, (B * tpi, 2, 256)), A.get("nb_m1:" + name, (1, 256)), A.get("nb_m2:" + name, (1, 256)))
            g_raw = ops.norm_backward(g_act[name], raw, A.get("mean:" + name, (1, 256)), A.get("rstd:" + name, (1, 256)), 1, slope,
                                      params[name + ".norm.weight"], params[name + ".norm.bias"],
                                      grads[name + ".norm.weight"], grads[name + ".norm.bias"], scratch=scratch, tf32=tf32)
        else:
            g_raw = A.get("g_raw:" + name, (B, L_out, 256))          # per layer: its wgrad may still be in flight
            ops.rownorm_act_bwd(g_act[name], raw, A.get("rmean:" + name, (B * L_out,)), A.get("rrstd:" + name, (B * L_out,)), slope,
                                out=g_raw, tf32=tf32)
        xin = acts["in:" + name]
        L_in = xin.shape[1]
        _wgrad(self, g, xin, g_raw, B, 1, L_in, grads[name + ".conv.weight"], name=name)
        w = params[name + ".conv.weight"]
        if kind == "x0" and self._x0_pad():
            cpad = self._x0_pad()
            g_x0 = A.get("g_x0_pad", (B, L_in, cpad))
            _dgrad(self, name + ":dpad", ConvGeom.conv1d(cpad, g.cout, g.kw, g.sw, g.pw), g_raw, params[name + ".conv.weight_dpad"],
                   g_x0, B, 1, L_in)
        elif kind == "x0":
            g_x0 = A.get("g_x0", tuple(xin.shape))
            _dgrad(self, name, g, g_raw, w, g_x0, B, 1, L_in)
        elif kind.startswith("act:"):
            src = kind[4:]
            if src in g_act:          # skip connection already deposited its gradient there
                _dgrad(self, name, g, g_raw, w, g_act[src], B, 1, L_in, accumulate=True)
            else:
                g_act[src] = A.get("g_act:" + src, (B, L_in, 256))
                _dgrad(self, name, g, g_raw, w, g_act[src], B, 1, L_in)
        else:
            prev, skip = kind[3:].split("+")
            g_xin = A.get("g_xin:" + name, (B, L_in, 256))
            _dgrad(self, name, g, g_raw, w, g_xin, B, 1, L_in)
            g_act[skip] = g_xin                                       # d(up(prev)+skip)/d skip = identity
            g_act[prev] = A.get("g_act:" + prev, (B, self.seq_len[prev], 256))
            ops.upsample_bwd(g_xin, self.seq_len[prev], out=g_act[prev])
    # ---- resize/concat adjoint
    h7, w7 = self.enc_hw[8]
    g_enc = A.get("g_enc:7", (B, h7, w7, 256))
    D = self.code_dim
    if self._x0_pad():
        # the padded gradient buffer reads as 256 encoder channels + (D + zero tail) code channels
        Dp = self._x0_pad() - 256
        gc_pad = A.get("g_code_pad", (B, Dp))
        ops.enc_to_seq_bwd(g_x0, h7, w7, 256, Dp, g_act=g_enc, g_code=gc_pad)
        if D > 0 and g_code is not None:
            g_code.copy_(gc_pad[:, :D])
    else:
        ops.enc_to_seq_bwd(g_x0, h7, w7, 256, D, g_act=g_enc, g_code=g_code if D > 0 else None)
    # ---- 2-D encoder in reverse
    groups = self._groups()
    for l in range(7, -1, -1):
        lname, co, ci, kh, kw, s, p = ENC2D[l]
        g = self.enc_geoms[l]
        name = ENC_PREFIX + lname
        H, W = self.enc_hw[l]
        oh, ow = self.enc_hw[l + 1]
        if l == 0 and self.fused_first:
            ops.first_layer_bwd(g_enc, A.get("act2d:" + name, (B, oh, ow, co)), self._mel, params[name + ".conv.weight"],
                                A.get("mom:" + name, (B, 54), torch.float64), A.get("scale:" + name, (groups, co)),
                                A.get("shift:" + name, (groups, co)), slope, grads[name + ".conv.weight"],
                                scratch=A.get("fl_partial:" + name, (B, ops.first_layer_units(H, W), 11, co)))
            break
        raw = A.get("raw:" + name, (B, oh, ow, co))
        tpi = ops.bwd_partial_tiles(oh * ow, B, co)
        scratch = (A.get("nb_partial:" + name, (B * tpi, 2, co)), A.get("nb_m1:" + name, (groups, co)), A.get("nb_m2:" + name, (groups, co)))
        if bn:
            ops.norm_backward(g_enc, raw, A.get("mean:" + name, (groups, co)), A.get("rstd:" + name, (groups, co)), groups, slope,
                              params[name + ".norm.weight"], params[name + ".norm.bias"],
                              grads[name + ".norm.weight"], grads[name + ".norm.bias"], scratch=scratch, tf32=tf32)
        else:
            ops.norm_backward(g_enc, raw, A.get("mean:" + name, (groups, co)), A.get("rstd:" + name, (groups, co)), groups, slope,
                              scratch=scratch, tf32=tf32)
        if l == 0:
            src, xf = self._mel.view(B, 80, self.T, 1), None
        else:
            pname = ENC_PREFIX + ENC2D[l - 1][0]
            pc = ENC2D[l - 1][1]
            if self.materialize:
                src, xf = A.get("act2d:" + pname, (B, H, W, pc)), None
            else:
                src = A.get("raw:" + pname, (B, H, W, pc))
                xf = (A.get("scale:" + pname, (groups, pc)), A.get("shift:" + pname, (groups, pc)), self._bstride(pc))
        _wgrad(self, g, src, g_enc, B, H, W, grads[name + ".conv.weight"], xf, slope, name=name)
        if l > 0:
            g_prev = A.get("g_enc:%d" % (l - 1), (B, H, W, ci))
            _dgrad(self, name, g, g_enc, params[name + ".conv.weight"], g_prev, B, H, W)
            g_enc = g_prev
    _wgrad_join(self)


GeneratorEngine.forward = _gen_forward
GeneratorEngine.prepare = _gen_prepare
GeneratorEngine.backward = _gen_backward


class PoseEncoderEngine:
    """PoseSeqEncoder (autoencoder.py:8-35), forward only: BatchNorm in train mode (batch statistics, running-stat
    update) or eval mode.  Raw conv outputs + per-channel (scale, shift); activations never materialised."""

    def __init__(self, n_landmarks, code_dim, leaky, device, math=None):
        self.kp2 = n_landmarks * 2
        self.code2 = code_dim * 2
        self.slope = 0.2 if leaky else 0.0
        self.math = ops.resolve_math(math)
        self.arena = Arena(device)
        self.geoms = ([ConvGeom.conv1d(self.kp2, 256, 3, 1, 1), ConvGeom.conv1d(256, 256, 3, 1, 1)]
                      + [ConvGeom.conv1d(256, 256, 4, 2, 1)] * 4 + [ConvGeom.conv1d(256, self.code2, 4, 2, 1)])
        self.wprep = WeightPrep(self.arena, self.math)

    def param_shapes(self):
        shapes = {}
        for i, g in enumerate(self.geoms):
            shapes["blocks.%d.conv.weight" % i] = (g.cout, g.cin, g.kw)
            shapes["blocks.%d.norm.weight" % i] = (g.cout,)
            shapes["blocks.%d.norm.bias" % i] = (g.cout,)
        return shapes

    def forward(self, poses, params, buffers, training, tag=""):
        """poses (B,F,2K) channels-last (== the reference's (B,F,2,K)) -> (mu, logvar) engine-owned (B,D) buffers."""
        A = self.arena
        B, L = poses.shape[0], poses.shape[1]
        src, xf = poses.view(B, 1, L, self.kp2), None
        # TMA-fed tensor-core path (math modes 2, 3): the 242 input channels are zero-padded to 256 (K-block = 32 channels)
        # and every block's activation is materialised (2 MB), because TMA delivers plain tensors only
        tma = self.math >= 2
        cpad = -(-self.kp2 // 32) * 32 if tma else 0
        if tag in ("", "/pred"):        # the two FGD passes of a step share one weight refresh
            self.wprep.ensure([("blocks.%d" % i, "blocks.%d.conv.weight" % i, g, (1, 1), False, True, cpad if i == 0 else 0)
                               for i, g in enumerate(self.geoms)], params, False)
            self.wprep.run()
        if tma and cpad != self.kp2:
            padded = A.get("poses_padded" + tag, (B, 1, L, cpad), zero=True)      # the pad columns stay zero
            padded[..., :self.kp2].copy_(src)
            src = padded
        for i, g in enumerate(self.geoms):
            name = "blocks.%d" % i
            lo = g.out_hw(1, L)[1]
            wt, wt_nk = self.wprep.fwd[name]
            if i == 0 and src.shape[-1] != g.cin:
                g = ConvGeom.conv1d(src.shape[-1], g.cout, g.kw, g.sw, g.pw)
            raw = A.get("raw%s:%s" % (tag, name), (B, 1, lo, g.cout))
            # BN statistics span the batch (row tiles may straddle clips); scale/shift are per channel (bstride 0)
            d = ops.fwd_desc(g, src, wt, raw, B, 1, L, xf, self.slope, wt_nk=wt_nk, math=self.math)
            sc = A.get("scale%s:%s" % (tag, name), (1, g.cout))
            sh = A.get("shift%s:%s" % (tag, name), (1, g.cout))
            if training:
                partial = A.get("partial%s:%s" % (tag, name), (ops.row_tiles(d), 2, g.cout))
                d.stat_partial = partial.data_ptr()
                ops.conv_gemm(d)
                running = (buffers[name + ".norm.running_mean"], buffers[name + ".norm.running_var"],
                           buffers[name + ".norm.num_batches_tracked"])
                ops.norm_finalize(partial, 1, g.cout, B * lo, params[name + ".norm.weight"], params[name + ".norm.bias"], running,
                                  out=(sc, sh, A.get("mean%s:%s" % (tag, name), (1, g.cout)), A.get("rstd%s:%s" % (tag, name), (1, g.cout))))
            else:
                ops.conv_gemm(d)
                ops.bn_eval_scale_shift(buffers[name + ".norm.running_mean"], buffers[name + ".norm.running_var"],
                                        params[name + ".norm.weight"], params[name + ".norm.bias"], out=(sc, sh))
            if tma and i + 1 < len(self.geoms):
                act = A.get("act%s:%s" % (tag, name), (B, 1, lo, g.cout))
                ops.scale_shift_act(raw.view(B, lo, g.cout), sc, sh, 0, self.slope, out=act.view(B, lo, g.cout), tf32=True)
                src, xf, L = act, None, lo
            else:
                src, xf, L = raw, (sc, sh, 0), lo
        mu = A.get("mu" + tag, (B, self.code2 // 2))
        logvar = A.get("logvar" + tag, (B, self.code2 // 2))
        ops.pose_head_fwd(src.view(B, L, self.code2), xf[0], xf[1], self.slope, mu, logvar)
        return mu, logvar
