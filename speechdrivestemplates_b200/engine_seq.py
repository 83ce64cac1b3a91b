"""1-D sequence engines with BatchNorm: the pose VAE (core/networks/poses_reconstruction/autoencoder.py) and the
motion discriminator (core/networks/keypoints_generation/discriminator.py), forward AND backward.

Both are stacks of ConvNormRelu('1d', norm='BN') blocks (building_blocks.py:31-46) on short sequences (L <= 64), so
activations are materialised (they are tiny) and every block is: implicit-GEMM conv (statistics in the epilogue) ->
norm_finalize (+ running-stat update) -> scale/shift/LeakyReLU; backward = fused BN/activation backward
(sdt_norm_bwd_*, which also yields dgamma/dbeta) -> wgrad -> dgrad.  Layers may be run several times per step under
different `tag`s (the discriminator sees real / fake / fake.detach(), voice2pose.py:191-193).
"""
import torch

from . import ops
from .engine import Arena, WeightPrep
from .ops import ConvGeom


class SeqBlock:
    """One Conv1d [+ BatchNorm1d + (Leaky)ReLU] block with materialised output."""

    def __init__(self, owner, name, geom, norm, slope, wkey=None):
        self.o, self.name, self.g, self.norm, self.slope = owner, name, geom, norm, slope
        self.wkey = wkey or (name + ".conv.weight" if norm else name + ".weight")
        self.saved = {}

    def forward(self, xin, B, L_in, params, buffers, training, tag=""):
        A, g, name = self.o.arena, self.g, self.name
        L_out = g.out_hw(1, L_in)[1]
        wt, wt_nk = self.o.wprep.fwd[name]
        raw = A.get("raw%s:%s" % (tag, name), (B, L_out, g.cout))
        bias = params.get(name + ".bias") if self.norm is None else None
        d = ops.fwd_desc(g, xin, wt, raw, B, 1, L_in, bias=bias, wt_nk=wt_nk, math=self.o.math)
        if self.norm == "BN":
            sc = A.get("scale%s:%s" % (tag, name), (1, g.cout))
            sh = A.get("shift%s:%s" % (tag, name), (1, g.cout))
            gamma, beta = params[name + ".norm.weight"], params[name + ".norm.bias"]
            if training:
                partial = A.get("partial%s:%s" % (tag, name), (ops.row_tiles(d), 2, g.cout))
                d.stat_partial = partial.data_ptr()
                ops.conv_gemm(d)
                running = (buffers[name + ".norm.running_mean"], buffers[name + ".norm.running_var"],
                           buffers[name + ".norm.num_batches_tracked"])
                ops.norm_finalize(partial, 1, g.cout, B * L_out, gamma, beta, running,
                                  out=(sc, sh, A.get("mean%s:%s" % (tag, name), (1, g.cout)), A.get("rstd%s:%s" % (tag, name), (1, g.cout))))
            else:
                ops.conv_gemm(d)
                ops.bn_eval_scale_shift(buffers[name + ".norm.running_mean"], buffers[name + ".norm.running_var"], gamma, beta,
                                        out=(sc, sh))
            act = A.get("act%s:%s" % (tag, name), (B, L_out, g.cout))
            ops.scale_shift_act(raw, sc, sh, 0, self.slope, out=act, tf32=self.o.math >= 1)
        else:
            ops.conv_gemm(d)
            act = raw
        self.saved[tag] = (xin, L_in, L_out)
        return act

    def backward(self, g_act, B, params, grads, tag="", dx=None, accumulate_dx=False, accumulate_params=False, need_dx=True):
        """g_act (B,L_out,Cout) is consumed (overwritten in place by the BN backward). Returns dx (or None)."""
        A, g, name = self.o.arena, self.g, self.name
        xin, L_in, L_out = self.saved[tag]
        if self.norm == "BN":
            raw = A.get("raw%s:%s" % (tag, name), (B, L_out, g.cout))
            tpi = ops.bwd_tiles(L_out, B)
            scratch = (A.get("nb_partial:" + name, (B * tpi, 2, g.cout)), A.get("nb_m1:" + name, (1, g.cout)), A.get("nb_m2:" + name, (1, g.cout)))
            ops.norm_backward(g_act, raw, A.get("mean%s:%s" % (tag, name), (1, g.cout)), A.get("rstd%s:%s" % (tag, name), (1, g.cout)), 1,
                              self.slope, params[name + ".norm.weight"], params[name + ".norm.bias"],
                              grads[name + ".norm.weight"], grads[name + ".norm.bias"], accumulate_params, scratch=scratch,
                              tf32=self.o.math >= 1)
        elif name + ".bias" in grads:
            ops.colsum(g_act, grads[name + ".bias"], accumulate_params)
        g_raw = g_act
        oh, ow = g.out_hw(1, L_in)
        splits = ops.wgrad_splits(g, B, oh, ow, math=self.o.math)
        need = splits * g.cout * g.k
        ws = A.get("wgrad_ws", (max(need, getattr(self.o, "_ws_elems", 0)),))
        self.o._ws_elems = ws.numel()
        ops.conv_wgrad(ops.wgrad_desc(g, xin, g_raw, ws, B, 1, L_in, splits, math=self.o.math))
        ops.wgrad_reduce(ws, splits, g, grads[self.wkey], accumulate_params)
        if not need_dx:
            return None
        if dx is None:
            dx = A.get("dx%s:%s" % (tag, name), (B, L_in, g.cin))
        for cls, wt, wt_nk in self.o.wprep.dgrad[name]:
            ops.conv_gemm(ops.dgrad_desc(g, cls, g_raw, wt, dx, B, 1, L_in, accumulate_dx, wt_nk=wt_nk, math=self.o.math))
        return dx


class _SeqEngine:
    def __init__(self, device, math=None):
        self.arena = Arena(device)
        self.device = device
        self.math = ops.resolve_math(math)       # this engine's convolution math mode (every descriptor carries it)
        self.wprep = WeightPrep(self.arena, self.math)
        self.blocks = []

    def _prep(self, params, lengths, with_dgrad=True):
        layers = [(b.name, b.wkey, b.g, (1, lengths[b.name]), True) for b in self.blocks]
        self.wprep.ensure(layers, params, with_dgrad)
        self.wprep.run()


class AutoencoderEngine(_SeqEngine):
    """Autoencoder (autoencoder.py:71-92): PoseSeqEncoder -> reparameterisation -> PoseSeqDecoder; parameter names are
    relative to the Autoencoder module ('encoder.blocks.0.conv.weight', 'decoder.d5.norm.bias', 'decoder.blocks.4.bias')."""

    def __init__(self, n_landmarks, code_dim, leaky, device, math=None):
        super().__init__(device, math)
        self.kp2, self.D = n_landmarks * 2, code_dim
        self.slope = 0.2 if leaky else 0.0
        s = self.slope
        eg = ([ConvGeom.conv1d(self.kp2, 256, 3, 1, 1), ConvGeom.conv1d(256, 256, 3, 1, 1)]
              + [ConvGeom.conv1d(256, 256, 4, 2, 1)] * 4 + [ConvGeom.conv1d(256, 2 * code_dim, 4, 2, 1)])
        self.enc = [SeqBlock(self, "encoder.blocks.%d" % i, g, "BN", s) for i, g in enumerate(eg)]
        self.dec_up = [SeqBlock(self, "decoder.d5", ConvGeom.conv1d(code_dim, 256, 3, 1, 1), "BN", s)] + \
                      [SeqBlock(self, "decoder." + n, ConvGeom.conv1d(256, 256, 3, 1, 1), "BN", s) for n in ("d4", "d3", "d2", "d1")]
        self.dec_tail = [SeqBlock(self, "decoder.blocks.%d" % i, ConvGeom.conv1d(256, 256, 3, 1, 1), "BN", s) for i in range(4)]
        self.dec_out = SeqBlock(self, "decoder.blocks.4", ConvGeom.conv1d(256, self.kp2, 1, 1, 0), None, 1.0)
        self.blocks = self.enc + self.dec_up + self.dec_tail + [self.dec_out]
        self.fwd_id = 0

    def _lengths(self, F):
        lens, L = {}, F
        for b in self.enc:
            lens[b.name] = L
            L = b.g.out_hw(1, L)[1]
        self.enc_out_len = L
        L = 4
        for b in self.dec_up:
            lens[b.name] = L
            L *= 2
        for b in self.dec_tail + [self.dec_out]:
            lens[b.name] = L // 2
        return lens

    def encode(self, poses, params, buffers, training, prep=True):
        """poses (B,F,2K) -> mu, logvar (B,D) engine-owned."""
        A = self.arena
        B, F = poses.shape[0], poses.shape[1]
        if prep:
            self._prep(params, self._lengths(F))
        x, L = poses, F
        for b in self.enc:
            x = b.forward(x, B, L, params, buffers, training)
            L = x.shape[1]
        self.enc_L = L
        ones = A.get("ones", (1, 2 * self.D))
        zeros = A.get("zeros", (1, 2 * self.D), zero=True)
        ones.fill_(1.0)
        mu, logvar = A.get("mu", (B, self.D)), A.get("logvar", (B, self.D))
        ops.pose_head_fwd(x, ones, zeros, 1.0, mu, logvar)          # activation already applied: identity transform
        return mu, logvar

    def decode(self, code, params, buffers, training):
        """code (B,D) -> pred (B,64,2K)."""
        A = self.arena
        B = code.shape[0]
        x = A.get("x0", (B, 2, self.D))
        x.copy_(code.unsqueeze(1).expand(B, 2, self.D))           # F.interpolate(x.unsqueeze(-1), 2): nearest duplicate
        L = 2
        for b in self.dec_up:
            up = A.get("up:" + b.name, (B, 2 * L, b.g.cin))
            ops.upsample_add_fwd(x, None, 2 * L, out=up, tf32=self.math >= 1)           # F.interpolate(..., mode='linear') x2 (autoencoder.py:62-66)
            x = b.forward(up, B, 2 * L, params, buffers, training)
            L *= 2
        for b in self.dec_tail:
            x = b.forward(x, B, L, params, buffers, training)
        return self.dec_out.forward(x, B, L, params, buffers, training)

    def forward(self, poses, eps, params, buffers, training, lambda_kl):
        A = self.arena
        self.fwd_id += 1
        self._params = params
        mu, logvar = self.encode(poses, params, buffers, training)
        B = poses.shape[0]
        code = A.get("code", (B, self.D))
        kl = A.get("kl_out", (1,))
        self._eps, self._lambda_kl = eps, lambda_kl
        ops.vae_reparam_kl(mu, logvar, eps, lambda_kl, code, kl)
        pred = self.decode(code, params, buffers, training)
        return pred, mu, logvar, kl

    def backward(self, g_pred, grads, include_kl=True, g_mu_ext=None, g_lv_ext=None):
        """g_pred (B,64,2K) = d loss / d pred; KL gradient added at the reparameterisation when include_kl;
        g_mu_ext / g_lv_ext: extra gradients arriving at mu / logvar from outside (autograd drop-in)."""
        A, params = self.arena, self._params
        B = g_pred.shape[0]
        g = self.dec_out.backward(g_pred, B, params, grads)
        for b in reversed(self.dec_tail):
            g = b.backward(g, B, params, grads)
        L = 64
        for b in reversed(self.dec_up):
            g_up = b.backward(g, B, params, grads)                 # (B, L, cin) gradient w.r.t. the upsampled input
            L //= 2
            g = A.get("g_pre:" + b.name, (B, L, b.g.cin))
            ops.upsample_bwd(g_up, L, out=g)
        if getattr(self, "on_decoder_done", None) is not None:     # every decoder.* gradient is final (multi-GPU: first bucket)
            self.on_decoder_done()
        g_code = A.get("g_code", (B, self.D))
        torch.add(g[:, 0, :], g[:, 1, :], out=g_code)              # adjoint of the nearest x2 duplicate
        mu, logvar = A.get("mu", (B, self.D)), A.get("logvar", (B, self.D))
        g_mu, g_lv = A.get("g_mu", (B, self.D)), A.get("g_logvar", (B, self.D))
        ops.vae_reparam_kl_bwd(mu, logvar, self._eps, g_code, self._lambda_kl if include_kl else 0.0, g_mu, g_lv)
        if g_mu_ext is not None:
            g_mu.add_(g_mu_ext)
        if g_lv_ext is not None:
            g_lv.add_(g_lv_ext)
        self.encoder_backward(g_mu, g_lv, grads)

    def encoder_backward(self, g_mu, g_lv, grads):
        A, params = self.arena, self._params
        B = g_mu.shape[0]
        g = A.get("g_head", (B, self.enc_L, 2 * self.D))
        ops.pose_head_bwd(g_mu, g_lv, self.enc_L, g)
        for i in range(len(self.enc) - 1, -1, -1):
            g = self.enc[i].backward(g, B, params, grads, need_dx=i > 0)


class DiscriminatorEngine(_SeqEngine):
    """PoseSequenceDiscriminator (discriminator.py:6-23) on channels-last motion sequences (B,T,2K) -> scores (B,T')."""

    def __init__(self, n_landmarks, leaky, device, math=None):
        super().__init__(device, math)
        self.kp2 = n_landmarks * 2
        self.slope = 0.2 if leaky else 0.0
        s = self.slope
        self.seq = [SeqBlock(self, "seq.0", ConvGeom.conv1d(self.kp2, 256, 4, 2, 1), "BN", s),
                    SeqBlock(self, "seq.1", ConvGeom.conv1d(256, 512, 4, 2, 1), "BN", s),
                    SeqBlock(self, "seq.2", ConvGeom.conv1d(512, 1024, 3, 1, 1), "BN", s),
                    SeqBlock(self, "seq.3", ConvGeom.conv1d(1024, 1, 3, 1, 1), None, 1.0)]
        self.blocks = self.seq

    def prepare(self, params, T):
        lens, L = {}, T
        for b in self.seq:
            lens[b.name] = L
            L = b.g.out_hw(1, L)[1]
        self._prep(params, lens)

    def forward(self, x, params, buffers, training, tag):
        B, L = x.shape[0], x.shape[1]
        for b in self.seq:
            x = b.forward(x, B, L, params, buffers, training, tag)
            L = x.shape[1]
        return x.view(B, L)

    def backward(self, g_scores, params, grads, tag, need_dx, accumulate_params):
        B = g_scores.shape[0]
        g = g_scores.reshape(B, -1, 1)
        for i in range(len(self.seq) - 1, -1, -1):
            g = self.seq[i].backward(g, B, params, grads, tag, need_dx=(need_dx or i > 0), accumulate_params=accumulate_params)
        return g
