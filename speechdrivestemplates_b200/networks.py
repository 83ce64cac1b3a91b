"""Drop-in replacements for the reference's network classes (core/networks/**), same constructor signature
``(cfg)``, same forward signatures, same parameter / buffer names and shapes, so that the
reference's ``main.py``, YAML configs, optimizers, DDP wrapper and checkpoints drive them unchanged.

The modules hold parameters only; all arithmetic runs in libsdt_b200 through ``engine.py``.  There is no PyTorch
fallback: on a CPU tensor or without the library they raise.
"""
import math

import torch
from torch import nn

from . import engine as E


class ConvWeights(nn.Module):
    """Parameter container of one convolution in the reference's (Cout, Cin, k...) layout.

    Init consumes the RNG like the reference: nn.ConvNd's default kaiming_uniform(a=sqrt(5)) [+ uniform bias], then
    ConvNormRelu's kaiming_normal_ (building_blocks.py:44) when ``kaiming_normal``.
    """

    def __init__(self, shape, bias=False, kaiming_normal=True):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(shape))
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        if bias:
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            bound = 1.0 / math.sqrt(fan_in)
            self.bias = nn.Parameter(torch.empty(shape[0]))
            nn.init.uniform_(self.bias, -bound, bound)
        if kaiming_normal:
            nn.init.kaiming_normal_(self.weight)


class BatchNormState(nn.Module):
    """Affine parameters + running statistics of nn.BatchNorm{1,2}d (names/shapes as in its state-dict)."""

    def __init__(self, c):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(c))
        self.bias = nn.Parameter(torch.zeros(c))
        self.register_buffer("running_mean", torch.zeros(c))
        self.register_buffer("running_var", torch.ones(c))
        self.register_buffer("num_batches_tracked", torch.tensor(0, dtype=torch.long))


class NoState(nn.Module):
    """nn.InstanceNorm{1,2}d as configured by the reference has neither parameters nor buffers."""


class ConvNormRelu(nn.Module):
    """Container with the attribute names of building_blocks.ConvNormRelu (``conv``, ``norm``)."""

    def __init__(self, wshape, norm):
        super().__init__()
        self.conv = ConvWeights(wshape)
        if norm == "BN":
            self.norm = BatchNormState(wshape[0])
        elif norm == "IN":
            self.norm = NoState()
        else:
            raise NotImplementedError            # building_blocks.py:27-28,42-43


def _seq(*mods):
    return nn.Sequential(*mods)


class _EngineModule(nn.Module):
    """Shared plumbing: lazily built engine, name-ordered parameter / buffer views.

    ``conv_math`` (class default None = the process default at the time the engine is built, ``ops.set_conv_math``) is the math
    mode of THIS module's engine; set it before the first forward (the fused trainers do, from their ``conv_math`` argument)."""

    conv_math = None

    def set_conv_math(self, mode):
        """Math mode of this module's engine (and of its sub-engines); rebuilds the engine on the next call."""
        self.conv_math = None if mode is None else int(mode)
        for attr in ("_eng", "_ae_eng", "_d_eng"):
            if hasattr(self, attr):
                setattr(self, attr, None)
        for child in self.children():
            if isinstance(child, _EngineModule):
                child.set_conv_math(mode)

    def _device(self):
        return next(self.parameters()).device

    def _require_cuda(self, *tensors):
        for t in tensors:
            if t is not None and not t.is_cuda:
                raise RuntimeError("%s runs only on CUDA tensors (libsdt_b200 has no CPU fallback)" % type(self).__name__)

    def _named(self):
        names = [n for n, _ in self.named_parameters()]
        return names, [p for _, p in self.named_parameters()]

    def _buffers_dict(self):
        return dict(self.named_buffers())


class _AudioEncoder(nn.Module):
    def __init__(self, norm):
        super().__init__()
        blocks = []
        for i in range(4):
            a, b = E.ENC2D[2 * i], E.ENC2D[2 * i + 1]
            blocks.append(_seq(ConvNormRelu((a[1], a[2], a[3], a[4]), norm), ConvNormRelu((b[1], b[2], b[3], b[4]), norm)))
        self.specgram_encoder_2d = _seq(*blocks)


class _UNet1D(nn.Module):
    def __init__(self, norm, code_dim):
        super().__init__()
        self.e0 = ConvNormRelu((256, 256 + (code_dim or 0), 3), norm)
        self.e1 = ConvNormRelu((256, 256, 3), norm)
        for n in E.UNET_E[2:]:
            setattr(self, n, ConvNormRelu((256, 256, 4), norm))
        for n in E.UNET_D:
            setattr(self, n, ConvNormRelu((256, 256, 3), norm))


class SequenceGeneratorCNN(_EngineModule):
    """core/networks/keypoints_generation/generator.py:87-117.

    forward(x: (B,80,T_mel) f32 CUDA, num_frames: int, code: (B,D) or None) -> (B, num_frames, 2, NUM_LANDMARKS),
    autograd-connected to the parameters and to ``code``.
    """

    def __init__(self, cfg):
        super().__init__()
        self.cfg = cfg
        gcfg = cfg.VOICE2POSE.GENERATOR
        self.norm_kind = gcfg.NORM
        self.leaky = bool(gcfg.LEAKY_RELU)
        self.code_dim = gcfg.CLIP_CODE.DIMENSION
        self.n_landmarks = cfg.DATASET.NUM_LANDMARKS
        self.audio_encoder = _AudioEncoder(self.norm_kind)
        self.unet = _UNet1D(self.norm_kind, self.code_dim)
        self.decoder = _seq(*[ConvNormRelu((256, 256, 3), self.norm_kind) for _ in range(4)],
                            ConvWeights((self.n_landmarks * 2, 256, 1), bias=True, kaiming_normal=False))
        self._eng = None

    def engine(self):
        dev = self._device()
        if self._eng is None or self._eng.device != dev:
            self._eng = E.GeneratorEngine(self.norm_kind, self.leaky, self.code_dim, self.n_landmarks, dev, self.conv_math)
        return self._eng

    def forward(self, x, num_frames, code=None):
        self._require_cuda(x, code)
        if (self.code_dim is not None) != (code is not None):
            raise ValueError("clip code must be given exactly when CLIP_CODE.DIMENSION is set")
        names, params = self._named()
        out = _GeneratorFn.apply(self, int(num_frames), x, code, *params)
        return out


class _GeneratorFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, num_frames, mel, code, *params):
        eng = module.engine()
        names = [n for n, _ in module.named_parameters()]
        pdict = {n: p.detach() for n, p in zip(names, params)}
        mel = mel.detach().contiguous().float()
        codec = code.detach().contiguous().float() if code is not None else None
        pred = eng.forward(mel, num_frames, codec, pdict, module.training, module._buffers_dict())
        ctx.module, ctx.names, ctx.fwd_id = module, names, eng.fwd_id
        ctx.has_code = code is not None
        ctx.shapes = [p.shape for p in params]
        B = mel.shape[0]
        return pred.view(B, num_frames, 2, module.n_landmarks).clone()

    @staticmethod
    def backward(ctx, g):
        module = ctx.module
        eng = module.engine()
        if eng.fwd_id != ctx.fwd_id:
            raise RuntimeError("SequenceGeneratorCNN.backward: the engine's saved activations belong to a later forward; "
                               "call backward before running the module again")
        dev = g.device
        grads = {n: torch.empty(s, device=dev) for n, s in zip(ctx.names, ctx.shapes)}
        B, F = g.shape[0], g.shape[1]
        g_code = torch.empty(B, module.code_dim, device=dev) if ctx.has_code else None
        eng.backward(g.contiguous().view(B, F, -1).float(), grads, g_code)
        return (None, None, None, g_code) + tuple(grads[n] for n in ctx.names)


class PoseSeqEncoder(_EngineModule):
    """core/networks/poses_reconstruction/autoencoder.py:8-35.  forward(x: (B,F,2,K)) -> (mu, logvar), each (B, CODE_DIM).

    Inference / no-grad use (the FGD feature extractor of Voice2Pose, voice2pose.py:162-176): BatchNorm follows
    ``self.training`` exactly like the reference (batch statistics + running-stat update when training).
    """

    def __init__(self, cfg):
        super().__init__()
        acfg = cfg.POSE2POSE.AUTOENCODER
        if acfg.NORM != "BN":
            raise NotImplementedError("PoseSeqEncoder with NORM=%s" % acfg.NORM)
        self.leaky = bool(acfg.LEAKY_RELU)
        self.code_dim = acfg.CODE_DIM
        self.n_landmarks = cfg.DATASET.NUM_LANDMARKS
        cin = self.n_landmarks * 2
        shapes = [(256, cin, 3), (256, 256, 3)] + [(256, 256, 4)] * 4 + [(self.code_dim * 2, 256, 4)]
        self.blocks = _seq(*[ConvNormRelu(s, "BN") for s in shapes])
        self._eng = None

    def engine(self):
        dev = self._device()
        if self._eng is None or self._eng.arena.device != dev:
            self._eng = E.PoseEncoderEngine(self.n_landmarks, self.code_dim, self.leaky, dev, self.conv_math)
        return self._eng

    def forward(self, x):
        self._require_cuda(x)
        if torch.is_grad_enabled() and x.requires_grad:
            raise NotImplementedError("PoseSeqEncoder backward (Pose2Pose training) is not implemented yet; "
                                      "call under torch.no_grad() as voice2pose.py:162 does")
        B, F = x.shape[0], x.shape[1]
        xin = x.detach().reshape(B, F, -1).contiguous().float()
        params = {n: p.detach() for n, p in self.named_parameters()}
        mu, logvar = self.engine().forward(xin, params, self._buffers_dict(), self.training)
        return mu.clone(), logvar.clone()


class _PoseSeqDecoder(nn.Module):
    """Parameter container with the attribute names of autoencoder.PoseSeqDecoder (autoencoder.py:37-57)."""

    def __init__(self, code_dim, kp2):
        super().__init__()
        self.d5 = ConvNormRelu((256, code_dim, 3), "BN")
        for n in ("d4", "d3", "d2", "d1"):
            setattr(self, n, ConvNormRelu((256, 256, 3), "BN"))
        self.blocks = _seq(*[ConvNormRelu((256, 256, 3), "BN") for _ in range(4)],
                           ConvWeights((kp2, 256, 1), bias=True, kaiming_normal=False))


class Autoencoder(_EngineModule):
    """core/networks/poses_reconstruction/autoencoder.py:71-92 (pose VAE).

    forward(x: (B,F,2,K), num_frames, mel=None, external_code=None) -> (pred (B,F,2,K), mu, logvar), autograd-connected.
    The N(0,1) draw uses ``torch.randn(logvar.shape, device=...)`` exactly like autoencoder.py:86, i.e. the same
    generator stream as the reference on the same device."""

    def __init__(self, cfg):
        super().__init__()
        self.cfg = cfg
        acfg = cfg.POSE2POSE.AUTOENCODER
        if acfg.NORM != "BN":
            raise NotImplementedError("Autoencoder with NORM=%s" % acfg.NORM)
        self.leaky = bool(acfg.LEAKY_RELU)
        self.code_dim = acfg.CODE_DIM
        self.n_landmarks = cfg.DATASET.NUM_LANDMARKS
        self.encoder = PoseSeqEncoder(cfg)
        self.decoder = _PoseSeqDecoder(self.code_dim, self.n_landmarks * 2)
        self._ae_eng = None
        self.lambda_kl_fused = 0.0          # the drop-in leaves the KL term to the caller's torch expression (pose2pose.py:77)

    def engine(self):
        from .engine_seq import AutoencoderEngine
        dev = self._device()
        if self._ae_eng is None or self._ae_eng.device != dev:
            self._ae_eng = AutoencoderEngine(self.n_landmarks, self.code_dim, self.leaky, dev, self.conv_math)
        return self._ae_eng

    def forward(self, x, num_frames, mel=None, external_code=None):
        self._require_cuda(x if x is not None else external_code)
        names, params = self._named()
        if external_code is not None:                                   # autoencoder.py:80-83
            pdict = {n: p.detach() for n, p in zip(names, params)}
            eng = self.engine()
            eng._prep(pdict, eng._lengths(num_frames), with_dgrad=False)
            with torch.no_grad():
                pred = eng.decode(external_code.detach().contiguous().float(), pdict, self._buffers_dict(), self.training)
            B = external_code.shape[0]
            return pred.view(B, num_frames, 2, self.n_landmarks).clone(), external_code, torch.zeros_like(external_code)
        eps = torch.randn((x.shape[0], self.code_dim), device=x.device)  # autoencoder.py:86
        return _AutoencoderFn.apply(self, int(num_frames), x, eps, *params)


class _AutoencoderFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, num_frames, x, eps, *params):
        eng = module.engine()
        names = [n for n, _ in module.named_parameters()]
        pdict = {n: p.detach() for n, p in zip(names, params)}
        B = x.shape[0]
        xin = x.detach().reshape(B, x.shape[1], -1).contiguous().float()
        pred, mu, logvar, _kl = eng.forward(xin, eps.contiguous(), pdict, module._buffers_dict(), module.training, 0.0)
        ctx.module, ctx.names, ctx.fwd_id = module, names, eng.fwd_id
        ctx.shapes = [p.shape for p in params]
        ctx.set_materialize_grads(False)
        return pred.view(B, num_frames, 2, module.n_landmarks).clone(), mu.clone(), logvar.clone()

    @staticmethod
    def backward(ctx, g_pred, g_mu, g_lv):
        module = ctx.module
        eng = module.engine()
        if eng.fwd_id != ctx.fwd_id:
            raise RuntimeError("Autoencoder.backward: saved activations were overwritten by a later forward")
        dev = eng.device
        grads = {n: torch.zeros(s, device=dev) for n, s in zip(ctx.names, ctx.shapes)}
        B = eng.arena.bufs["mu"].shape[0]
        gp = g_pred.contiguous().view(B, -1, module.n_landmarks * 2).float().clone() if g_pred is not None else \
            torch.zeros(B, 64, module.n_landmarks * 2, device=dev)
        eng.backward(gp, grads, include_kl=False, g_mu_ext=g_mu, g_lv_ext=g_lv)
        return (None, None, None, None) + tuple(grads[n] for n in ctx.names)


class PoseSequenceDiscriminator(_EngineModule):
    """core/networks/keypoints_generation/discriminator.py:6-23.  forward(x: (B,T,2,K)) -> (B,T') scores.

    Each forward call gets its own activation slot, so the three calls of a step (real, fake, fake.detach();
    voice2pose.py:191-193) can all be back-propagated, each with its own BatchNorm batch statistics."""

    SLOTS = 4

    def __init__(self, cfg):
        super().__init__()
        self.cfg = cfg
        self.leaky = bool(cfg.VOICE2POSE.POSE_DISCRIMINATOR.LEAKY_RELU)
        self.n_landmarks = cfg.DATASET.NUM_LANDMARKS
        kp2 = self.n_landmarks * 2
        self.seq = _seq(ConvNormRelu((256, kp2, 4), "BN"), ConvNormRelu((512, 256, 4), "BN"), ConvNormRelu((1024, 512, 3), "BN"),
                        ConvWeights((1, 1024, 3), bias=True, kaiming_normal=False))
        self._d_eng = None
        self._calls = 0

    def engine(self):
        from .engine_seq import DiscriminatorEngine
        dev = self._device()
        if self._d_eng is None or self._d_eng.device != dev:
            self._d_eng = DiscriminatorEngine(self.n_landmarks, self.leaky, dev, self.conv_math)
        return self._d_eng

    def forward(self, x):
        self._require_cuda(x)
        names, params = self._named()
        self._calls += 1
        return _DiscriminatorFn.apply(self, "/s%d" % (self._calls % self.SLOTS), x, *params)


class _DiscriminatorFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, tag, x, *params):
        eng = module.engine()
        names = [n for n, _ in module.named_parameters()]
        pdict = {n: p.detach() for n, p in zip(names, params)}
        B, T = x.shape[0], x.shape[1]
        xin = x.detach().reshape(B, T, -1).contiguous().float()
        eng.prepare(pdict, T)
        scores = eng.forward(xin, pdict, module._buffers_dict(), module.training, tag)
        ctx.module, ctx.names, ctx.tag, ctx.pdict = module, names, tag, pdict
        ctx.shapes = [p.shape for p in params]
        ctx.x_shape, ctx.need_dx = x.shape, x.requires_grad
        ctx.calls = module._calls
        return scores.clone()

    @staticmethod
    def backward(ctx, g):
        module = ctx.module
        eng = module.engine()
        if module._calls - ctx.calls >= module.SLOTS:
            raise RuntimeError("PoseSequenceDiscriminator.backward: activation slot was reused by later forwards")
        grads = {n: torch.zeros(s, device=eng.device) for n, s in zip(ctx.names, ctx.shapes)}
        eng.prepare(ctx.pdict, ctx.x_shape[1])
        dx = eng.backward(g.contiguous().float().clone(), ctx.pdict, grads, ctx.tag, ctx.need_dx, False)
        gx = dx.view(ctx.x_shape).clone() if ctx.need_dx else None
        return (None, None, gx) + tuple(grads[n] for n in ctx.names)


def module_dict():
    """Entries for the reference's registry core.networks.module_dict (core/networks/__init__.py:6-11)."""
    return {"SequenceGeneratorCNN": SequenceGeneratorCNN, "PoseSequenceDiscriminator": PoseSequenceDiscriminator,
            "Autoencoder": Autoencoder, "PoseSeqEncoder": PoseSeqEncoder}


def get_model(name):
    """core/networks/__init__.py:14-19 semantics (KeyError on unknown names)."""
    obj = module_dict().get(name)
    if obj is None:
        raise KeyError("Unknown model: %s" % name)
    return obj
