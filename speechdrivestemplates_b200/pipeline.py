"""Voice2Pose step model and train step on the B200 kernels.

``Voice2PoseModel`` mirrors core/pipelines/voice2pose.py:22-210 (constructor signature, attribute / parameter /
buffer names, ``forward(batch, dataset) -> (losses_dict, results_dict)`` with an autograd-connected ``G_loss``), so
the reference's ``Voice2Pose`` trainer can hold it unchanged.  ``Voice2PoseTrainer.train_step`` is the numeric part
of core/pipelines/voice2pose.py:281-312 as ONE fused device program: host batch -> H2D -> mel -> generator -> losses
-> FGD encoder x2 -> f64 final results + metrics -> backward -> (NCCL all-reduce) -> Adam, replayable as a CUDA graph.
"""
import math
from collections import OrderedDict

import os

import torch
from torch import nn

# diagnostic switch (wrong results: the FGD / metrics side stream is skipped); bench.py refuses to run with it set
_DIAG_SKIP_SIDE = bool(os.environ.get("SDT_DIAG_SKIP_SIDE"))
# multi-GPU communication mode when SDT_COMM is unset (profiles/r2_multi_gpu_modes.txt, per step at 2 / 8 GPUs):
#   "p2p"     3.07 / 3.12 ms  the library's own all-reduce over peer memory (csrc/p2p.cu: peer loads at 2 GPUs, NVLS multicast
#                             from 3 GPUs up) between two cross-GPU barriers, captured inside the ONE step graph; no NCCL in the step
#   "serial"  3.08 / 3.19 ms  one flat NCCL all-reduce between two graphs (round 1)
#   "overlap" 3.11 / 3.23 ms  NCCL all-reduce of three gradient buckets beside the backward pass, inside the step graph
# p2p falls back to serial (with a warning) where torch's symmetric memory cannot be set up.
_DEFAULT_COMM = "p2p"

from . import _lib, ops, parallel
from .networks import PoseSeqEncoder, SequenceGeneratorCNN, get_model


# ------------------------------------------------------------------------------------------------
# mel front end module (state-dict compatible with torchaudio.transforms.MelSpectrogram)
# ------------------------------------------------------------------------------------------------
def melscale_fbanks_htk(n_freqs=257, f_min=55.0, f_max=7500.0, n_mels=80, sample_rate=16000):
    """HTK triangular filterbank, no normalisation, evaluated in fp32 as torchaudio does -> (n_freqs, n_mels)."""
    all_freqs = torch.linspace(0, sample_rate // 2, n_freqs)
    m_min = 2595.0 * math.log10(1.0 + f_min / 700.0)
    m_max = 2595.0 * math.log10(1.0 + f_max / 700.0)
    m_pts = torch.linspace(m_min, m_max, n_mels + 2)
    f_pts = 700.0 * (10.0 ** (m_pts / 2595.0) - 1.0)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
    down = -slopes[:, :-2] / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    return torch.clamp(torch.minimum(down, up), min=0.0)


class _Buffer(nn.Module):
    def __init__(self, name, value):
        super().__init__()
        self.register_buffer(name, value)


class MelSpectrogram(nn.Module):
    """Buffers ``spectrogram.window`` (400) and ``mel_scale.fb`` (257, 80) like the torchaudio module built at
    voice2pose.py:27-30; forward(audio (B, L)) -> (B, 80, 1 + L//160) power mel spectrogram via sdt_mel_fwd."""

    def __init__(self):
        super().__init__()
        self.spectrogram = _Buffer("window", torch.hann_window(400))
        self.mel_scale = _Buffer("fb", melscale_fbanks_htk())
        self._tables = None
        self._tables_key = None

    def tables(self):
        fb = self.mel_scale.fb
        key = (fb.data_ptr(), fb._version, fb.device)
        if self._tables_key != key:
            self._tables = ops.mel_band_tables(fb)
            self._tables_key = key
        return self._tables

    def forward(self, audio, out=None):
        if not audio.is_cuda:
            raise RuntimeError("MelSpectrogram runs only on CUDA tensors (libsdt_b200 has no CPU fallback)")
        lead = audio.shape[:-1]
        a2 = audio.detach().reshape(-1, audio.shape[-1]).contiguous().float()
        mel = ops.mel_fwd(a2, self.spectrogram.window, self.tables(), out=out)
        return mel.view(*lead, 80, mel.shape[-1])


# ------------------------------------------------------------------------------------------------
# step engine (no autograd): forward of the whole loss graph + backward into caller-provided gradient tensors
# ------------------------------------------------------------------------------------------------
class Voice2PoseStepEngine:
    """Sequences the kernels of Voice2PoseModel.forward (training branch) and its backward."""

    def __init__(self, model):
        self.model = model
        self.arena = None

    def _arena(self, device):
        from .engine import Arena
        if self.arena is None or self.arena.device != device:
            self.arena = Arena(device)
        return self.arena

    def forward(self, audio, poses, clip_index, stat=None, code_table=None, p2g_stats=None, defer_side=False):
        """audio (B,L) f32, poses (B,F,2,K) f32, clip_index (B) i64 on device; stat = (mean, std, scale) f64 or None;
        p2g_stats = (mean_parted, std_parted, mean_global, std_global) f32 (242) when HIERARCHICAL_POSE is False.

        Returns a dict of engine-owned device tensors (valid until the next forward).
        """
        m = self.model
        cfg = m.cfg
        dev = audio.device
        A = self._arena(dev)
        B, F = poses.shape[0], poses.shape[1]
        K2 = poses.shape[2] * poses.shape[3]
        gcfg = cfg.VOICE2POSE.GENERATOR
        out = OrderedDict()
        mel = m.mel_transfm(audio, out=A.get("mel", (audio.shape[0], 80, 1 + audio.shape[1] // 160)))
        D = gcfg.CLIP_CODE.DIMENSION
        code = None
        self._code_live = False
        if D is not None:
            table = code_table if code_table is not None else m.clips_code
            if table.device != dev:                      # external (frozen) code table loaded from a checkpoint (voice2pose.py:40-55)
                table = table.to(dev).contiguous().float()
                if code_table is None:
                    m.clips_code = table
            code = A.get("code", (B, D))
            kl = A.get("kl_out", (2,))
            ops.code_gather_kl(table.detach(), clip_index, float(gcfg.LAMBDA_CLIP_KL), code, kl, A.get("g_code_kl", (B, D)))
            out["G_clipcode_kl_loss"] = kl[0:1]
            out["kl_applied"] = kl[1:2]
            self._code_live = True
        gparams = {n: p.detach() for n, p in m.netG.named_parameters()}
        pred = m.netG.engine().forward(mel, F, code, gparams, m.netG.training, m.netG._buffers_dict())
        reg = A.get("reg_out", (1,))
        self._g_pred = A.get("g_pred", (B, F, K2))
        ops.l1_loss(pred, poses, float(gcfg.LAMBDA_REG), reg, self._g_pred, A.get("l1_partial", (1024,)))
        out["G_reg_loss"] = reg
        g_loss = A.get("g_loss", (1,))
        if D is not None:
            torch.add(reg, out["G_clipcode_kl_loss"], out=g_loss)       # the KL slot holds 0 when the guard skipped it
        else:
            g_loss.copy_(reg)
        out["G_loss"] = g_loss
        out["poses_pred_batch"] = pred.view(B, F, 2, -1)
        out["condition_code"] = code
        self._side_args = (out, pred, poses, stat, p2g_stats)
        if not defer_side:
            self.run_side()
        self._clip_index = clip_index
        self._B, self._F = B, F
        return out

    def run_side(self):
        """The part of the step nothing else depends on: FGD features of prediction / ground truth and the f64 final
        results + metrics.  The fused trainer runs it on a second stream, concurrently with the backward pass."""
        out, pred, poses, stat, p2g_stats = self._side_args
        m = self.model
        cfg = m.cfg
        A = self.arena
        B, F = poses.shape[0], poses.shape[1]
        K2 = poses.shape[2] * poses.shape[3]
        # FGD feature extractor on prediction and ground truth (voice2pose.py:162-176), no gradient
        if cfg.VOICE2POSE.POSE_ENCODER.NAME is not None:
            pe = m.pose_encoder
            pparams = {n: p.detach() for n, p in pe.named_parameters()}
            pbuf = pe._buffers_dict()
            pe_pred, pe_gt = pred, poses.view(B, F, K2)
            if not cfg.DATASET.HIERARCHICAL_POSE:          # dataset.transform_normalized_parted2global (voice2pose.py:168-169)
                if p2g_stats is None:
                    raise ValueError("HIERARCHICAL_POSE=False needs the speaker's parted and global statistics (p2g_stats)")
                pe_pred = ops.pose_parted2global(pred, *p2g_stats, out=A.get("p2g_pred", (B, F, K2)))
                pe_gt = ops.pose_parted2global(poses.view(B, F, K2), *p2g_stats, out=A.get("p2g_gt", (B, F, K2)))
            out["mu_pred"], out["logvar_pred"] = pe.engine().forward(pe_pred, pparams, pbuf, pe.training, tag="/pred")
            out["mu_gt"], out["logvar_gt"] = pe.engine().forward(pe_gt, pparams, pbuf, pe.training, tag="/gt")
        if stat is not None:          # dataset.get_final_results x2 + evaluate_step (voice2pose.py:289-292)
            mean, std, scale = stat
            hier = bool(cfg.DATASET.HIERARCHICAL_POSE)
            fp = ops.pose_final_results(pred.view(B, F, 2, -1), mean, std, scale, hier, out=A.get("final_pred", (B, F, 2, K2 // 2), torch.float64))
            fg = ops.pose_final_results(poses, mean, std, scale, hier, out=A.get("final_gt", (B, F, 2, K2 // 2), torch.float64))
            met = ops.pose_metrics(fp, fg, A.get("met_partial", (2 * B,), torch.float64), A.get("met_out", (2,), torch.float64))
            out["final_pred"], out["final_gt"] = fp, fg
            out["L2_dist"], out["lip_sync_error_n"] = met[0:1], met[1:2]
        return out

    def backward(self, g_grads, g_table=None, g_pred=None):
        """d G_loss: generator parameter gradients into g_grads[name]; dense clip-code gradient added into g_table
        (which the caller has zeroed).  g_pred overrides the stored L1 gradient (B,F,2K)."""
        m = self.model
        D = m.cfg.VOICE2POSE.GENERATOR.CLIP_CODE.DIMENSION
        A = self.arena
        g_code = A.get("g_code_e0", (self._B, D)) if D is not None else None
        m.netG.engine().backward(self._g_pred if g_pred is None else g_pred, g_grads, g_code)
        if D is not None and g_table is not None:
            ops.code_scatter_grad(g_code, A.get("g_code_kl", (self._B, D)), self._clip_index, g_table)
        return g_code


class _V2PLossFn(torch.autograd.Function):
    """Autograd bridge for the drop-in model: (pred, G_reg_loss, KL) as one node over the step engine."""

    @staticmethod
    def forward(ctx, model, audio, poses, clip_index, clips_code, p2g_stats, *gparams):
        eng = model.step_engine()
        out = eng.forward(audio, poses, clip_index, None, clips_code, p2g_stats)
        ctx.model, ctx.eng = model, eng
        ctx.fwd_id = model.netG.engine().fwd_id
        ctx.names = [n for n, _ in model.netG.named_parameters()]
        ctx.shapes = [p.shape for p in gparams]
        ctx.table_shape = clips_code.shape if clips_code is not None else None
        ctx.code_needs_grad = clips_code is not None and clips_code.requires_grad
        ctx.set_materialize_grads(False)
        ctx.results = out
        kl = out.get("G_clipcode_kl_loss")
        return (out["poses_pred_batch"].clone(), out["G_reg_loss"].clone().squeeze(0),
                kl.clone().squeeze(0) if kl is not None else torch.zeros((), device=audio.device))

    @staticmethod
    def backward(ctx, g_pred_ext, g_reg, g_kl):
        model, eng = ctx.model, ctx.eng
        if model.netG.engine().fwd_id != ctx.fwd_id:
            raise RuntimeError("Voice2PoseModel.backward: saved activations were overwritten by a later forward")
        dev = eng._g_pred.device
        gp = eng._g_pred
        if g_reg is None:
            gp = torch.zeros_like(gp)
        elif float(g_reg) != 1.0:
            gp = gp * g_reg
        if g_pred_ext is not None:
            gp = gp + g_pred_ext.reshape(gp.shape)
        grads = {n: torch.empty(s, device=dev) for n, s in zip(ctx.names, ctx.shapes)}
        g_table = None
        if ctx.code_needs_grad:
            g_table = torch.zeros(ctx.table_shape, device=dev)
            if g_kl is None or float(g_kl) != 1.0:
                eng.arena.bufs["g_code_kl"].mul_(0.0 if g_kl is None else float(g_kl))
        eng.backward(grads, g_table, gp.contiguous())
        return (None, None, None, None, g_table, None) + tuple(grads[n] for n in ctx.names)


# ------------------------------------------------------------------------------------------------
# drop-in step model
# ------------------------------------------------------------------------------------------------
class Voice2PoseModel(nn.Module):
    """core/pipelines/voice2pose.py:22-210 on libsdt_b200.  State-dict groups: ``clips_code``,
    ``mel_transfm.spectrogram.window``, ``mel_transfm.mel_scale.fb``, ``netG.*``, ``pose_encoder.*``."""

    def __init__(self, cfg, state_dict=None, num_train_samples=None, rank=0):
        super().__init__()
        self.cfg = cfg
        self.mel_transfm = MelSpectrogram()
        self.netG = get_model(cfg.VOICE2POSE.GENERATOR.NAME)(cfg)
        ccfg = cfg.VOICE2POSE.GENERATOR.CLIP_CODE
        if ccfg.DIMENSION is not None:
            if ccfg.EXTERNAL_CODE:
                path = ccfg.EXTERNAL_CODE_PTH or cfg.VOICE2POSE.POSE_ENCODER.AE_CHECKPOINT
                if path is None:
                    raise RuntimeError("External code not provide.")             # voice2pose.py:48
                ckpt = torch.load(path, map_location={"cuda:0": "cuda:%d" % rank})
                codes = {k.replace("module.", ""): v for k, v in ckpt["model_state_dict"].items() if "clip_code" in k}
                self.clips_code = codes["clip_code_mu"]                              # plain tensor, not trained (:50-55)
            else:
                if num_train_samples is None:
                    assert state_dict is not None, "No state_dict available, while no dataset is configured."
                    num_train_samples = state_dict["module.clips_code"].shape[0]
                if ccfg.FRAME_VARIANT:
                    # the reference builds an (N, D, NUM_FRAMES) table (voice2pose.py:66-67) but its own generator then fails:
                    # generator.py:110 does code.unsqueeze(2).repeat([1, 1, T]) on the 3-D code -> a 4-D tensor repeated with 3
                    # factors, a RuntimeError in torch.  The option cannot run in the reference either; refuse it up front.
                    raise NotImplementedError("CLIP_CODE.FRAME_VARIANT (unusable in the reference too: generator.py:110)")
                self.clips_code = nn.Parameter(torch.zeros(num_train_samples, ccfg.DIMENSION), requires_grad=bool(ccfg.TRAIN))
        else:
            self.clips_code = None
        if cfg.VOICE2POSE.POSE_ENCODER.NAME is not None:
            self.pose_encoder = get_model(cfg.VOICE2POSE.POSE_ENCODER.NAME)(cfg)
            self.pose_encoder.eval()                                                 # voice2pose.py:77
        if cfg.VOICE2POSE.POSE_DISCRIMINATOR.NAME is not None:
            self.netD_pose = get_model(cfg.VOICE2POSE.POSE_DISCRIMINATOR.NAME)(cfg)
            self.pose_gan_criterion = nn.MSELoss()                                   # voice2pose.py:81-82 (LSGAN)
        self._step_engine = None

    def step_engine(self):
        if self._step_engine is None:
            self._step_engine = Voice2PoseStepEngine(self)
        return self._step_engine

    def set_conv_math(self, mode):
        """Convolution math mode of every engine of this model (networks._EngineModule.set_conv_math)."""
        for name in ("netG", "pose_encoder", "netD_pose"):
            mod = getattr(self, name, None)
            if mod is not None:
                mod.set_conv_math(mode)

    def _p2g_stats(self, batch, dataset, device):
        """Parted and global statistics of the batch's (single) speaker as fp32 device tensors (gesture_dataset.py:228-229)."""
        if "stat_parted" in batch and batch["stat_parted"] is not None:
            sp, sg = batch["stat_parted"], batch["stat_global"]
        else:
            k = self.cfg.DATASET.NUM_LANDMARKS
            sp = dataset.get_speaker_stat(batch["speaker"][0], k, True)
            sg = dataset.get_speaker_stat(batch["speaker"][0], k, False)
        f32 = lambda a: torch.as_tensor(a, dtype=torch.float64).float().to(device).contiguous()
        return (f32(sp["mean"]), f32(sp["std"]), f32(sg["mean"]), f32(sg["std"]))

    def _condition_code(self, batch, dataset, clip_indices, audio, poses_gt, return_loss, interpolation_coeff):
        """Eval-time code selection (voice2pose.py:96-120)."""
        cfg = self.cfg
        ccfg = cfg.VOICE2POSE.GENERATOR.CLIP_CODE
        dev = audio.device
        if ccfg.SAMPLE_FROM_NORMAL:
            return torch.randn([len(clip_indices), ccfg.DIMENSION]).to(dev)
        if ccfg.TEST_WITH_GT_CODE:
            assert cfg.VOICE2POSE.POSE_ENCODER.NAME is not None
            with torch.no_grad():
                mu_gt, _ = self.pose_encoder(self._fgd_input(poses_gt, batch, dataset))     # voice2pose.py:102-105
            return mu_gt
        table = self.clips_code.to(dev)
        if cfg.DEMO.CODE_INDEX is not None:
            assert not return_loss, 'WARNING: Do not set "DEMO.CODE_INDEX" in train or test mode!'
            assert 0 <= cfg.DEMO.CODE_INDEX < table.size(0)
            code = table[torch.full((len(audio),), cfg.DEMO.CODE_INDEX, dtype=torch.long, device=dev)]
            if interpolation_coeff is not None:
                assert cfg.DEMO.CODE_INDEX_B < table.size(0)
                code_b = table[torch.full((len(audio),), cfg.DEMO.CODE_INDEX_B, dtype=torch.long, device=dev)]
                code = code * (1 - interpolation_coeff) + code_b * interpolation_coeff
            return code
        return table[torch.randint(table.size(0), (len(audio),)).to(dev)]

    def _fgd_input(self, poses, batch, dataset):
        """Input of the FGD feature extractor: the poses as they are (hierarchical statistics) or
        ``dataset.transform_normalized_parted2global`` of them (voice2pose.py:102-105,164-169) as one kernel."""
        if self.cfg.DATASET.HIERARCHICAL_POSE:
            return poses
        p2g = self._p2g_stats(batch, dataset, poses.device)
        return ops.pose_parted2global(poses.detach().contiguous().float(), *p2g)

    def _discriminator_losses(self, losses, g_loss, poses_gt, pred):
        """voice2pose.py:179-208: three discriminator passes (each with its own BatchNorm statistics in train mode) + LSGAN terms."""
        dcfg = self.cfg.VOICE2POSE.POSE_DISCRIMINATOR
        real_batch, fake_batch = poses_gt, pred
        if dcfg.WHITE_LIST is not None:
            real_batch, fake_batch = real_batch[..., dcfg.WHITE_LIST], fake_batch[..., dcfg.WHITE_LIST]
        if dcfg.MOTION:
            real_batch = real_batch[:, 1:, ...] - real_batch[:, :-1, ...]
            fake_batch = fake_batch[:, 1:, ...] - fake_batch[:, :-1, ...]
        score_real = self.netD_pose(real_batch)
        score_fake = self.netD_pose(fake_batch)
        score_fake_detach = self.netD_pose(fake_batch.detach())
        g_gan = self.pose_gan_criterion(score_fake, torch.ones_like(score_fake)) * dcfg.LAMBDA_GAN
        losses["G_pose_gan_loss"] = g_gan
        losses["G_loss"] = g_loss + g_gan
        d_fake = self.pose_gan_criterion(score_fake_detach, torch.zeros_like(score_fake_detach))
        d_real = self.pose_gan_criterion(score_real, torch.ones_like(score_real))
        losses["D_pose_gan_loss"] = (d_real + d_fake) * dcfg.LAMBDA_GAN
        losses["pose_score_fake"] = score_fake.mean()
        losses["pose_score_real"] = score_real.mean()

    def forward(self, batch, dataset=None, return_loss=True, interpolation_coeff=None):
        cfg = self.cfg
        audio = batch["audio"].cuda()
        clip_indices = batch["clip_index"].cuda()
        num_frames = int(batch["num_frames"][0].item())
        poses_gt = batch["poses"].cuda() if return_loss else None
        D = cfg.VOICE2POSE.GENERATOR.CLIP_CODE.DIMENSION

        if self.training and return_loss:
            gparams = [p for _, p in self.netG.named_parameters()]
            p2g = None
            if cfg.VOICE2POSE.POSE_ENCODER.NAME is not None and not cfg.DATASET.HIERARCHICAL_POSE:
                p2g = self._p2g_stats(batch, dataset, audio.device)
            pred, reg, kl = _V2PLossFn.apply(self, audio.contiguous().float(), poses_gt.contiguous().float(),
                                             clip_indices.contiguous(), self.clips_code if D is not None else None, p2g, *gparams)
            res = _V2PLossFn_last_results(self)
            losses = OrderedDict()
            losses["G_reg_loss"] = reg
            g_loss = reg.clone()
            if D is not None and float(res["kl_applied"]) != 0.0:          # the reference's host-side guard (voice2pose.py:154)
                losses["G_clipcode_kl_loss"] = kl
                g_loss = g_loss + kl
            losses["G_loss"] = g_loss
            results = {"poses_pred_batch": pred, "condition_code": res["condition_code"], "poses_gt_batch": poses_gt}
            for k in ("mu_pred", "mu_gt", "logvar_pred", "logvar_gt"):
                if k in res:
                    results[k] = res[k].clone()
            if hasattr(self, "netD_pose"):                                           # voice2pose.py:179-208
                self._discriminator_losses(losses, g_loss, poses_gt, pred)
            return losses, results

        # eval / demo: code selection (voice2pose.py:96-120), then the same loss graph without gradients.  (In the reference
        # `self.training` only switches the code selection and the BatchNorm mode; the loss terms are computed either way.)
        code = None
        if D is not None:
            code = self._condition_code(batch, dataset, clip_indices, audio, poses_gt, return_loss, interpolation_coeff)
        chunk = int(getattr(self, "demo_chunk_frames", 0) or os.environ.get("SDT_DEMO_CHUNK_FRAMES", "0"))
        if not return_loss and chunk > 0 and num_frames > 2 * chunk and audio.shape[0] == 1 and getattr(self.netG, "norm_kind", None) == "IN":
            # long-audio demo (Trainer.demo -> demo_step, trainer.py:459-484, voice2pose.py:386-410): the time-tiled forward of
            # inference.StreamingGenerator over THIS model's generator and mel front end -- same result, a fraction of the memory
            from .inference import StreamingGenerator
            sg = getattr(self, "_streaming", None)
            if sg is None or sg.chunk_frames != chunk:
                sg = StreamingGenerator(self.cfg, audio.device, conv_math=self.netG.conv_math, chunk_frames=chunk, netG=self.netG, mel=self.mel_transfm)
                object.__setattr__(self, "_streaming", sg)               # not a sub-module: the generator is already registered
            pred = sg.forward_device(audio.contiguous().float(), num_frames, code)
            return {"poses_pred_batch": pred, "condition_code": code}
        with torch.no_grad():
            mel = self.mel_transfm(audio)
            pred = self.netG(mel, num_frames, code)
        results = {"poses_pred_batch": pred, "condition_code": code}
        if not return_loss:
            return results
        results["poses_gt_batch"] = poses_gt
        losses = OrderedDict()
        with torch.no_grad():
            gt32 = poses_gt.contiguous().float()
            reg = torch.empty(1, device=pred.device)
            ops.l1_loss(pred.contiguous(), gt32, float(cfg.VOICE2POSE.GENERATOR.LAMBDA_REG), reg, None,
                        torch.empty(1024, device=pred.device))                                # voice2pose.py:141-142
            losses["G_reg_loss"] = reg.squeeze(0)
            g_loss = losses["G_reg_loss"].clone()
            if code is not None:                                                               # voice2pose.py:147-157, (B, D) glue
                mu, var = code.mean(dim=0), code.var(dim=0)
                if bool((var != 0).all()):
                    kl = 0.5 * (-torch.log(var) + mu ** 2 + var - 1).mean() * cfg.VOICE2POSE.GENERATOR.LAMBDA_CLIP_KL
                    losses["G_clipcode_kl_loss"] = kl
                    g_loss = g_loss + kl
            losses["G_loss"] = g_loss
            if cfg.VOICE2POSE.POSE_ENCODER.NAME is not None:                                   # voice2pose.py:162-176
                results["mu_pred"], results["logvar_pred"] = self.pose_encoder(self._fgd_input(pred, batch, dataset))
                results["mu_gt"], results["logvar_gt"] = self.pose_encoder(self._fgd_input(gt32, batch, dataset))
            if hasattr(self, "netD_pose"):
                self._discriminator_losses(losses, g_loss, gt32, pred)
        return losses, results


def _V2PLossFn_last_results(model):
    eng = model.step_engine()
    A = eng.arena
    out = {"kl_applied": A.bufs["kl_out"][1] if "kl_out" in A.bufs else torch.zeros(()),
           "condition_code": A.bufs["code"].clone() if "code" in A.bufs else None}
    pe = getattr(model, "pose_encoder", None)
    if pe is not None and pe._eng is not None:
        b = pe._eng.arena.bufs
        out.update(mu_pred=b["mu/pred"], logvar_pred=b["logvar/pred"], mu_gt=b["mu/gt"], logvar_gt=b["logvar/gt"])
    return out


# ------------------------------------------------------------------------------------------------
# fused train step
# ------------------------------------------------------------------------------------------------
def _cuda_device(device):
    d = torch.device(device)
    if d.type != "cuda":
        raise RuntimeError("the fused trainers run only on CUDA devices (libsdt_b200 has no CPU fallback), got %s" % d)
    return torch.device("cuda", torch.cuda.current_device() if d.index is None else d.index)


def _peer_exchange(trainer, numel):
    """parallel.PeerExchange for SDT_COMM=p2p, or None.  If the symmetric allocation cannot be set up on this software / hardware
    stack the trainer says so and drops to the NCCL all-reduce ("serial")."""
    if trainer.world <= 1 or trainer.comm_mode != "p2p":
        return None
    try:
        px = parallel.PeerExchange(numel, trainer.device, trainer.pg)
        if os.environ.get("SDT_P2P_2GRAPHS"):            # tuning aid: two graphs around an eager exchange instead of one graph
            trainer.comm_mode = "p2p-2graphs"
        return px
    except Exception as e:                      # noqa: BLE001 - no symmetric memory / peer access here: fall back, loudly
        import warnings
        warnings.warn("SDT_COMM=p2p: symmetric memory is not available (%s: %s); falling back to SDT_COMM=serial" % (type(e).__name__, e))
        trainer.comm_mode = "serial"
        return None


def _on_device(fn):
    """Run a trainer method with the trainer's device current: streams, events and ``ops._stream()`` all follow torch's
    current device, so ``Trainer(cfg, n, 'cuda:1')`` must not depend on the caller having called ``torch.cuda.set_device``."""
    import functools

    @functools.wraps(fn)
    def wrapped(self, *args, **kwargs):
        if torch.cuda.current_device() == self.device.index:
            return fn(self, *args, **kwargs)
        with torch.cuda.device(self.device):
            return fn(self, *args, **kwargs)
    return wrapped


class Voice2PoseTrainer:
    """The numeric part of Voice2Pose.train_step (voice2pose.py:281-312) for the SDT configs as a fused device program.

    * parameters of netG (and the dense clips_code table) live in ONE flat fp32 buffer; the nn.Parameters of
      ``self.model`` are views into it, so ``state_dict()`` keeps the reference layout;
    * gradients are written by the backward kernels straight into a flat gradient buffer (no autograd), all-reduced
      with a single NCCL call when world_size > 1, then consumed by a flat fused Adam (K18);
    * after one eager warm-up step the whole step is captured into CUDA graphs and replayed.
    """

    def __init__(self, cfg, num_train_samples, device, use_cuda_graph=True, process_group=None, seed=0, conv_math=None):
        self.cfg = cfg
        self.device = _cuda_device(device)
        with torch.cuda.device(self.device):
            self._init(num_train_samples, use_cuda_graph, process_group, seed, conv_math)

    def _init(self, num_train_samples, use_cuda_graph, process_group, seed, conv_math):
        cfg = self.cfg
        # convolution math mode of THIS trainer's engines (None = the process default, 3 = tcgen05 TF32 unless SDT_CONV_MATH /
        # ops.set_conv_math say otherwise); 0 = fp32 FFMA.  TF32 is the reference's own GPU default (cudnn.allow_tf32)
        self.conv_math = ops.resolve_math(conv_math)
        self.has_d = cfg.VOICE2POSE.POSE_DISCRIMINATOR.NAME is not None
        if self.has_d and cfg.VOICE2POSE.POSE_DISCRIMINATOR.WHITE_LIST is not None:
            raise NotImplementedError("the fused trainer feeds the discriminator all keypoints (POSE_DISCRIMINATOR.WHITE_LIST=None); "
                                      "a white list trains through the drop-in Voice2PoseModel (tests/test_gpu_step.py)")
        torch.manual_seed(seed)                                       # main.py:37
        self.model = Voice2PoseModel(cfg, num_train_samples=num_train_samples).to(self.device)
        self.model.set_conv_math(self.conv_math)
        ae_ckpt = cfg.VOICE2POSE.POSE_ENCODER.AE_CHECKPOINT
        if cfg.VOICE2POSE.POSE_ENCODER.NAME is not None and ae_ckpt is not None:      # voice2pose.py:234-242
            ckpt = torch.load(ae_ckpt, map_location="cpu")
            enc = OrderedDict((k.replace("module.ae.encoder.", ""), v) for k, v in ckpt["model_state_dict"].items() if "encoder" in k)
            self.model.pose_encoder.load_state_dict(enc)
        self.model.train()                                            # trainer.py:382
        self.pg = process_group
        self.world = torch.distributed.get_world_size(process_group) if process_group is not None else 1
        self.use_graph = use_cuda_graph
        self.lr = float(cfg.TRAIN.LR)
        self.code_lr = self.lr * float(cfg.VOICE2POSE.GENERATOR.CLIP_CODE.LR_SCALING)
        m = self.model
        self.g_names = [n for n, _ in m.netG.named_parameters()]
        g_params = [p for _, p in m.netG.named_parameters()]
        self.train_code = isinstance(m.clips_code, nn.Parameter) and m.clips_code.requires_grad
        # flat layout: [netG parameters | pad to a multiple of 4 | clips_code (N*D) | pad | netD_pose parameters | pad]
        self.n_g = sum(p.numel() for p in g_params)
        self.n_g_pad = self.n_g + ((-self.n_g) % 4)
        self.n_code = m.clips_code.numel() if self.train_code else 0
        self.n_code_pad = self.n_code + ((-self.n_code) % 4)
        self.d_names = [n for n, _ in m.netD_pose.named_parameters()] if self.has_d else []
        d_params = [p for _, p in m.netD_pose.named_parameters()] if self.has_d else []
        self.n_d = sum(p.numel() for p in d_params)
        self.off_d = self.n_g_pad + self.n_code_pad
        self.flat_p = torch.zeros(self.off_d + self.n_d + ((-self.n_d) % 4), device=self.device)
        off = 0
        for p in g_params:
            v = self.flat_p[off:off + p.numel()].view(p.shape)
            v.copy_(p.data)
            p.data = v
            off += p.numel()
        if self.train_code:
            v = self.flat_p[self.n_g_pad:self.n_g_pad + self.n_code].view(m.clips_code.shape)
            v.copy_(m.clips_code.data)
            m.clips_code.data = v
        self.comm_mode = os.environ.get("SDT_COMM", _DEFAULT_COMM) if self.world > 1 else "none"
        self._px = _peer_exchange(self, self.flat_p.numel())          # SDT_COMM=p2p: the gradient buffer is a symmetric allocation
        self.flat_g = self._px.flat if self._px is not None else torch.zeros_like(self.flat_p)
        self.exp_avg = torch.zeros_like(self.flat_p)
        self.exp_avg_sq = torch.zeros_like(self.flat_p)
        self.grads = {}
        off = 0
        for n, p in zip(self.g_names, g_params):
            self.grads[n] = self.flat_g[off:off + p.numel()].view(p.shape)
            off += p.numel()
        self.g_table = self.flat_g[self.n_g_pad:self.n_g_pad + self.n_code].view(-1, m.clips_code.shape[1]) if self.train_code else None
        self.d_grads, self.d_scratch = {}, {}
        if self.has_d:
            off = self.off_d
            scratch = torch.zeros(self.n_d, device=self.device)       # parameter gradients of the fake pass of G_loss: discarded
            so = 0                                                    # (the reference zeroes them before D_loss.backward, voice2pose.py:306)
            for n, p in zip(self.d_names, d_params):
                v = self.flat_p[off:off + p.numel()].view(p.shape)
                v.copy_(p.data)
                p.data = v
                self.d_grads[n] = self.flat_g[off:off + p.numel()].view(p.shape)
                self.d_scratch[n] = scratch[so:so + p.numel()].view(p.shape)
                off += p.numel()
                so += p.numel()
        self.adam_g = torch.zeros(8, device=self.device)
        self.adam_c = torch.zeros(8, device=self.device)
        self.adam_d = torch.zeros(8, device=self.device)
        self.set_lr(self.lr)
        self.engine = m.step_engine()
        self._wg_stream = torch.cuda.Stream(device=self.device, priority=int(os.environ.get("SDT_WG_PRIORITY", "0")))   # weight gradients overlap the dgrad chain (engine._wgrad)
        m.netG.engine().wg_stream = self._wg_stream
        m.netG.engine().defer_reduce = os.environ.get("SDT_DEFER_REDUCE", "1") != "0"   # gradient buffers are static: batch the split-K reductions
        self._overlap = True
        # multi-GPU: "overlap" = gradient buckets all-reduced on a communication stream while the backward pass still runs, the
        # clip-code gradient exchanged as B x (index, 32) rows, NCCL captured inside the step's CUDA graph; "serial" = one flat
        # all-reduce between two graphs (round-1 behaviour; also the fallback when NCCL cannot be captured on this stack)
        self._comm = torch.cuda.Stream(device=self.device) if self.world > 1 else None
        self._works = []
        self._aux = None
        self._inbox, self._prefetched = None, None
        self._scal = torch.zeros(12, device=self.device, dtype=torch.float64)
        self._scal_host = torch.zeros(2, 12, dtype=torch.float64).pin_memory()
        self._scal_ev = [torch.cuda.Event(), torch.cuda.Event()]
        self._staging = None
        self._graphs = None
        self._warm = 0
        self.kernels_per_step = 0
        self.steps_done = 0

    @_on_device
    def close(self):
        """Drop the captured graphs and pending collectives (call before tearing the process group down: NCCL communicators
        cannot be destroyed while graphs that hold their kernels are alive)."""
        self._graphs, self._works = None, []
        torch.cuda.synchronize(self.device)

    def set_overlap(self, on):
        """Multi-stream overlap of the step (FGD/metrics and weight gradients beside the dgrad chain). bench.py switches
        it off for the per-kernel roofline pass so that every launch is timed alone."""
        self._overlap = bool(on)
        self.model.netG.engine().wg_stream = self._wg_stream if on else None
        self._graphs = None

    # ---- schedule hook (MultiStepLR steps per epoch in the reference, voice2pose.py:251-257)
    def set_lr(self, lr):
        self.lr = float(lr)
        self.code_lr = self.lr * float(self.cfg.VOICE2POSE.GENERATOR.CLIP_CODE.LR_SCALING)
        self.adam_g[3] = self.lr
        self.adam_c[3] = self.code_lr
        self.adam_d[3] = self.lr                                       # optimizerD_pose, same schedule (voice2pose.py:262-268)

    # ---- host -> device staging
    def _stage(self, batch):
        dev = self.device
        audio, poses, idx = batch["audio"], batch["poses"], batch["clip_index"]
        st = batch["speaker_stat"]
        if self._staging is None or self._staging["audio"].shape != audio.shape or self._staging["poses"].shape != poses.shape:
            self._staging = dict(
                audio=torch.empty(audio.shape, device=dev), poses=torch.empty(poses.shape, device=dev),
                idx=torch.empty(idx.shape, device=dev, dtype=torch.long),
                mean=torch.empty(tuple(st["mean"].shape), device=dev, dtype=torch.float64),
                std=torch.empty(tuple(st["std"].shape), device=dev, dtype=torch.float64),
                scale=torch.empty(tuple(st["scale_factor"].shape), device=dev, dtype=torch.float64))
            self._graphs = None
        s = self._staging
        if batch is self._prefetched:                   # uploaded ahead of time by prefetch(): device -> device hand-over
            main = torch.cuda.current_stream()
            main.wait_event(self._inbox_ready)
            for k in ("audio", "poses", "idx", "mean", "std", "scale"):
                s[k].copy_(self._inbox[k], non_blocking=True)
            self._inbox_free.record(main)
            self._prefetched = None
            return s
        s["audio"].copy_(audio, non_blocking=True)
        s["poses"].copy_(poses, non_blocking=True)
        s["idx"].copy_(idx, non_blocking=True)
        s["mean"].copy_(torch.as_tensor(st["mean"]), non_blocking=True)
        s["std"].copy_(torch.as_tensor(st["std"]), non_blocking=True)
        s["scale"].copy_(torch.as_tensor(st["scale_factor"]), non_blocking=True)
        release = batch.get("_release")
        if release is not None:                         # data.DeviceBatchBuilder: its device buffers may be refilled from here on
            release()
        return s

    @_on_device
    def prefetch(self, batch):
        """Start the host->device copy of a FUTURE batch on a copy stream, overlapping the step in flight.  The next
        ``train_step(batch)`` with this very object picks the device copy up (one inbox: prefetch one batch ahead)."""
        dev = self.device
        st = batch["speaker_stat"]
        src = dict(audio=batch["audio"], poses=batch["poses"], idx=batch["clip_index"], mean=torch.as_tensor(st["mean"]),
                   std=torch.as_tensor(st["std"]), scale=torch.as_tensor(st["scale_factor"]))
        if self._inbox is None or any(self._inbox[k].shape != v.shape for k, v in src.items()):
            dt = dict(audio=torch.float32, poses=torch.float32, idx=torch.long, mean=torch.float64, std=torch.float64, scale=torch.float64)
            self._inbox = {k: torch.empty(tuple(v.shape), device=dev, dtype=dt[k]) for k, v in src.items()}
            self._copy_stream = torch.cuda.Stream(device=self.device)
            self._inbox_ready, self._inbox_free = torch.cuda.Event(), torch.cuda.Event()
            self._inbox_free.record(torch.cuda.current_stream())
        cs = self._copy_stream
        cs.wait_event(self._inbox_free)                 # the previous occupant has been handed over to the step's staging
        if any(v.is_cuda for v in src.values()):
            # device-resident sources (data.DeviceBatchBuilder): their producer (upload + sdt_pose_preprocess) was ordered
            # before the CURRENT stream by builder.batch(); the copy stream has to see that order too
            cs.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(cs):
            for k, v in src.items():
                self._inbox[k].copy_(v, non_blocking=True)
            self._inbox_ready.record(cs)
            release = batch.get("_release")
            if release is not None:
                release(cs)
        self._prefetched = batch

    def set_p2g_stats(self, stat_parted, stat_global):
        """HIERARCHICAL_POSE=False (voice2pose_s2g): the speaker's parted and global statistics for the FGD extractor's input
        transform (dataset.transform_normalized_parted2global, voice2pose.py:168-169).  {'mean': (242), 'std': (242)} each."""
        def f32(v):
            return torch.as_tensor(__import__("numpy").asarray(v, "float64").astype("float32").reshape(-1)).to(self.device)
        self._p2g = (f32(stat_parted["mean"]), f32(stat_parted["std"]), f32(stat_global["mean"]), f32(stat_global["std"]))

    def _gan(self):
        """Discriminator passes + LSGAN terms of the step (voice2pose.py:179-208) and their backward passes.

        score_real = D(real), score_fake = D(fake), score_fake_detach = D(fake.detach()): three forward passes with their own
        BatchNorm batch statistics and running-stat updates, in the reference's order.  G_loss gets MSE(score_fake, 1) * lambda:
        its gradient goes through D to the prediction (the D parameter gradients of that pass are the ones the reference
        zeroes again before D_loss.backward).  D_loss = (MSE(score_real, 1) + MSE(score_fake_detach, 0)) * lambda gives the
        discriminator gradients.  Optimizer order (G step before D_loss.backward, :298-309) does not matter for the values: the
        D passes all ran on the pre-step generator output.  Returns the total gradient w.r.t. the prediction (B,F,2K)."""
        cfg, m, s = self.cfg, self.model, self._staging
        dcfg = cfg.VOICE2POSE.POSE_DISCRIMINATOR
        lam = float(dcfg.LAMBDA_GAN)
        A = self.engine._arena(self.device)
        out = self.out
        B, F = s["poses"].shape[0], s["poses"].shape[1]
        K2 = s["poses"].shape[2] * s["poses"].shape[3]
        pred, gt = out["poses_pred_batch"].view(B, F, K2), s["poses"].view(B, F, K2)
        if dcfg.MOTION:                                                 # frame differences (voice2pose.py:187-189)
            real = ops.motion_diff_fwd(gt, out=A.get("gan_real", (B, F - 1, K2)))
            fake = ops.motion_diff_fwd(pred, out=A.get("gan_fake", (B, F - 1, K2)))
        else:
            real, fake = gt, pred
        D = m.netD_pose.engine()
        dp = {n: p.detach() for n, p in m.netD_pose.named_parameters()}
        dbuf = m.netD_pose._buffers_dict()
        T = real.shape[1]
        D.prepare(dp, T)
        s_real = D.forward(real, dp, dbuf, True, "/real")
        s_fake = D.forward(fake, dp, dbuf, True, "/fake")
        s_fd = D.forward(fake, dp, dbuf, True, "/fake_detach")
        g_fake, g_real, g_fd = (A.get("gan_g_" + k, tuple(s_fake.shape)) for k in ("fake", "real", "fd"))
        gan, d_real, d_fake = A.get("gan_loss", (1,)), A.get("gan_d_real", (1,)), A.get("gan_d_fake", (1,))
        ops.mse_const_loss(s_fake, 1.0, lam, gan, g_fake)               # G_pose_gan_loss
        ops.mse_const_loss(s_real, 1.0, lam, d_real, g_real)
        ops.mse_const_loss(s_fd, 0.0, lam, d_fake, g_fd)
        out["G_pose_gan_loss"] = gan
        out["D_pose_gan_loss"] = torch.add(d_real, d_fake, out=A.get("gan_d_loss", (1,)))
        out["pose_score_fake"] = torch.mean(s_fake, dim=(0, 1), keepdim=False, out=A.get("gan_sf", ()))
        out["pose_score_real"] = torch.mean(s_real, dim=(0, 1), keepdim=False, out=A.get("gan_sr", ()))
        torch.add(out["G_loss"], gan, out=out["G_loss"])                # G_loss = reg (+ kl) + gan
        # generator path: d G_pose_gan_loss / d prediction
        dx = D.backward(g_fake, dp, self.d_scratch, "/fake", True, False)
        g_total = A.get("gan_g_pred", (B, F, K2))
        g_total.copy_(self.engine._g_pred)
        if dcfg.MOTION:
            ops.motion_diff_bwd(dx.view(B, T, K2), out=g_total, accumulate=True)
        else:
            g_total.add_(dx.view(B, F, K2))
        # discriminator path: real, then fake.detach() accumulated on top
        D.backward(g_real, dp, self.d_grads, "/real", False, False)
        D.backward(g_fd, dp, self.d_grads, "/fake_detach", False, True)
        return g_total

    # ---- the device program, in two halves around the all-reduce
    # ---- multi-GPU: gradient buckets in the order the backward pass completes them (still ONE logical exchange of the
    # flat gradient per step, cut where the weight-gradient stream passes two layers so that the wire overlaps the remaining math)
    def _buckets(self):
        """[(engine mark or None, start, end)] over flat_g.  A mark names the layer whose weight gradient completes the bucket."""
        n = self.flat_g.numel()
        if self.has_d or self.comm_mode != "overlap" or not self._overlap:
            return [(None, 0, n)]
        from .engine import ENC_PREFIX
        tail = self.n_g_pad if self.train_code else n          # the dense clip-code gradient never goes on the wire (rows do)
        plan = parallel.bucket_plan([(k, p.numel()) for k, p in self.model.netG.named_parameters()],
                                    ["unet.e0.conv.weight", ENC_PREFIX + "2.0.conv.weight"], n, tail)
        return [(m[:-len(".conv.weight")] if m else None, lo, hi) for m, lo, hi in plan]      # engine marks are layer names

    def _reduce_async(self, ready, lo, hi):
        """all-reduce(sum) of flat_g[lo:hi] on the communication stream, starting when `ready` (event) has happened."""
        with torch.cuda.stream(self._comm):
            self._comm.wait_event(ready)
            self._works.append(torch.distributed.all_reduce(self.flat_g[lo:hi], op=torch.distributed.ReduceOp.SUM, group=self.pg,
                                                            async_op=True))

    def _exchange(self, g_code):
        """Everything of the step that crosses GPUs and is not a marked bucket: the last bucket, the clip-code gradient rows and
        (C5, trainer.py:323-327) the loss / metric scalars.  Ends with the main stream waiting for all of the step's collectives."""
        dist = torch.distributed
        main = torch.cuda.current_stream()
        done = torch.cuda.Event()
        done.record(main)
        A = self.engine._arena(self.device)
        rows_all = idx_all = None
        with torch.cuda.stream(self._comm):
            self._comm.wait_event(done)
            for mark, lo, hi in self._bucket_plan:
                if mark is None:
                    self._works.append(dist.all_reduce(self.flat_g[lo:hi], op=dist.ReduceOp.SUM, group=self.pg, async_op=True))
            if self.train_code and self.comm_mode == "overlap" and not self.has_d:
                B, D = g_code.shape
                rows = A.get("xchg_rows", (B, D))
                torch.add(g_code, A.get("g_code_kl", (B, D)), out=rows)
                rows_all, idx_all, works = parallel.gather_rows(
                    rows, self.engine._clip_index, self.pg, async_op=True,
                    out=(A.get("xchg_rows_all", (self.world * B, D)), A.get("xchg_idx_all", (self.world * B,), torch.long)))
                self._works += works
            self._works.append(dist.all_reduce(self._scal, op=dist.ReduceOp.SUM, group=self.pg, async_op=True))
        for w in self._works:
            w.wait()                                    # the main stream waits; nothing blocks the host
        self._works = []
        main.wait_stream(self._comm)
        if rows_all is not None:                        # dense clip-code gradient = sum over ranks of the scattered rows (K12)
            ops.code_scatter_grad(rows_all, None, idx_all, self.g_table)

    def _fwd_bwd(self):
        s = self._staging
        if self.train_code:
            self.g_table.zero_()                                       # optimizerClipCode.zero_grad(); dense grad (K12)
        p2g = getattr(self, "_p2g", None)
        multi = self.world > 1 and self.comm_mode == "overlap"
        sparse_code = multi and self.train_code and not self.has_d
        eng = self.model.netG.engine()
        if multi:
            self._bucket_plan = self._buckets()
            ranges = {m: (lo, hi) for m, lo, hi in self._bucket_plan if m is not None}
            eng.grad_marks = {m: None for m in ranges}
            eng.on_mark = lambda name, ev: self._reduce_async(ev, *ranges[name])
        else:
            eng.grad_marks, eng.on_mark = None, None
        if not self._overlap:
            self.out = self.engine.forward(s["audio"], s["poses"], s["idx"], (s["mean"], s["std"], s["scale"]), p2g_stats=p2g)
            g_code = self.engine.backward(self.grads, None if sparse_code else self.g_table, g_pred=self._gan() if self.has_d else None)
            self._pack_scalars()
            if multi:
                self._exchange(g_code)
            return
        self.out = self.engine.forward(s["audio"], s["poses"], s["idx"], (s["mean"], s["std"], s["scale"]), p2g_stats=p2g, defer_side=True)
        # fork: FGD features + f64 results/metrics on a second stream while the backward pass runs on this one
        main = torch.cuda.current_stream()
        if self._aux is None:
            self._aux = torch.cuda.Stream(device=self.device)
        fork, join = torch.cuda.Event(), torch.cuda.Event()
        fork.record(main)
        with torch.cuda.stream(self._aux):
            self._aux.wait_event(fork)
            if not _DIAG_SKIP_SIDE:                            # diagnostic only: cost of the FGD / metrics side stream
                self.engine.run_side()
            join.record(self._aux)
        g_code = self.engine.backward(self.grads, None if sparse_code else self.g_table, g_pred=self._gan() if self.has_d else None)
        main.wait_event(join)
        self._pack_scalars()
        if multi:
            self._exchange(g_code)

    def _optim(self):
        gs = 1.0 / self.world
        ops.adam_advance(self.adam_g, -1.0)
        ops.adam_flat(self.flat_p[:self.n_g_pad], self.flat_g[:self.n_g_pad], self.exp_avg[:self.n_g_pad],
                      self.exp_avg_sq[:self.n_g_pad], self.adam_g, grad_scale=gs,
                      weight_decay=float(self.cfg.TRAIN.WD))            # optimizerG only (voice2pose.py:249-250)
        if self.train_code:
            ops.adam_advance(self.adam_c, -1.0)
            sl = slice(self.n_g_pad, self.n_g_pad + self.n_code_pad)
            ops.adam_flat(self.flat_p[sl], self.flat_g[sl], self.exp_avg[sl], self.exp_avg_sq[sl], self.adam_c, grad_scale=gs)
        if self.has_d:                                                 # optimizerD_pose (voice2pose.py:305-309)
            ops.adam_advance(self.adam_d, -1.0)
            sl = slice(self.off_d, self.flat_p.numel())
            ops.adam_flat(self.flat_p[sl], self.flat_g[sl], self.exp_avg[sl], self.exp_avg_sq[sl], self.adam_d, grad_scale=gs)

    def _allreduce(self):
        """comm_mode "serial": ONE flat NCCL all-reduce of the whole gradient buffer between the two graphs + the scalars (C5).
        comm_mode "p2p": the same exchange through peer memory (parallel.PeerExchange / csrc/p2p.cu), no NCCL in the step."""
        if self.world > 1 and self.comm_mode.startswith("p2p"):
            self._px.allreduce(self._scal)
        elif self.world > 1 and self.comm_mode != "overlap":
            parallel.allreduce_flat_(self.flat_g, self.pg)
            torch.distributed.all_reduce(self._scal, group=self.pg)

    def _step(self):
        """The whole device program of one step in stream order (collectives included in comm_mode "overlap")."""
        self._fwd_bwd()
        self._allreduce()
        self._optim()

    @_on_device
    def run_staged(self):
        """Run one step on the already-staged device batch (bench.py's device-resident timing)."""
        if self.use_graph and self._graphs is None and self._warm >= 2:
            self._capture()
        if self._graphs is not None:
            self._graphs[0].replay()
            if len(self._graphs) == 2:
                self._allreduce()
                self._graphs[1].replay()
        else:
            n0 = _lib.launch_count
            self._step()
            self.kernels_per_step = _lib.launch_count - n0
            self._warm += 1
        self.steps_done += 1
        return self.out

    @_on_device
    def _capture(self):
        """world 1 and comm_mode "overlap": ONE graph holds the step (NCCL collectives are captured with it, on the streams
        torch's process group gives them).  comm_mode "serial": two graphs around the eager all-reduce.  If capturing NCCL fails
        on this software stack the trainer drops to "serial" and says so."""
        torch.cuda.synchronize()
        if self.world > 1:
            torch.distributed.barrier(group=self.pg)
            torch.cuda.synchronize()
        # Stream priorities were measured and rejected (profiles/r2_ablation_stream_priority.txt): capturing the critical path
        # (forward, data-gradient chain, Adam) on a high-priority stream makes the step 0.27 ms SLOWER -- the weight gradients then
        # pile up behind it and run alone at the end.  SDT_MAIN_PRIORITY / SDT_WG_PRIORITY remain as tuning aids (default 0 = equal).
        side = torch.cuda.Stream(device=self.device, priority=int(os.environ.get("SDT_MAIN_PRIORITY", "0")))
        side.wait_stream(torch.cuda.current_stream())
        if self.world == 1 or self.comm_mode in ("overlap", "p2p"):       # p2p: barriers + the exchange kernel are plain launches
            try:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.stream(side):
                    with torch.cuda.graph(g, stream=side, capture_error_mode="thread_local"):
                        self._step()
                self._graphs = (g,)
            except Exception as e:                      # noqa: BLE001 - NCCL capture unsupported: fall back, loudly
                if self.world == 1:
                    raise
                import warnings
                warnings.warn("capturing the collectives inside the step graph failed (%s: %s); falling back to two graphs around an eager exchange"
                              % (type(e).__name__, e))
                self.comm_mode, self._works = ("p2p-2graphs" if self.comm_mode == "p2p" else "serial"), []
                torch.cuda.synchronize()
        if self._graphs is None:
            g1, g2 = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
            with torch.cuda.stream(side):
                with torch.cuda.graph(g1, stream=side):
                    self._fwd_bwd()
                with torch.cuda.graph(g2, stream=side):
                    self._optim()
            self._graphs = (g1, g2)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()

    @_on_device
    def train_step(self, batch):
        """batch: the reference's batch dict (host tensors; pinned memory recommended). Returns a dict of device
        tensors: G_reg_loss, [G_clipcode_kl_loss], G_loss, L2_dist, lip_sync_error_n, poses_pred_batch, mu_*/logvar_*."""
        self._stage(batch)
        return self.run_staged()

    _SCALARS = ("G_reg_loss", "G_loss", "L2_dist", "lip_sync_error_n", "G_clipcode_kl_loss", "kl_applied",
                "G_pose_gan_loss", "D_pose_gan_loss", "pose_score_fake", "pose_score_real")

    def _pack_scalars(self):
        """Last node of the step: the loss / metric scalars as one f64 vector (a single 64-byte D2H read per step)."""
        for i, k in enumerate(self._SCALARS):
            if k in self.out:
                self._scal[i:i + 1].copy_(self.out[k])

    def _scalars_dict(self, vals):
        if self.world > 1:              # the all-reduced sums -> means over ranks (reduce_tensor_dict, trainer.py:323-327)
            vals = [v / self.world for v in vals]
        d = {k: v for k, v in zip(self._SCALARS, vals) if k in self.out}
        if "kl_applied" in d and d.pop("kl_applied") == 0.0:          # the reference's guard (voice2pose.py:154)
            d.pop("G_clipcode_kl_loss", None)
        return d

    @_on_device
    def losses_to_host(self, out=None):
        """One small D2H read of the last step's scalars (the reference logs these every LOG_INTERVAL steps). Blocking."""
        return self._scalars_dict(self._scal.cpu().tolist())

    @_on_device
    def post_losses(self, slot):
        """Asynchronous variant: enqueue the D2H of the last step's scalars into pinned slot 0/1; collect_losses(slot)
        waits for exactly that copy, so the host can run one step ahead of the device."""
        self._scal_host[slot].copy_(self._scal, non_blocking=True)
        self._scal_ev[slot].record(torch.cuda.current_stream())

    def collect_losses(self, slot):
        self._scal_ev[slot].synchronize()
        return self._scalars_dict(self._scal_host[slot].tolist())

    @_on_device
    def run_epoch(self, batches, on_losses=None):
        """The reference's inner loop (trainer.py: ``for batch in dataloader: train_step; log``) as a software pipeline:
        batch k+1 is uploaded on a copy stream while step k runs, and the scalars of step k are read while step k+1 is
        already enqueued.  ``on_losses(step_index, dict)`` is called for EVERY step, in order.  Returns the step count."""
        it = iter(batches)
        nxt = next(it, None)
        if nxt is not None:
            self.prefetch(nxt)
        k, pending = 0, None
        while nxt is not None:
            self.train_step(nxt)
            nxt = next(it, None)
            if nxt is not None:
                self.prefetch(nxt)
            self.post_losses(k & 1)
            if pending is not None and on_losses is not None:
                on_losses(pending, self.collect_losses(pending & 1))
            pending = k
            k += 1
        if pending is not None and on_losses is not None:
            on_losses(pending, self.collect_losses(pending & 1))
        return k


# ------------------------------------------------------------------------------------------------
# Pose2Pose (pose VAE): drop-in step model + fused train step
# ------------------------------------------------------------------------------------------------
class Pose2PoseModel(nn.Module):
    """core/pipelines/pose2pose.py:20-89.  State-dict groups: buffers ``clip_code_mu`` /
    ``clip_code_logvar`` (N, CODE_DIM), ``mel_transfm.*``, ``ae.encoder.*``, ``ae.decoder.*``.

    The reference computes a mel spectrogram that ``Autoencoder.forward`` never reads (dead compute, pose2pose.py:48,65);
    it is not computed here."""

    def __init__(self, cfg, state_dict=None, num_train_samples=None, rank=0):
        super().__init__()
        self.cfg = cfg
        self.mel_transfm = MelSpectrogram()
        self.ae = get_model(cfg.POSE2POSE.AUTOENCODER.NAME)(cfg)
        if num_train_samples is None:
            assert state_dict is not None, "No state_dict available, while no dataset is configured."
            num_train_samples = state_dict["module.clip_code_mu"].shape[0]
        d = cfg.POSE2POSE.AUTOENCODER.CODE_DIM
        self.register_buffer("clip_code_mu", torch.zeros([num_train_samples, d]))
        self.register_buffer("clip_code_logvar", torch.zeros([num_train_samples, d]))

    def forward(self, batch, return_loss=True, is_testing=False, interpolation_coeff=None):
        cfg = self.cfg
        num_frames = int(batch["num_frames"][0].item())
        if not return_loss:                                                                        # pose2pose.py:52-65
            assert cfg.DEMO.CODE_PATH is not None
            import numpy as np
            idx = int((cfg.DEMO.MULTIPLE - 1) * interpolation_coeff)
            code = np.load(cfg.DEMO.CODE_PATH)["v"][idx] * 10                                        # "10 is empirically selected"
            code = torch.Tensor(code).cuda().unsqueeze(0)
            pred, mu, logvar = self.ae(None, cfg.DATASET.NUM_FRAMES, external_code=code)
            return {"poses_pred_batch": pred, "clip_code_mu": mu, "clip_code_logvar": logvar}
        poses_gt = batch["poses"].cuda()
        pred, mu, logvar = self.ae(poses_gt, num_frames, None)
        losses = OrderedDict()
        reg = (torch.abs(pred - poses_gt) * cfg.POSE2POSE.LAMBDA_REG).mean()                       # pose2pose.py:71-73
        losses["reg_loss"] = reg
        kl = 0.5 * (-logvar + mu ** 2 + torch.exp(logvar) - 1).mean() * cfg.POSE2POSE.LAMBDA_KL     # pose2pose.py:77
        losses["kl_loss"] = kl
        losses["loss"] = reg + kl
        results = {"poses_pred_batch": pred, "poses_gt_batch": poses_gt, "clip_code_mu": mu, "clip_code_logvar": logvar}
        return losses, results


class Pose2PoseTrainer:
    """The numeric part of Pose2Pose.train_step (pose2pose.py:124-150) as a fused device program: VAE forward, L1 + KL,
    f64 final results + metrics, clip-code buffer scatter, backward, (NCCL all-reduce), flat Adam over ``ae``."""

    def __init__(self, cfg, num_train_samples, device, use_cuda_graph=True, process_group=None, seed=0, conv_math=None):
        self.cfg = cfg
        self.device = _cuda_device(device)
        with torch.cuda.device(self.device):
            self._init(num_train_samples, use_cuda_graph, process_group, seed, conv_math)

    def _init(self, num_train_samples, use_cuda_graph, process_group, seed, conv_math):
        cfg = self.cfg
        self.conv_math = ops.resolve_math(conv_math)
        torch.manual_seed(seed)
        self.model = Pose2PoseModel(cfg, num_train_samples=num_train_samples).to(self.device)
        self.model.ae.set_conv_math(self.conv_math)
        self.model.train()
        self.pg = process_group
        self.world = torch.distributed.get_world_size(process_group) if process_group is not None else 1
        self.use_graph = use_cuda_graph
        ae = self.model.ae
        self.names = [n for n, _ in ae.named_parameters()]
        params = [p for _, p in ae.named_parameters()]
        n = sum(p.numel() for p in params)
        self.flat_p = torch.zeros(n + ((-n) % 4), device=self.device)
        self.comm_mode = os.environ.get("SDT_COMM", _DEFAULT_COMM) if self.world > 1 else "none"
        self._px = _peer_exchange(self, self.flat_p.numel())
        self.flat_g = self._px.flat if self._px is not None else torch.zeros_like(self.flat_p)
        self.exp_avg = torch.zeros_like(self.flat_p)
        self.exp_avg_sq = torch.zeros_like(self.flat_p)
        self.grads, off = {}, 0
        for nme, p in zip(self.names, params):
            v = self.flat_p[off:off + p.numel()].view(p.shape)
            v.copy_(p.data)
            p.data = v
            self.grads[nme] = self.flat_g[off:off + p.numel()].view(p.shape)
            off += p.numel()
        self.adam = torch.zeros(8, device=self.device)
        self.set_lr(float(cfg.TRAIN.LR))
        from .engine import Arena
        self.arena = Arena(self.device)
        self._staging, self._graphs, self._warm = None, None, 0
        self.kernels_per_step = 0
        self.eps_override = None          # tests inject the N(0,1) draw here
        self._comm = torch.cuda.Stream(device=self.device) if self.world > 1 else None
        self._works, self._reduce_scalars = [], False
        self._scal = torch.zeros(len(self._SCALARS), device=self.device, dtype=torch.float64)
        self._dec_off = sum(p.numel() for nme, p in zip(self.names, params) if nme.startswith("encoder."))

    def close(self):
        self._graphs, self._works = None, []
        torch.cuda.synchronize(self.device)

    def set_lr(self, lr):
        self.lr = float(lr)
        self.adam[3] = self.lr

    def _stage(self, batch):
        dev = self.device
        poses, idx, st = batch["poses"], batch["clip_index"], batch["speaker_stat"]
        if self._staging is None or self._staging["poses"].shape != poses.shape:
            self._staging = dict(
                poses=torch.empty(poses.shape, device=dev), idx=torch.empty(idx.shape, device=dev, dtype=torch.long),
                mean=torch.empty(tuple(st["mean"].shape), device=dev, dtype=torch.float64),
                std=torch.empty(tuple(st["std"].shape), device=dev, dtype=torch.float64),
                scale=torch.empty(tuple(st["scale_factor"].shape), device=dev, dtype=torch.float64),
                eps=torch.empty(poses.shape[0], self.cfg.POSE2POSE.AUTOENCODER.CODE_DIM, device=dev))
            self._graphs = None
        s = self._staging
        s["poses"].copy_(poses, non_blocking=True)
        s["idx"].copy_(idx, non_blocking=True)
        s["mean"].copy_(torch.as_tensor(st["mean"]), non_blocking=True)
        s["std"].copy_(torch.as_tensor(st["std"]), non_blocking=True)
        s["scale"].copy_(torch.as_tensor(st["scale_factor"]), non_blocking=True)
        return s

    def _fwd_bwd(self):
        s, cfg, A = self._staging, self.cfg, self.arena
        m = self.model
        ae = m.ae
        eng = ae.engine()
        B, F = s["poses"].shape[0], s["poses"].shape[1]
        K2 = ae.n_landmarks * 2
        if self.eps_override is not None:
            s["eps"].copy_(self.eps_override)
        else:
            s["eps"].normal_()                                           # torch.randn(logvar.shape) (autoencoder.py:86)
        params = {n: p.detach() for n, p in ae.named_parameters()}
        poses = s["poses"].view(B, F, K2)
        pred, mu, logvar, kl = eng.forward(poses, s["eps"], params, ae._buffers_dict(), True, float(cfg.POSE2POSE.LAMBDA_KL))
        reg = A.get("reg_out", (1,))
        g_pred = A.get("g_pred", (B, F, K2))
        ops.l1_loss(pred, poses, float(cfg.POSE2POSE.LAMBDA_REG), reg, g_pred, A.get("l1_partial", (1024,)))
        loss = A.get("loss", (1,))
        torch.add(reg, kl, out=loss)
        hier = bool(cfg.DATASET.HIERARCHICAL_POSE)
        fp = ops.pose_final_results(pred.view(B, F, 2, -1), s["mean"], s["std"], s["scale"], hier, out=A.get("final_pred", (B, F, 2, K2 // 2), torch.float64))
        fg = ops.pose_final_results(s["poses"], s["mean"], s["std"], s["scale"], hier, out=A.get("final_gt", (B, F, 2, K2 // 2), torch.float64))
        met = ops.pose_metrics(fp, fg, A.get("met_partial", (2 * B,), torch.float64), A.get("met_out", (2,), torch.float64))
        ops.code_store_rows(mu, m.clip_code_mu, s["idx"], logvar, m.clip_code_logvar)      # pose2pose.py:135-137
        eng.backward(g_pred, self.grads, include_kl=True)
        self.out = OrderedDict(reg_loss=reg, kl_loss=kl, loss=loss, L2_dist=met[0:1], lip_sync_error_n=met[1:2],
                               poses_pred_batch=pred.view(B, F, 2, -1), clip_code_mu=mu, clip_code_logvar=logvar,
                               final_pred=fp, final_gt=fg)

    def _optim(self):
        ops.adam_advance(self.adam, -1.0)
        ops.adam_flat(self.flat_p, self.flat_g, self.exp_avg, self.exp_avg_sq, self.adam, grad_scale=1.0 / self.world,
                      weight_decay=float(self.cfg.TRAIN.WD))            # pose2pose.py:114-115

    # ---- multi-GPU (pose2pose.py:101-102: DDP over `ae`): two gradient buckets, decoder first (its gradients are final when the
    # backward pass enters the encoder), all-reduced on a communication stream inside the step's graph
    def _reduce_async(self, lo, hi):
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream())
        with torch.cuda.stream(self._comm):
            self._comm.wait_event(ev)
            self._works.append(torch.distributed.all_reduce(self.flat_g[lo:hi], op=torch.distributed.ReduceOp.SUM, group=self.pg,
                                                            async_op=True))

    def _step(self):
        multi = self.world > 1 and self.comm_mode == "overlap"
        eng = self.model.ae.engine()
        eng.on_decoder_done = (lambda: self._reduce_async(self._dec_off, self.flat_g.numel())) if multi else None
        self._fwd_bwd()
        if multi:
            self._reduce_async(0, self._dec_off)
            self._scal.copy_(torch.cat([self.out[k].double().view(1) for k in self._SCALARS]))
            self._reduce_scalars = True
            with torch.cuda.stream(self._comm):
                self._works.append(torch.distributed.all_reduce(self._scal, group=self.pg, async_op=True))
            for w in self._works:
                w.wait()
            self._works = []
            torch.cuda.current_stream().wait_stream(self._comm)
        elif self.world > 1:
            self._allreduce()
        self._optim()

    def _allreduce(self):
        if self.comm_mode.startswith("p2p"):
            self._scal.copy_(torch.cat([self.out[k].double().view(1) for k in self._SCALARS]))
            self._reduce_scalars = True
            self._px.allreduce(self._scal)
        else:
            parallel.allreduce_flat_(self.flat_g, self.pg)

    @_on_device
    def run_staged(self):
        if self.use_graph and self._graphs is None and self._warm >= 2:
            torch.cuda.synchronize()
            if self.world > 1:
                torch.distributed.barrier(group=self.pg)
                torch.cuda.synchronize()
            side = torch.cuda.Stream(device=self.device)
            side.wait_stream(torch.cuda.current_stream())
            if self.world == 1 or self.comm_mode in ("overlap", "p2p"):       # p2p: barriers + the exchange kernel are plain launches
                try:
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.stream(side):
                        with torch.cuda.graph(g, stream=side, capture_error_mode="thread_local"):
                            self._step()
                    self._graphs = (g,)
                except Exception as e:                      # noqa: BLE001
                    if self.world == 1:
                        raise
                    import warnings
                    warnings.warn("capturing the collectives inside the step graph failed (%s: %s); falling back to two graphs around an eager exchange"
                                  % (type(e).__name__, e))
                    self.comm_mode, self._works = ("p2p-2graphs" if self.comm_mode == "p2p" else "serial"), []
                    torch.cuda.synchronize()
            if self._graphs is None:
                g1, g2 = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
                with torch.cuda.stream(side):
                    with torch.cuda.graph(g1, stream=side):
                        self._fwd_bwd()
                    with torch.cuda.graph(g2, stream=side):
                        self._optim()
                self._graphs = (g1, g2)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
        if self._graphs is not None:
            self._graphs[0].replay()
            if len(self._graphs) == 2:
                if self.world > 1:
                    self._allreduce()
                self._graphs[1].replay()
        else:
            n0 = _lib.launch_count
            self._step()
            self.kernels_per_step = _lib.launch_count - n0
            self._warm += 1
        return self.out

    @_on_device
    def train_step(self, batch):
        self._stage(batch)
        return self.run_staged()

    _SCALARS = ("reg_loss", "kl_loss", "loss", "L2_dist", "lip_sync_error_n")

    @_on_device
    def losses_to_host(self, out):
        """One small D2H read of the step's scalars; with several ranks in comm_mode "overlap" these are the means over ranks
        (the all-reduce rode along with the gradient exchange: reduce_tensor_dict, trainer.py:323-327)."""
        if self._reduce_scalars:
            return dict(zip(self._SCALARS, (self._scal / self.world).cpu().tolist()))
        vals = torch.cat([out[k].double().view(1) for k in self._SCALARS]).cpu().tolist()
        return dict(zip(self._SCALARS, vals))
