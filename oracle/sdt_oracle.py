"""CPU oracle for the Voice2Pose / Pose2Pose training-step hot path.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it; the product package
(``speechdrivestemplates_b200``) never does and fails loudly when its CUDA library is missing.

It is a plain restatement, on CPU tensors (torch CPU ops / numpy, fp32 by default, fp64 on request),
of what the reference computes on this path.  The arithmetic of the reference lives in un-vendored
third-party dependencies (``torch==1.7.0``, ``torchaudio==0.7.0`` per /root/reference/requirements.txt:8-9;
2.11.0 / 2.11.0 in this image), so each function restates the *published algorithm* of the library
call at the cited reference call site.

Parity pin: the reference ships no tests or golden vectors.  The oracle is pinned
against the reference ITSELF: ``tests/golden/make_golden.py`` imports the unmodified reference from
/root/reference (CPU) and records its outputs in ``tests/golden/*.npz``; ``tests/test_oracle_golden.py``
checks this file against those fixtures.

Layouts here follow the reference (NCHW / NCL), not the product's channels-last HBM layout.
"""
import math
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

EPS_NORM = 1e-5

# --------------------------------------------------------------------------------------------
# configuration (mirrors the keys of /root/reference/configs/default.py that matter on the path)
# --------------------------------------------------------------------------------------------

def make_cfg(name="voice2pose_sdt_bp", **over):
    """Return a plain dict with the hot-path subset of the reference's config tree.

    name: one of the reference's YAML overlays (configs/*.yaml).
    """
    cfg = dict(
        pipeline="Voice2Pose",
        g_norm="IN", g_leaky=True,                 # default.py:11-12
        lambda_reg=1.0, lambda_clip_kl=0.1,        # default.py:13-14
        code_dim=None, code_train=True, external_code=False, code_lr_scaling=1.0,  # default.py:16-23
        pose_encoder=True,                         # default.py:26
        disc=False, d_leaky=False, lambda_gan=1.0, d_motion=True,  # default.py:30-34
        ae_leaky=True, ae_norm="BN", ae_code_dim=32, p2p_lambda_reg=1.0, p2p_lambda_kl=0.1,  # default.py:37-43
        n_landmarks=121, hierarchical=True, num_frames=64,  # default.py:49-52
        lr=1e-4, wd=0.0,                           # default.py:64-65
    )
    if name == "voice2pose_sdt_bp":      # configs/voice2pose_sdt_bp.yaml
        cfg.update(code_dim=32, external_code=False)
    elif name == "voice2pose_sdt_vae":   # configs/voice2pose_sdt_vae.yaml
        cfg.update(code_dim=32, external_code=True)
    elif name == "voice2pose_s2g":       # configs/voice2pose_s2g.yaml
        cfg.update(g_norm="BN", disc=True, lambda_gan=0.1, d_leaky=True, hierarchical=False)
    elif name == "pose2pose":            # configs/pose2pose.yaml
        cfg.update(pipeline="Pose2Pose")
    else:
        raise KeyError("Unknown config: %s" % name)
    cfg.update(over)
    return cfg


# (state-dict suffix, Cout, Cin, (kh, kw), stride, pad)  -- generator.py:15-30
ENC2D_LAYERS = [
    ("0.0", 64, 1, (3, 3), 1, 1),
    ("0.1", 64, 64, (4, 4), 2, 1),
    ("1.0", 128, 64, (3, 3), 1, 1),
    ("1.1", 128, 128, (4, 4), 2, 1),
    ("2.0", 256, 128, (3, 3), 1, 1),
    ("2.1", 256, 256, (4, 4), 2, 1),
    ("3.0", 256, 256, (3, 3), 1, 1),
    ("3.1", 256, 256, (6, 3), 1, 0),
]
UNET_ENC = ["e0", "e1", "e2", "e3", "e4", "e5", "e6"]   # generator.py:53-62 (e2..e6 downsample)
UNET_DEC = ["d5", "d4", "d3", "d2", "d1"]               # generator.py:64-68


# --------------------------------------------------------------------------------------------
# parameter construction in the reference's RNG order
# --------------------------------------------------------------------------------------------

def _conv_weight(shape, kaiming_normal):
    """nn.ConvNd default init (kaiming_uniform, a=sqrt(5)) then building_blocks.py:44 kaiming_normal_."""
    w = torch.empty(shape)
    torch.nn.init.kaiming_uniform_(w, a=math.sqrt(5))
    if kaiming_normal:
        torch.nn.init.kaiming_normal_(w)
    return w


def _conv_bias(wshape):
    fan_in = int(np.prod(wshape[1:]))
    bound = 1.0 / math.sqrt(fan_in)
    b = torch.empty(wshape[0])
    torch.nn.init.uniform_(b, -bound, bound)
    return b


def _add_block(sd, prefix, wshape, norm):
    """One ConvNormRelu (building_blocks.py:4-46): bias-free conv (+ BN affine/buffers when norm=='BN')."""
    sd[prefix + ".conv.weight"] = _conv_weight(wshape, True)
    if norm == "BN":
        c = wshape[0]
        sd[prefix + ".norm.weight"] = torch.ones(c)
        sd[prefix + ".norm.bias"] = torch.zeros(c)
        sd[prefix + ".norm.running_mean"] = torch.zeros(c)
        sd[prefix + ".norm.running_var"] = torch.ones(c)
        sd[prefix + ".norm.num_batches_tracked"] = torch.tensor(0, dtype=torch.long)
    elif norm != "IN":
        raise NotImplementedError(norm)


def init_generator(cfg, sd=None, prefix="netG."):
    """SequenceGeneratorCNN.__init__ (generator.py:87-104), consuming torch's RNG in the same order."""
    sd = OrderedDict() if sd is None else sd
    norm = cfg["g_norm"]
    for suf, co, ci, (kh, kw), _s, _p in ENC2D_LAYERS:
        _add_block(sd, prefix + "audio_encoder.specgram_encoder_2d." + suf, (co, ci, kh, kw), norm)
    cin0 = 256 + (cfg["code_dim"] or 0)
    _add_block(sd, prefix + "unet.e0", (256, cin0, 3), norm)
    _add_block(sd, prefix + "unet.e1", (256, 256, 3), norm)
    for n in UNET_ENC[2:]:
        _add_block(sd, prefix + "unet." + n, (256, 256, 4), norm)
    for n in UNET_DEC:
        _add_block(sd, prefix + "unet." + n, (256, 256, 3), norm)
    for i in range(4):
        _add_block(sd, prefix + "decoder.%d" % i, (256, 256, 3), norm)
    wshape = (cfg["n_landmarks"] * 2, 256, 1)
    sd[prefix + "decoder.4.weight"] = _conv_weight(wshape, False)
    sd[prefix + "decoder.4.bias"] = _conv_bias(wshape)
    return sd


def init_pose_encoder(cfg, sd=None, prefix="pose_encoder."):
    """PoseSeqEncoder.__init__ (autoencoder.py:8-25)."""
    sd = OrderedDict() if sd is None else sd
    norm = cfg["ae_norm"]
    cin = cfg["n_landmarks"] * 2
    shapes = [(256, cin, 3), (256, 256, 3)] + [(256, 256, 4)] * 4 + [(cfg["ae_code_dim"] * 2, 256, 4)]
    for i, s in enumerate(shapes):
        _add_block(sd, prefix + "blocks.%d" % i, s, norm)
    return sd


def init_pose_decoder(cfg, sd=None, prefix="decoder."):
    """PoseSeqDecoder.__init__ (autoencoder.py:37-57)."""
    sd = OrderedDict() if sd is None else sd
    norm = cfg["ae_norm"]
    _add_block(sd, prefix + "d5", (256, cfg["ae_code_dim"], 3), norm)
    for n in ["d4", "d3", "d2", "d1"]:
        _add_block(sd, prefix + n, (256, 256, 3), norm)
    for i in range(4):
        _add_block(sd, prefix + "blocks.%d" % i, (256, 256, 3), norm)
    wshape = (cfg["n_landmarks"] * 2, 256, 1)
    sd[prefix + "blocks.4.weight"] = _conv_weight(wshape, False)
    sd[prefix + "blocks.4.bias"] = _conv_bias(wshape)
    return sd


def init_discriminator(cfg, sd=None, prefix="netD_pose."):
    """PoseSequenceDiscriminator.__init__ (discriminator.py:6-17); norm is always BN (ConvNormRelu default)."""
    sd = OrderedDict() if sd is None else sd
    cin = cfg["n_landmarks"] * 2
    _add_block(sd, prefix + "seq.0", (256, cin, 4), "BN")
    _add_block(sd, prefix + "seq.1", (512, 256, 4), "BN")
    _add_block(sd, prefix + "seq.2", (1024, 512, 3), "BN")
    wshape = (1, 1024, 3)
    sd[prefix + "seq.3.weight"] = _conv_weight(wshape, False)
    sd[prefix + "seq.3.bias"] = _conv_bias(wshape)
    return sd


def init_voice2pose(cfg, num_train_samples, seed=0):
    """Voice2PoseModel.__init__ (voice2pose.py:22-82) under torch.manual_seed(seed) (main.py:37)."""
    torch.manual_seed(seed)
    sd = OrderedDict()
    if cfg["code_dim"] is not None and not cfg["external_code"]:
        sd["clips_code"] = torch.zeros(num_train_samples, cfg["code_dim"])  # voice2pose.py:63-70
    sd["mel_transfm.spectrogram.window"] = hann_window_periodic(400).float()
    sd["mel_transfm.mel_scale.fb"] = mel_filterbank().float()
    # NB: state-dict order puts clips_code first, but construction order is netG, pose_encoder, netD.
    init_generator(cfg, sd)
    if cfg["pose_encoder"]:
        init_pose_encoder(cfg, sd)
    if cfg["disc"]:
        init_discriminator(cfg, sd)
    return sd


def init_pose2pose(cfg, num_train_samples, seed=0):
    """Pose2PoseModel.__init__ (pose2pose.py:20-39)."""
    torch.manual_seed(seed)
    sd = OrderedDict()
    sd["clip_code_mu"] = torch.zeros(num_train_samples, cfg["ae_code_dim"])
    sd["clip_code_logvar"] = torch.zeros(num_train_samples, cfg["ae_code_dim"])
    sd["mel_transfm.spectrogram.window"] = hann_window_periodic(400).float()
    sd["mel_transfm.mel_scale.fb"] = mel_filterbank().float()
    init_pose_encoder(cfg, sd, "ae.encoder.")
    init_pose_decoder(cfg, sd, "ae.decoder.")
    return sd


# --------------------------------------------------------------------------------------------
# mel front end  (torchaudio.transforms.MelSpectrogram as configured at voice2pose.py:27-30)
# --------------------------------------------------------------------------------------------
SAMPLE_RATE = 16000
N_FFT, WIN_LENGTH, HOP, N_MELS = 512, 400, 160, 80
F_MIN, F_MAX = 55.0, 7500.0
N_FREQS = N_FFT // 2 + 1


def hann_window_periodic(n=WIN_LENGTH):
    """torch.hann_window(n, periodic=True) = 0.5 - 0.5 cos(2 pi k / n).

    The reference's buffer ``mel_transfm.spectrogram.window`` is the fp32 evaluation by torch itself (its fp32 cos
    differs from the exactly-rounded formula by ~7e-9), so the oracle takes the library value to stay bit-identical.
    """
    return torch.hann_window(n, periodic=True, dtype=torch.float32)


def _hz_to_mel_htk(f):
    return 2595.0 * math.log10(1.0 + f / 700.0)


def mel_filterbank():
    """torchaudio.functional.melscale_fbanks(257, 55, 7500, 80, 16000, norm=None, 'htk') -> (257, 80) fp32.

    Triangular filters on the HTK mel scale, no area normalisation.  Evaluated in fp32 like
    torchaudio does, which reproduces the reference's ``mel_transfm.mel_scale.fb`` buffer bit for bit.
    """
    all_freqs = torch.linspace(0, SAMPLE_RATE // 2, N_FREQS)
    m_pts = torch.linspace(_hz_to_mel_htk(F_MIN), _hz_to_mel_htk(F_MAX), N_MELS + 2)
    f_pts = 700.0 * (10.0 ** (m_pts / 2595.0) - 1.0)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)          # (257, 82)
    down = -slopes[:, :-2] / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    return torch.clamp(torch.minimum(down, up), min=0.0)


def num_mel_frames(n_samples):
    return 1 + n_samples // HOP


def mel_spectrogram(audio, window=None, fb=None, dtype=torch.float32):
    """Power mel spectrogram. audio (B, L) -> (B, 80, 1 + L//160).  No log.

    center=True reflect-pad by n_fft//2, hann(400) centred in the 512 frame (56 zeros each side),
    one-sided rFFT, |X|^2, then P^T . fb.
    """
    audio = audio.to(dtype)
    w = (hann_window_periodic() if window is None else window).to(device=audio.device, dtype=dtype)
    fbm = (mel_filterbank() if fb is None else fb).to(device=audio.device, dtype=dtype)
    lpad = (N_FFT - WIN_LENGTH) // 2
    wfull = torch.zeros(N_FFT, dtype=dtype, device=audio.device)
    wfull[lpad:lpad + WIN_LENGTH] = w
    x = F.pad(audio.unsqueeze(1), (N_FFT // 2, N_FFT // 2), mode="reflect").squeeze(1)
    frames = x.unfold(-1, N_FFT, HOP)                     # (B, T, 512)
    spec = torch.fft.rfft(frames * wfull, dim=-1)         # (B, T, 257)
    power = spec.real ** 2 + spec.imag ** 2
    return torch.matmul(power, fbm).transpose(1, 2).contiguous()


def parse_audio_length(audio_length, sr, fps):
    """core/utils/audio_processing.py:5-11 (float arithmetic kept as is)."""
    bit_per_frames = sr / fps
    num_frames = int(audio_length / bit_per_frames)
    audio_length = int(num_frames * bit_per_frames)
    return audio_length, num_frames


def crop_pad_audio(wav, audio_length):
    """core/utils/audio_processing.py:14-19."""
    if len(wav) > audio_length:
        wav = wav[:audio_length]
    elif len(wav) < audio_length:
        wav = np.pad(wav, [0, audio_length - len(wav)], mode="constant", constant_values=0)
    return wav


# --------------------------------------------------------------------------------------------
# building block  (core/networks/building_blocks.py:4-55)
# --------------------------------------------------------------------------------------------

def _act(x, leaky):
    """LeakyReLU(0.2) / ReLU (building_blocks.py:46).  A float selects an arbitrary negative slope: the tests use
    slope 1.0 (identity) to compare gradients without the activation's derivative discontinuity."""
    if isinstance(leaky, float):
        return F.leaky_relu(x, leaky)
    return F.leaky_relu(x, 0.2) if leaky else F.relu(x)


def _batch_norm(x, sd, prefix, training, momentum=0.1):
    """BatchNorm{1,2}d: batch stats (biased var) in training, running stats in eval; running_var unbiased."""
    dims = [0] + list(range(2, x.dim()))
    shape = [1, -1] + [1] * (x.dim() - 2)
    g, b = sd[prefix + ".norm.weight"], sd[prefix + ".norm.bias"]
    if training:
        mean = x.mean(dim=dims)
        var = x.var(dim=dims, unbiased=False)
        n = x.numel() // x.shape[1]
        with torch.no_grad():
            rm, rv = sd[prefix + ".norm.running_mean"], sd[prefix + ".norm.running_var"]
            rm.mul_(1 - momentum).add_(momentum * mean.detach().to(rm.dtype))
            rv.mul_(1 - momentum).add_(momentum * (var.detach() * (n / max(n - 1, 1))).to(rv.dtype))
            sd[prefix + ".norm.num_batches_tracked"] += 1
    else:
        mean = sd[prefix + ".norm.running_mean"].to(x.dtype)
        var = sd[prefix + ".norm.running_var"].to(x.dtype)
    xh = (x - mean.view(shape)) / torch.sqrt(var.view(shape) + EPS_NORM)
    return xh * g.to(x.dtype).view(shape) + b.to(x.dtype).view(shape)


def conv_norm_relu(x, sd, prefix, norm, leaky, stride, pad, training=True):
    """ConvNormRelu.forward (building_blocks.py:48-55).

    'IN' on a 1-D block normalises over CHANNELS per (b, t) (permute trick at :50-51), on a 2-D block
    over H*W per (b, c); no affine, no running stats.  'BN' is per-channel over batch+space.
    """
    w = sd[prefix + ".conv.weight"].to(x.dtype)
    if w.dim() == 4:
        x = F.conv2d(x, w, None, stride, pad)
    else:
        x = F.conv1d(x, w, None, stride, pad)
    if norm == "IN":
        dims = (1,) if w.dim() == 3 else (2, 3)
        mean = x.mean(dim=dims, keepdim=True)
        var = x.var(dim=dims, unbiased=False, keepdim=True)
        x = (x - mean) / torch.sqrt(var + EPS_NORM)
    elif norm == "BN":
        x = _batch_norm(x, sd, prefix, training)
    else:
        raise NotImplementedError(norm)
    return _act(x, leaky)


# --------------------------------------------------------------------------------------------
# generator  (core/networks/keypoints_generation/generator.py)
# --------------------------------------------------------------------------------------------

def interp_linear_x2_to(x, size):
    """F.interpolate(x, size, mode='linear'), align_corners=False (generator.py:79-83)."""
    return F.interpolate(x, size, mode="linear", align_corners=False)


def audio_encoder_forward(mel, sd, cfg, num_frames, training=True, prefix="netG.", taps=None):
    """AudioEncoder.forward (generator.py:39-43): 8 Conv2d blocks, bilinear resize to (1, F), squeeze."""
    x = mel.unsqueeze(1)
    for suf, _co, _ci, _k, s, p in ENC2D_LAYERS:
        x = conv_norm_relu(x, sd, prefix + "audio_encoder.specgram_encoder_2d." + suf,
                           cfg["g_norm"], cfg["g_leaky"], s, p, training)
        if taps is not None:
            taps["enc." + suf] = x
    x = F.interpolate(x, (1, num_frames), mode="bilinear", align_corners=False)
    return x.squeeze(2)


def unet_forward(x, sd, cfg, training=True, prefix="netG.unet.", taps=None):
    """UNet_1D.forward (generator.py:70-85)."""
    norm, leaky = cfg["g_norm"], cfg["g_leaky"]
    e = []
    for i, n in enumerate(UNET_ENC):
        down = i >= 2
        x = conv_norm_relu(x, sd, prefix + n, norm, leaky, 2 if down else 1, 1, training)
        e.append(x)
        if taps is not None:
            taps["unet." + n] = x
    d = e[6]
    for j, n in enumerate(UNET_DEC):
        skip = e[5 - j]
        d = conv_norm_relu(interp_linear_x2_to(d, skip.size(-1)) + skip, sd, prefix + n, norm, leaky, 1, 1, training)
        if taps is not None:
            taps["unet." + n] = d
    return d


def generator_forward(mel, num_frames, code, sd, cfg, training=True, prefix="netG.", taps=None):
    """SequenceGeneratorCNN.forward (generator.py:106-117) -> (B, F, 2, K)."""
    x = audio_encoder_forward(mel, sd, cfg, num_frames, training, prefix, taps)
    if taps is not None:
        taps["enc.out"] = x
    if cfg["code_dim"] is not None:
        x = torch.cat([x, code.to(x.dtype).unsqueeze(2).repeat(1, 1, x.shape[-1])], 1)
    x = unet_forward(x, sd, cfg, training, prefix + "unet.", taps)
    for i in range(4):
        x = conv_norm_relu(x, sd, prefix + "decoder.%d" % i, cfg["g_norm"], cfg["g_leaky"], 1, 1, training)
        if taps is not None:
            taps["dec.%d" % i] = x
    x = F.conv1d(x, sd[prefix + "decoder.4.weight"].to(x.dtype), sd[prefix + "decoder.4.bias"].to(x.dtype))
    return x.permute(0, 2, 1).reshape(-1, num_frames, 2, cfg["n_landmarks"])


# --------------------------------------------------------------------------------------------
# pose VAE + discriminator  (autoencoder.py, discriminator.py)
# --------------------------------------------------------------------------------------------

def pose_encoder_forward(poses, sd, cfg, training, prefix="pose_encoder.", taps=None):
    """PoseSeqEncoder.forward (autoencoder.py:27-35): (B,F,2,K) -> mu, logvar (B, code)."""
    x = poses.reshape(poses.shape[0], poses.shape[1], -1).permute(0, 2, 1)
    for i in range(7):
        x = conv_norm_relu(x, sd, prefix + "blocks.%d" % i, cfg["ae_norm"], cfg["ae_leaky"],
                           2 if i >= 2 else 1, 1, training)
        if taps is not None:
            taps["penc.%d" % i] = x
    x = x[..., 0]                      # F.interpolate(x, 1) == nearest, index 0
    return x[:, 0::2], x[:, 1::2]


def pose_decoder_forward(code, sd, cfg, training, prefix="decoder.", taps=None):
    """PoseSeqDecoder.forward (autoencoder.py:59-69): (B, code) -> (B, 2K, 64)."""
    x = code.unsqueeze(-1).repeat(1, 1, 2)                 # nearest x2 of a length-1 signal
    for n in ["d5", "d4", "d3", "d2", "d1"]:
        x = conv_norm_relu(interp_linear_x2_to(x, x.shape[-1] * 2), sd, prefix + n,
                           cfg["ae_norm"], cfg["ae_leaky"], 1, 1, training)
        if taps is not None:
            taps["pdec." + n] = x
    for i in range(4):
        x = conv_norm_relu(x, sd, prefix + "blocks.%d" % i, cfg["ae_norm"], cfg["ae_leaky"], 1, 1, training)
    return F.conv1d(x, sd[prefix + "blocks.4.weight"].to(x.dtype), sd[prefix + "blocks.4.bias"].to(x.dtype))


def autoencoder_forward(poses, num_frames, sd, cfg, eps, training=True, prefix="ae."):
    """Autoencoder.forward (autoencoder.py:79-92) with the N(0,1) draw ``eps`` injected (SURVEY §7 hard part 5)."""
    mu, logvar = pose_encoder_forward(poses, sd, cfg, training, prefix + "encoder.")
    code = mu + torch.exp(0.5 * logvar) * eps.to(mu.dtype)
    x = pose_decoder_forward(code, sd, cfg, training, prefix + "decoder.")
    x = x.permute(0, 2, 1).reshape(-1, num_frames, 2, cfg["n_landmarks"])
    return x, mu, logvar


def discriminator_forward(x, sd, cfg, training=True, prefix="netD_pose."):
    """PoseSequenceDiscriminator.forward (discriminator.py:19-23): (B,T,2,K) -> (B,T')."""
    x = x.reshape(x.size(0), x.size(1), -1).transpose(1, 2)
    x = conv_norm_relu(x, sd, prefix + "seq.0", "BN", cfg["d_leaky"], 2, 1, training)
    x = conv_norm_relu(x, sd, prefix + "seq.1", "BN", cfg["d_leaky"], 2, 1, training)
    x = conv_norm_relu(x, sd, prefix + "seq.2", "BN", cfg["d_leaky"], 1, 1, training)
    x = F.conv1d(x, sd[prefix + "seq.3.weight"].to(x.dtype), sd[prefix + "seq.3.bias"].to(x.dtype), 1, 1)
    return x.squeeze(1)


# --------------------------------------------------------------------------------------------
# keypoint indexing / normalisation  (core/datasets/gesture_dataset.py:131-220) -- bit-exact gates
# --------------------------------------------------------------------------------------------
ROOT_NODE, HAND_ROOT_L, HAND_ROOT_R, HEAD_ROOT = 1, 6, 3, 39     # gesture_dataset.py:42-45
IDX_137_TO_122 = list(range(0, 8)) + [15, 16] + list(range(25, 137))   # :134
IDX_122_TO_121 = [0] + list(range(2, 122))                             # :143
HEAD_SET = list(range(9, HEAD_ROOT)) + list(range(HEAD_ROOT + 1, 79))  # :149 / :159


def remove_unused_kp(poses):
    assert poses.shape[-1] == 137
    return poses[..., :, IDX_137_TO_122]


def absolute_to_relative(poses):
    poses = poses.copy()
    poses[..., :2, :] = poses[..., :2, :] - poses[..., :2, ROOT_NODE, None]
    return poses[..., :, IDX_122_TO_121]


def global_to_parted(poses):
    poses = poses.copy()
    poses[..., :2, HEAD_SET] = poses[..., :2, HEAD_SET] - poses[..., :2, HEAD_ROOT, None]
    poses[..., :2, 79:100] = poses[..., :2, 79:100] - poses[..., :2, HAND_ROOT_L, None]
    poses[..., :2, 100:121] = poses[..., :2, 100:121] - poses[..., :2, HAND_ROOT_R, None]
    return poses


def parted_to_global(poses):
    poses = poses.copy()
    poses[..., :2, HEAD_SET] = poses[..., :2, HEAD_SET] + poses[..., :2, HEAD_ROOT, None]
    poses[..., :2, 79:100] = poses[..., :2, 79:100] + poses[..., :2, HAND_ROOT_L, None]
    poses[..., :2, 100:121] = poses[..., :2, 100:121] + poses[..., :2, HAND_ROOT_R, None]
    return poses


def normalize_poses(kp, mean, std):
    """(kp - f32(mean)) / f32(std), IEEE division; kp (T,2,K) f32, stats (2K,) (gesture_dataset.py:173-191)."""
    k = kp.shape[-1]
    m = np.asarray(mean, np.float64).astype(np.float32).reshape(1, 2, k)
    s = np.asarray(std, np.float64).astype(np.float32).reshape(1, 2, k)
    return (kp.astype(np.float32) - m) / s


def preprocess_pose(raw, mean, std, hierarchical=True):
    """Dataset pose preprocessing (gesture_dataset.py:95-105): raw (T,3,137) f32 -> (T,2,121) f32."""
    p = remove_unused_kp(np.asarray(raw, np.float32))
    p = absolute_to_relative(p)
    if hierarchical:
        p = global_to_parted(p)
    return normalize_poses(p[:, :2, :], mean, std)


def get_final_results(poses, mean, std, scale, hierarchical=True):
    """gesture_dataset.py:213-220: f64 ``x*std + mean`` (two roundings) -> parted_to_global -> ``* scale``.

    poses (B,T,2,K) f32; mean/std (B,2K) f64; scale (B,) f64.  Returns f64.
    """
    b, _t, _two, k = poses.shape
    m = np.asarray(mean, np.float64).reshape(b, 1, 2, k)
    s = np.asarray(std, np.float64).reshape(b, 1, 2, k)
    x = poses.astype(np.float64) * s
    x = x + m
    if hierarchical:
        x = parted_to_global(x)
    return x * np.asarray(scale, np.float64).reshape(b, 1, 1, 1)


def transform_normalized_parted2global(poses, stat_parted, stat_global):
    """gesture_dataset.py:221-234 (FGD input prep when HIERARCHICAL_POSE is False), fp32 throughout.

    poses (B,T,2,K) torch f32; stats: dicts of (2K,) arrays for the batch's (single) speaker.
    """
    k = poses.shape[-1]
    f32 = lambda a: torch.from_numpy(np.asarray(a, np.float64).astype(np.float32)).reshape(1, 2, k)
    x = poses.float() * f32(stat_parted["std"]) + f32(stat_parted["mean"])
    x = torch.from_numpy(parted_to_global(x.numpy()))
    return ((x - f32(stat_global["mean"])) / f32(stat_global["std"])).to(poses.dtype)


def evaluate_step(pred, gt):
    """Voice2Pose.evaluate_step (voice2pose.py:412-430) on final-result poses (f64)."""
    d = pred - gt
    l2 = np.sqrt((d * d).sum(axis=2))
    lp = pred[:, :, :, 75] - pred[:, :, :, 71]
    lg = gt[:, :, :, 75] - gt[:, :, :, 71]
    lip_pred = np.sqrt((lp * lp).sum(-1))
    lip_gt = np.sqrt((lg * lg).sum(-1))
    den = lip_gt.max(-1, keepdims=True) + 1e-4
    return {"L2_dist": l2.mean(), "lip_sync_error_n": np.abs(lip_pred / den - lip_gt / den).mean()}


# --------------------------------------------------------------------------------------------
# losses + train step  (voice2pose.py:84-210, 281-309; pose2pose.py:41-89, 124-147)
# --------------------------------------------------------------------------------------------

def clip_code_kl(code, lam):
    """voice2pose.py:147-157.  Returns None when any batch variance is exactly 0 (the reference's guard)."""
    mu = code.mean(dim=0)
    var = code.var(dim=0)              # unbiased
    if not bool((var != 0).all()):
        return None
    return 0.5 * (-torch.log(var) + mu ** 2 + var - 1).mean() * lam


def adam_update(p, g, m, v, step, lr, beta1=0.9, beta2=0.999, eps=1e-8):
    """torch.optim.Adam, single-tensor form; in place on p, m, v. wd == 0 (default.py:65)."""
    m.mul_(beta1).add_(g, alpha=1 - beta1)
    v.mul_(beta2).addcmul_(g, g, value=1 - beta2)
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
    p.addcdiv_(m, denom, value=-(lr / bc1))


class Voice2PoseOracle:
    """State + train step of the Voice2Pose pipeline (model, three Adam optimizers) on CPU."""

    def __init__(self, cfg, num_train_samples, seed=0, sd=None, dtype=torch.float32):
        self.cfg = cfg
        self.dtype = dtype
        self.sd = init_voice2pose(cfg, num_train_samples, seed) if sd is None else sd
        if dtype != torch.float32:
            for k, t in self.sd.items():
                if t.is_floating_point():
                    self.sd[k] = t.to(dtype)
        self.step = 0
        self.g_names = [k for k in self.sd if k.startswith("netG.") and _is_param(k)]
        self.d_names = [k for k in self.sd if k.startswith("netD_pose.") and _is_param(k)]
        self.code_trainable = "clips_code" in self.sd and cfg["code_train"]
        self.adam = {}

    def _adam_group(self, names, grads, lr):
        for n in names:
            g = grads.get(n)
            if g is None:
                continue
            st = self.adam.setdefault(n, dict(m=torch.zeros_like(self.sd[n]), v=torch.zeros_like(self.sd[n]), t=0))
            st["t"] += 1
            adam_update(self.sd[n], g, st["m"], st["v"], st["t"], lr)

    def forward(self, batch, taps=None, training=True, code=None):
        """Voice2PoseModel.forward with return_loss=True (voice2pose.py:84-210).

        training=True: ``model.train()`` -- clip codes gathered from the table (:94), BatchNorm layers on batch statistics.
        training=False: ``model.eval()`` (Voice2Pose.test_step, :333-352) -- BatchNorm from running statistics; the code is
        ``code`` if given, else mu of the ground-truth poses when cfg['test_with_gt_code'] (:100-106); the reference's other
        eval-time selections are random draws and have to be passed in as ``code``.  The loss terms are the same either way."""
        cfg, sd = self.cfg, self.sd
        audio = batch["audio"].to(self.dtype)
        gt = batch["poses"].to(self.dtype)
        nf = int(batch["num_frames"][0])

        def fgd_input(x):
            if cfg["hierarchical"]:
                return x
            return transform_normalized_parted2global(x, batch["stat_parted"], batch["stat_global"])

        if cfg["code_dim"] is not None and code is None:
            if training:
                table = batch["external_code_table"] if cfg["external_code"] else sd["clips_code"]
                code = table[batch["clip_index"]].to(self.dtype)               # voice2pose.py:94
            elif cfg.get("test_with_gt_code"):
                with torch.no_grad():
                    code, _ = pose_encoder_forward(fgd_input(gt), sd, cfg, False)
            else:
                raise ValueError("eval forward: pass the (randomly selected) condition code explicitly")
        mel = mel_spectrogram(audio, sd["mel_transfm.spectrogram.window"], sd["mel_transfm.mel_scale.fb"], self.dtype)
        if taps is not None:
            taps["mel"] = mel
        pred = generator_forward(mel, nf, code, sd, cfg, training, "netG.", taps)
        losses = OrderedDict()
        reg = (torch.abs(pred - gt) * cfg["lambda_reg"]).mean()           # voice2pose.py:141-142
        losses["G_reg_loss"] = reg
        g_loss = reg.clone()
        if code is not None:
            kl = clip_code_kl(code, cfg["lambda_clip_kl"])
            if kl is not None:
                losses["G_clipcode_kl_loss"] = kl
                g_loss = g_loss + kl
        losses["G_loss"] = g_loss
        results = {"poses_pred_batch": pred, "poses_gt_batch": gt, "condition_code": code}
        if cfg["pose_encoder"]:
            with torch.no_grad():       # voice2pose.py:162-176; BN follows the model's mode (train in a train step)
                results["mu_pred"], results["logvar_pred"] = pose_encoder_forward(fgd_input(pred.detach()), sd, cfg, training)
                results["mu_gt"], results["logvar_gt"] = pose_encoder_forward(fgd_input(gt), sd, cfg, training)
        if cfg["disc"]:                 # voice2pose.py:179-208
            real, fake = gt, pred
            if cfg["d_motion"]:
                real = real[:, 1:] - real[:, :-1]
                fake = fake[:, 1:] - fake[:, :-1]
            s_real = discriminator_forward(real, sd, cfg, training)
            s_fake = discriminator_forward(fake, sd, cfg, training)
            s_fake_d = discriminator_forward(fake.detach(), sd, cfg, training)
            g_gan = F.mse_loss(s_fake, torch.ones_like(s_fake)) * cfg["lambda_gan"]
            losses["G_pose_gan_loss"] = g_gan
            losses["G_loss"] = g_loss + g_gan
            d_loss = (F.mse_loss(s_real, torch.ones_like(s_real)) +
                      F.mse_loss(s_fake_d, torch.zeros_like(s_fake_d))) * cfg["lambda_gan"]
            losses["D_pose_gan_loss"] = d_loss
            losses["pose_score_fake"] = s_fake.mean()
            losses["pose_score_real"] = s_real.mean()
        return losses, results

    def train_step(self, batch, taps=None, world_size=1):
        """Voice2Pose.train_step's numeric part (voice2pose.py:288-309). Returns (losses, results, grads)."""
        cfg, sd = self.cfg, self.sd
        leaves = list(self.g_names) + (["clips_code"] if self.code_trainable else []) + list(self.d_names)
        for n in leaves:
            sd[n].requires_grad_(True)
            sd[n].grad = None
        losses, results = self.forward(batch, taps)
        stat = batch["speaker_stat"]
        fin_pred = get_final_results(results["poses_pred_batch"].detach().float().numpy(), stat["mean"], stat["std"],
                                     stat["scale_factor"], cfg["hierarchical"])
        fin_gt = get_final_results(results["poses_gt_batch"].detach().float().numpy(), stat["mean"], stat["std"],
                                   stat["scale_factor"], cfg["hierarchical"])
        metrics = evaluate_step(fin_pred, fin_gt)
        g_leaves = [sd[n] for n in leaves if not n.startswith("netD_pose.")]
        g_grads = torch.autograd.grad(losses["G_loss"], g_leaves, retain_graph=True, allow_unused=True)
        grads = {n: g for n, g in zip([n for n in leaves if not n.startswith("netD_pose.")], g_grads)}
        if cfg["disc"]:
            d_grads = torch.autograd.grad(losses["D_pose_gan_loss"], [sd[n] for n in self.d_names])
            grads.update({n: g for n, g in zip(self.d_names, d_grads)})
        for n in leaves:
            sd[n].requires_grad_(False)
        if self.code_trainable and grads.get("clips_code") is None:
            grads["clips_code"] = torch.zeros_like(sd["clips_code"])
        if world_size > 1:
            grads = {n: (g / world_size if g is not None else None) for n, g in grads.items()}
        self.last_grads = grads
        return ({k: v.detach() for k, v in losses.items()},
                dict(results, final_pred=fin_pred, final_gt=fin_gt, **metrics), grads)

    def apply_optimizers(self, grads):
        """optimizerClipCode.step(); optimizerG.step(); optimizerD_pose.step() (voice2pose.py:302-309)."""
        cfg = self.cfg
        with torch.no_grad():
            if self.code_trainable:
                self._adam_group(["clips_code"], grads, cfg["lr"] * cfg["code_lr_scaling"])
            self._adam_group(self.g_names, grads, cfg["lr"])
            if cfg["disc"]:
                self._adam_group(self.d_names, grads, cfg["lr"])
        self.step += 1


class Pose2PoseOracle:
    """State + train step of the Pose2Pose (pose VAE) pipeline on CPU (pose2pose.py:41-89,124-147)."""

    def __init__(self, cfg, num_train_samples, seed=0, sd=None, dtype=torch.float32):
        self.cfg, self.dtype = cfg, dtype
        self.sd = init_pose2pose(cfg, num_train_samples, seed) if sd is None else sd
        if dtype != torch.float32:
            for k, t in self.sd.items():
                if t.is_floating_point():
                    self.sd[k] = t.to(dtype)
        self.names = [k for k in self.sd if k.startswith("ae.") and _is_param(k)]
        self.adam = {}

    def train_step(self, batch, eps):
        cfg, sd = self.cfg, self.sd
        for n in self.names:
            sd[n].requires_grad_(True)
        gt = batch["poses"].to(self.dtype)
        nf = int(batch["num_frames"][0])
        pred, mu, logvar = autoencoder_forward(gt, nf, sd, cfg, eps, True, "ae.")
        reg = (torch.abs(pred - gt) * cfg["p2p_lambda_reg"]).mean()
        kl = 0.5 * (-logvar + mu ** 2 + torch.exp(logvar) - 1).mean() * cfg["p2p_lambda_kl"]
        loss = reg + kl
        gs = torch.autograd.grad(loss, [sd[n] for n in self.names])
        grads = dict(zip(self.names, gs))
        for n in self.names:
            sd[n].requires_grad_(False)
        with torch.no_grad():
            sd["clip_code_mu"][batch["clip_index"]] = mu.detach().to(sd["clip_code_mu"].dtype)       # pose2pose.py:135-137
            sd["clip_code_logvar"][batch["clip_index"]] = logvar.detach().to(sd["clip_code_logvar"].dtype)
        losses = OrderedDict(reg_loss=reg.detach(), kl_loss=kl.detach(), loss=loss.detach())
        return losses, {"poses_pred_batch": pred.detach(), "clip_code_mu": mu.detach(), "clip_code_logvar": logvar.detach()}, grads

    def apply_optimizers(self, grads):
        with torch.no_grad():
            for n in self.names:
                st = self.adam.setdefault(n, dict(m=torch.zeros_like(self.sd[n]), v=torch.zeros_like(self.sd[n]), t=0))
                st["t"] += 1
                adam_update(self.sd[n], grads[n], st["m"], st["v"], st["t"], self.cfg["lr"])


def _is_param(key):
    return not (key.endswith("running_mean") or key.endswith("running_var") or key.endswith("num_batches_tracked"))


# --------------------------------------------------------------------------------------------
# synthetic inputs  
# --------------------------------------------------------------------------------------------

def synthetic_batch(batch_size, num_train_samples, speaker_stat, seed=1, audio_len=68266, num_frames=64, k=121,
                    stat_parted=None, stat_global=None):
    """Device-independent seeded batch with the reference's batch-dict schema (gesture_dataset.py:107-119)."""
    g = torch.Generator().manual_seed(seed)
    audio = 0.1 * torch.randn(batch_size, audio_len, generator=g)
    poses = torch.randn(batch_size, num_frames, 2, k, generator=g)
    idx = torch.randint(0, num_train_samples, (batch_size,), generator=g)
    return {
        "audio": audio, "poses": poses, "clip_index": idx,
        "num_frames": torch.full((batch_size,), num_frames, dtype=torch.long),
        "speaker": ["oliver"] * batch_size,
        "stat_parted": stat_parted, "stat_global": stat_global,
        "speaker_stat": {
            "mean": np.tile(np.asarray(speaker_stat["mean"], np.float64)[None], (batch_size, 1)),
            "std": np.tile(np.asarray(speaker_stat["std"], np.float64)[None], (batch_size, 1)),
            "scale_factor": np.full((batch_size,), float(speaker_stat["scale_factor"]), np.float64),
        },
    }
