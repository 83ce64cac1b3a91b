"""Pin oracle/sdt_oracle.py against fixtures recorded from the unmodified reference (tests/golden/make_golden.py)."""
import numpy as np
import pytest
import torch

from oracle import sdt_oracle as O
from util import golden, oliver_stat, rel_err, samples_of, tensor_digest


def _check_digests(g, prefix, tensors, rtol, atol=0.0, what="", outlier_frac=0.0):
    """Compare sampled elements + |x| digest of each tensor with the fixture.

    Error is measured against the tensor's RMS.  ``outlier_frac`` > 0 is for steps after the first: the
    gradient field is discontinuous (LeakyReLU masks, sign() of the L1 loss), so a 1e-6 weight difference
    left by the previous Adam step may flip a unit and move a few individual gradient entries by percents.
    """
    n = 0
    n_bad = n_tot = 0
    for k, v in tensors.items():
        key = "%s/%s/samples" % (prefix, k)
        if key not in g.files or v is None:
            continue
        ref = g[key]
        got = samples_of(v)
        scale = float(np.sqrt(g["%s/%s/digest" % (prefix, k)][1] / max(v.numel(), 1))) + 1e-30   # rms of the tensor
        errs = np.abs(got.astype(np.float64) - ref.astype(np.float64))
        n_bad += int((errs > rtol * scale + atol).sum())
        n_tot += errs.size
        if outlier_frac == 0.0:
            assert errs.max() <= rtol * scale + atol, "%s %s: sample err %.3e (rms %.3e)" % (what, k, errs.max(), scale)
        d_ref = g["%s/%s/digest" % (prefix, k)]
        d_got = tensor_digest(v)
        assert abs(d_got[2] - d_ref[2]) <= 10 * rtol * abs(d_ref[2]) + atol * v.numel(), "%s %s: |x| digest" % (what, k)
        n += 1
    assert n > 0
    assert n_bad <= outlier_frac * n_tot, "%s: %d of %d samples out of tolerance" % (what, n_bad, n_tot)


def test_mel_filterbank_and_window_bit_exact():
    g = golden("mel_golden")
    assert np.array_equal(O.hann_window_periodic().float().numpy(), g["window"])
    assert np.array_equal(O.mel_filterbank().float().numpy(), g["fb"])
    fb = g["fb"]
    assert (fb != 0).sum() == 468 and (fb != 0).sum(0).max() <= 15


@pytest.mark.parametrize("case", ["real", "syn", "short", "odd"])
def test_mel_matches_reference(case):
    g = golden("mel_golden")
    gen = torch.Generator().manual_seed(int(g["syn_seed"]))
    syn = 0.1 * torch.randn(2, 68266, generator=gen)
    short = 0.1 * torch.randn(1, 4000, generator=gen)
    odd = 0.1 * torch.randn(1, 1601, generator=gen)
    audio = {"real": torch.from_numpy(g["real_audio"]), "syn": syn, "short": short, "odd": odd}[case]
    ref = g[case + "_mel"]
    got32 = O.mel_spectrogram(audio).numpy()
    got64 = O.mel_spectrogram(audio, dtype=torch.float64).numpy()
    assert got32.shape == ref.shape == (audio.shape[0], 80, 1 + audio.shape[1] // 160)
    # tolerance stated in SURVEY §4: abs <= 1e-5 * max
    assert np.abs(got32 - ref).max() <= 1e-5 * np.abs(ref).max()
    assert np.abs(got64 - ref).max() <= 1e-5 * np.abs(ref).max()


def test_audio_length():
    assert O.parse_audio_length(68267, 16000, 15) == (68266, 64)
    assert O.num_mel_frames(68266) == 427
    assert len(O.crop_pad_audio(np.zeros(10), 16)) == 16 and len(O.crop_pad_audio(np.zeros(20), 16)) == 16


@pytest.mark.parametrize("parted", [True, False])
def test_keypoints_bit_exact(parted):
    g = golden("keypoints_golden")
    st = oliver_stat(parted)
    tag = "parted" if parted else "global"
    got = O.preprocess_pose(g["raw"], st["mean"], st["std"], hierarchical=parted)
    assert got.dtype == np.float32 and got.shape == (8, 2, 121)
    assert np.array_equal(got, g[tag + "_normalized"])
    x = g[tag + "_final_in"]
    b = x.shape[0]
    fin = O.get_final_results(x, np.tile(st["mean"][None], (b, 1)), np.tile(st["std"][None], (b, 1)),
                              np.full((b,), st["scale_factor"]), hierarchical=parted)
    assert fin.dtype == np.float64
    assert np.array_equal(fin, g[tag + "_final_out"])


def test_init_matches_reference_rng_order():
    g = golden("sdt_bp_zero_code_golden")
    sd = O.init_voice2pose(O.make_cfg("voice2pose_sdt_bp"), int(g["n_train"]), seed=0)
    for k, v in sd.items():
        assert np.array_equal(samples_of(v), g["init/%s/samples" % k]), k
    assert len([k for k in g.files if k.startswith("init/") and k.endswith("/digest")]) == len(sd)


def _run_v2p(config, fixture, live_code, grad_rtol0=1e-3, later_grads=True):
    g = golden(fixture)
    cfg = O.make_cfg(config)
    n_train, bs = int(g["n_train"]), int(g["batch_size"])
    orc = O.Voice2PoseOracle(cfg, n_train, seed=0)
    if live_code:
        gen = torch.Generator().manual_seed(11)
        orc.sd["clips_code"] = 0.1 * torch.randn(n_train, 32, generator=gen)
    stat = oliver_stat(cfg["hierarchical"])
    for s in range(int(g["steps"])):
        batch = O.synthetic_batch(bs, n_train, stat, seed=100 + s, stat_parted=oliver_stat(True),
                                  stat_global=oliver_stat(False))
        losses, results, grads = orc.train_step(batch)
        p = "step%d" % s
        # After the first Adam step (p -= lr*g/|g| at t=1) every fp32-noise-level gradient difference becomes a
        # weight difference of up to 2*lr, so later steps are compared at a looser forward tolerance.
        ftol = 1e-4 if s == 0 else 5e-3
        for k, v in losses.items():
            assert abs(float(v) - float(g["%s/loss/%s" % (p, k)])) <= ftol * max(1.0, abs(float(v))), k
        assert ("G_clipcode_kl_loss" in losses) == (("%s/loss/G_clipcode_kl_loss" % p) in g.files)
        assert rel_err(results["poses_pred_batch"].detach().numpy(), g[p + "/pred"]) < ftol
        assert rel_err(results["final_pred"], g[p + "/final_pred"]) < ftol
        for k in ("L2_dist", "lip_sync_error_n"):
            assert abs(results[k] - float(g["%s/loss/%s" % (p, k)])) <= ftol * abs(results[k])
        if cfg["pose_encoder"]:
            assert rel_err(results["mu_pred"].numpy(), g[p + "/mu_pred"]) < 10 * ftol
            assert rel_err(results["mu_gt"].numpy(), g[p + "/mu_gt"]) < 10 * ftol
        if s == 0:
            _check_digests(g, p + "/grad", grads, rtol=grad_rtol0, what="grad")
        elif later_grads:   # one flipped LeakyReLU unit moves every upstream gradient by ~1e-3 rms (see _check_digests)
            _check_digests(g, p + "/grad", grads, rtol=3e-2, what="grad", outlier_frac=0.02)
        orc.apply_optimizers(grads)
        # Adam amplifies tiny grad differences where |g| ~ eps; lr = 1e-4 bounds the per-step difference
        _check_digests(g, p + "/state", orc.sd, rtol=1e-4, atol=2.5e-4 * (s + 1), what="state")


def test_sdt_bp_step_matches_reference():
    _run_v2p("voice2pose_sdt_bp", "sdt_bp_step_golden", True)


def test_sdt_bp_zero_code_skips_kl():
    _run_v2p("voice2pose_sdt_bp", "sdt_bp_zero_code_golden", False)


def test_s2g_step_matches_reference():
    # BN over a batch of 2: the fp32 noise floor of the early-layer gradients is ~2e-2 of their rms (fp32 vs fp64
    # oracle), and the explicit BN here rounds differently from ATen's fused kernel -> 3e-2.
    # Step-1 gradients of this B=2 BatchNorm net are chaotic at the 1e-1 level (measured: fp32 vs fp64 oracle), so
    # only losses / predictions / states are pinned there.
    _run_v2p("voice2pose_s2g", "s2g_step_golden", False, grad_rtol0=3e-2, later_grads=False)


def test_s2g_forward_parity_gate():
    """BASELINE.json configs[0]: s2g generator forward, 1 clip, eval-mode BN."""
    g = golden("s2g_forward_golden")
    cfg = O.make_cfg("voice2pose_s2g")
    sd = O.init_voice2pose(cfg, 4, seed=0)
    for k in g.files:
        if k.startswith("buf/"):
            sd[k[4:]] = torch.from_numpy(g[k])
    audio = torch.from_numpy(g["audio"])
    mel = O.mel_spectrogram(audio)
    pred = O.generator_forward(mel, 64, None, sd, cfg, training=False)
    assert pred.shape == (1, 64, 2, 121)
    assert rel_err(pred.numpy(), g["pred"]) < 1e-4


def test_pose2pose_step_matches_reference():
    g = golden("pose2pose_step_golden")
    cfg = O.make_cfg("pose2pose")
    n_train, bs = int(g["n_train"]), int(g["batch_size"])
    orc = O.Pose2PoseOracle(cfg, n_train, seed=0)
    for k, v in orc.sd.items():
        assert np.array_equal(samples_of(v), g["init/%s/samples" % k]), k
    stat = oliver_stat(True)
    for s in range(int(g["steps"])):
        batch = O.synthetic_batch(bs, n_train, stat, seed=200 + s)
        eps = torch.from_numpy(g["step%d/eps" % s])
        losses, results, grads = orc.train_step(batch, eps)
        p = "step%d" % s
        for k, v in losses.items():
            assert abs(float(v) - float(g["%s/loss/%s" % (p, k)])) <= 2e-5 * max(1.0, abs(float(v))), k
        assert rel_err(results["poses_pred_batch"].numpy(), g[p + "/pred"]) < 1e-4
        assert rel_err(results["clip_code_mu"].numpy(), g[p + "/mu"]) < 1e-4
        _check_digests(g, p + "/grad", grads, rtol=1e-3 if s == 0 else 3e-2, what="grad", outlier_frac=0.0 if s == 0 else 0.02)
        orc.apply_optimizers(grads)
        _check_digests(g, p + "/state", orc.sd, rtol=1e-4, atol=2.5e-4 * (s + 1), what="state")


@pytest.mark.parametrize("tag,name,over", [("s2g", "voice2pose_s2g", {}), ("gtcode", "voice2pose_sdt_bp", {"test_with_gt_code": True})])
def test_eval_forward_matches_reference(tag, name, over):
    """Validation forward (model.eval(), return_loss=True) of the reference: BatchNorm from running statistics, the FGD input
    through transform_normalized_parted2global (s2g), the ground-truth code + its KL (TEST_WITH_GT_CODE), LSGAN terms."""
    g = golden("eval_golden")
    cfg = O.make_cfg(name, **over)
    n_train, bs = int(g[tag + "/n_train"]), int(g[tag + "/batch_size"])
    orc = O.Voice2PoseOracle(cfg, n_train, seed=0)
    for k in g.files:
        if k.startswith(tag + "/buf/"):
            orc.sd[k[len(tag) + 5:]] = torch.from_numpy(g[k])
    batch = O.synthetic_batch(bs, n_train, oliver_stat(cfg["hierarchical"]), seed=410, stat_parted=oliver_stat(True),
                              stat_global=oliver_stat(False))
    with torch.no_grad():
        losses, results = orc.forward(batch, training=False)
    ref_keys = {k.split("/")[-1] for k in g.files if k.startswith(tag + "/loss/")}
    assert set(losses) == ref_keys
    for k, v in losses.items():
        ref = float(g["%s/loss/%s" % (tag, k)])
        assert abs(float(v) - ref) <= 1e-4 * max(1.0, abs(ref)), (k, float(v), ref)
    assert rel_err(results["poses_pred_batch"].numpy(), g[tag + "/pred"]) < 1e-4
    for k in ("mu_pred", "mu_gt", "logvar_pred", "logvar_gt"):
        assert rel_err(results[k].numpy(), g["%s/%s" % (tag, k)]) < 1e-3, k
    if tag == "gtcode":
        assert rel_err(results["condition_code"].numpy(), g[tag + "/condition_code"]) < 1e-4
