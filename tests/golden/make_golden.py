"""Generate the golden fixtures in tests/golden/ by running the UNMODIFIED reference on CPU.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

The reference ships no tests or golden vectors, so these fixtures ARE the parity pin:
they are outputs of the reference's own classes (``Voice2PoseModel``, ``Pose2PoseModel``,
``GestureDataset`` methods, ``torchaudio.transforms.MelSpectrogram`` as configured at
core/pipelines/voice2pose.py:27-30) driven with the reference's optimizer choreography
(voice2pose.py:298-309, pose2pose.py:145-147) on seeded synthetic inputs.
``tests/test_oracle_golden.py`` checks ``oracle/sdt_oracle.py`` against them; the GPU parity tests
check the CUDA path against the oracle and, for the small cases, directly against these files.
"""
import os
import sys
from collections import OrderedDict

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import refshim  # noqa: E402

refshim.install()
from oracle import sdt_oracle as O  # noqa: E402  (only for synthetic_batch: shared seeded inputs)

N_SAMPLES = 32  # sampled elements per tensor


def sample_idx(numel, n=N_SAMPLES):
    """Deterministic pseudo-random flat indices (LCG), identical in the tests."""
    out, x = [], 12345 + numel
    for _ in range(min(n, numel)):
        x = (1103515245 * x + 12345) % (2 ** 31)
        out.append(x % numel)
    return np.asarray(out, np.int64)


def tensor_digest(t):
    t = t.detach().double().flatten()
    return np.asarray([t.sum().item(), (t * t).sum().item(), t.abs().sum().item()], np.float64)


def digest_dict(prefix, d, out):
    for k, v in d.items():
        if v is None:
            continue
        v = v.detach()
        out["%s/%s/digest" % (prefix, k)] = tensor_digest(v)
        out["%s/%s/samples" % (prefix, k)] = v.flatten()[torch.from_numpy(sample_idx(v.numel()))].numpy()


def speaker_stats():
    from core.datasets.speakers_stat import SPEAKERS_STAT_121, SPEAKERS_STAT_121_parted
    out = {}
    for tag, table in (("global", SPEAKERS_STAT_121), ("parted", SPEAKERS_STAT_121_parted)):
        s = table["oliver"]
        out[tag + "_mean"] = np.asarray(s["mean"], np.float64)
        out[tag + "_std"] = np.asarray(s["std"], np.float64)
        out[tag + "_scale_factor"] = np.float64(s["scale_factor"])
    np.savez(os.path.join(HERE, "speaker_stat_oliver.npz"), **out)
    return out


def stat_of(stats, parted):
    tag = "parted" if parted else "global"
    return {"mean": stats[tag + "_mean"], "std": stats[tag + "_std"], "scale_factor": stats[tag + "_scale_factor"]}


def to_ref_batch(batch):
    """oracle-style batch (numpy stats) -> what the reference's default collate would hand over."""
    b = dict(batch)
    b["speaker_stat"] = {k: torch.from_numpy(np.asarray(v)) for k, v in batch["speaker_stat"].items()}
    return b


def golden_mel():
    import torchaudio
    from scipy.io import wavfile
    tf = torchaudio.transforms.MelSpectrogram(win_length=400, hop_length=160, n_fft=512, f_min=55, f_max=7500.0, n_mels=80)
    sr, wav = wavfile.read(os.path.join(refshim.REFERENCE_ROOT, "demo_audio.wav"))
    assert sr == 16000
    wav = np.asarray(wav, np.float32)
    real = torch.from_numpy(wav[16000:16000 + 68266].copy())[None]
    g = torch.Generator().manual_seed(7)
    syn = 0.1 * torch.randn(2, 68266, generator=g)
    short = 0.1 * torch.randn(1, 4000, generator=g)          # ragged / short input: 26 frames
    odd = 0.1 * torch.randn(1, 1601, generator=g)            # T = 11
    np.savez_compressed(
        os.path.join(HERE, "mel_golden.npz"),
        window=tf.spectrogram.window.numpy(), fb=tf.mel_scale.fb.numpy(),
        real_audio=real.numpy(), real_mel=tf(real).numpy(),
        syn_seed=np.int64(7), syn_mel=tf(syn).numpy().astype(np.float32),
        short_mel=tf(short).numpy(), odd_mel=tf(odd).numpy(),
    )


def golden_keypoints(stats):
    cfg = refshim.get_cfg("voice2pose_sdt_bp")
    ds = refshim.make_dataset_stub(cfg)
    rng = np.random.RandomState(3)
    raw = np.stack([rng.uniform(0, 1280, (8, 137)), rng.uniform(0, 720, (8, 137)), rng.uniform(0, 1, (8, 137))], 1).astype(np.float32)
    out = {"raw": raw}
    for parted in (True, False):
        tag = "parted" if parted else "global"
        st = stat_of(stats, parted)
        p = torch.Tensor(raw.copy())
        p = ds.remove_unuesd_kp(p)
        p = ds.absolute_to_relative(p)
        if parted:
            p = ds.global_to_parted(p)
        rel = p[:, :2, :]
        norm = ds.normalize_poses(rel, {"mean": st["mean"], "std": st["std"]})
        out[tag + "_relative"] = rel.numpy().copy()
        out[tag + "_normalized"] = norm.numpy().copy()
        # get_final_results on a (B,T,2,K) f32 batch with collated f64 stats
        b = 3
        x = torch.from_numpy(rng.standard_normal((b, 8, 2, 121)).astype(np.float32))
        cstat = {"mean": torch.from_numpy(np.tile(st["mean"][None], (b, 1))),
                 "std": torch.from_numpy(np.tile(st["std"][None], (b, 1))),
                 "scale_factor": torch.from_numpy(np.full((b,), st["scale_factor"]))}
        ds.cfg = refshim.get_cfg("voice2pose_sdt_bp", ["DATASET.HIERARCHICAL_POSE", parted]).DATASET
        fin = ds.get_final_results(x.clone(), cstat)
        assert fin.dtype == torch.float64
        out[tag + "_final_in"] = x.numpy()
        out[tag + "_final_out"] = fin.numpy()
    np.savez_compressed(os.path.join(HERE, "keypoints_golden.npz"), **out)


def run_voice2pose(config_name, stats, batch_size, n_train, steps, live_code, tag):
    from core.pipelines.voice2pose import Voice2PoseModel
    cfg = refshim.get_cfg(config_name)
    ocfg = O.make_cfg(config_name)
    torch.manual_seed(0)
    model = Voice2PoseModel(cfg, num_train_samples=n_train)
    if live_code:
        g = torch.Generator().manual_seed(11)
        with torch.no_grad():
            model.clips_code.copy_(0.1 * torch.randn(n_train, 32, generator=g))
    model.train()                                              # trainer.py:382
    ds = refshim.make_dataset_stub(cfg)
    out = {"batch_size": np.int64(batch_size), "n_train": np.int64(n_train), "steps": np.int64(steps),
           "live_code": np.int64(live_code)}
    digest_dict("init", model.state_dict(), out)
    optG = torch.optim.Adam(model.netG.parameters(), lr=cfg.TRAIN.LR, weight_decay=cfg.TRAIN.WD)
    optC = torch.optim.Adam([model.clips_code], lr=cfg.TRAIN.LR * cfg.VOICE2POSE.GENERATOR.CLIP_CODE.LR_SCALING) \
        if isinstance(model.clips_code, torch.nn.Parameter) else None
    optD = torch.optim.Adam(model.netD_pose.parameters(), lr=cfg.TRAIN.LR) if hasattr(model, "netD_pose") else None
    for s in range(steps):
        batch = O.synthetic_batch(batch_size, n_train, stat_of(stats, ocfg["hierarchical"]), seed=100 + s)
        rb = to_ref_batch(batch)
        losses, results = model(rb, ds)                                                     # voice2pose.py:288
        fin_pred = ds.get_final_results(results["poses_pred_batch"].detach(), rb["speaker_stat"])
        fin_gt = ds.get_final_results(results["poses_gt_batch"].detach(), rb["speaker_stat"])
        d = fin_pred - fin_gt                                                               # evaluate_step :412-430
        l2 = torch.norm(d, p=2, dim=2)
        lp = torch.norm(fin_pred[:, :, :, 75] - fin_pred[:, :, :, 71], p=2, dim=-1)
        lg = torch.norm(fin_gt[:, :, :, 75] - fin_gt[:, :, :, 71], p=2, dim=-1)
        den = lg.max(-1, keepdim=True).values + 1e-4
        losses["L2_dist"] = l2.mean()
        losses["lip_sync_error_n"] = torch.abs(lp / den - lg / den).mean()
        if optC is not None:
            optC.zero_grad()
        optG.zero_grad()
        losses["G_loss"].backward(retain_graph=True)                                        # :301
        grads = OrderedDict((k, p.grad.clone() if p.grad is not None else None) for k, p in model.named_parameters()
                            if k.startswith("netG.") or k == "clips_code")
        if optC is not None:
            optC.step()
        optG.step()
        if optD is not None:
            optD.zero_grad()
            losses["D_pose_gan_loss"].backward()
            grads.update((k, p.grad.clone()) for k, p in model.named_parameters() if k.startswith("netD_pose."))
            optD.step()
        p = "step%d" % s
        for k, v in losses.items():
            out["%s/loss/%s" % (p, k)] = np.float64(v.item())
        out[p + "/pred"] = results["poses_pred_batch"].detach().numpy().astype(np.float32)
        out[p + "/final_pred"] = fin_pred.numpy()
        for k in ("mu_pred", "mu_gt", "logvar_pred", "logvar_gt"):
            if k in results:
                out["%s/%s" % (p, k)] = results[k].numpy()
        digest_dict(p + "/grad", grads, out)
        digest_dict(p + "/state", model.state_dict(), out)
    np.savez_compressed(os.path.join(HERE, tag + ".npz"), **out)
    print(tag, {k: float(v) for k, v in out.items() if "/loss/" in k})


def golden_s2g_forward(stats):
    """BASELINE.json configs[0]: voice2pose_s2g forward, 1 clip x 64 frames, eval mode (BN running stats)."""
    from core.pipelines.voice2pose import Voice2PoseModel
    cfg = refshim.get_cfg("voice2pose_s2g")
    torch.manual_seed(0)
    model = Voice2PoseModel(cfg, num_train_samples=4)
    # non-trivial running stats so that BN-eval is exercised: two training-mode forwards first
    model.train()
    g = torch.Generator().manual_seed(5)
    with torch.no_grad():
        for _ in range(2):
            model.netG(model.mel_transfm(0.1 * torch.randn(2, 68266, generator=g)), 64, None)
    model.eval()
    audio = 0.1 * torch.randn(1, 68266, generator=g)
    with torch.no_grad():
        mel = model.mel_transfm(audio)
        pred = model.netG(mel, 64, None)
    out = {"audio": audio.numpy(), "pred": pred.numpy()}
    digest_dict("state", model.state_dict(), out)
    # full BN buffers of netG (small) so the oracle / CUDA path can load the exact eval-time state
    for k, v in model.state_dict().items():
        if k.startswith("netG.") and ("running_" in k or "num_batches" in k):
            out["buf/" + k] = v.numpy()
    np.savez_compressed(os.path.join(HERE, "s2g_forward_golden.npz"), **out)


def golden_eval(stats):
    """Validation forward (``model.eval()``, return_loss=True: Voice2Pose.test_step, voice2pose.py:333-352) of the UNMODIFIED
    reference for the two configurations whose eval path differs from training beyond the BatchNorm mode:
      s2g     -- BatchNorm generator / discriminator / FGD extractor from running statistics, HIERARCHICAL_POSE=False so the FGD
                 input goes through dataset.transform_normalized_parted2global (:164-169), LSGAN terms in the losses;
      gtcode  -- voice2pose_sdt_bp with CLIP_CODE.TEST_WITH_GT_CODE: condition_code = mu of the ground-truth poses (:100-106),
                 clip-code KL on that code (:147-157).
    Parameters come from torch.manual_seed(0) + the reference constructors (reproduced by the drop-in modules / the oracle);
    the BatchNorm buffers after a few training-mode forwards are stored in full (they are small)."""
    from core.pipelines.voice2pose import Voice2PoseModel
    out = {}
    for tag, name, opts, bs in (("s2g", "voice2pose_s2g", [], 3),
                                ("gtcode", "voice2pose_sdt_bp", ["VOICE2POSE.GENERATOR.CLIP_CODE.TEST_WITH_GT_CODE", True], 3)):
        cfg = refshim.get_cfg(name, opts)
        ocfg = O.make_cfg(name)
        n_train = 8
        torch.manual_seed(0)
        model = Voice2PoseModel(cfg, num_train_samples=n_train)
        ds = refshim.make_dataset_stub(cfg)
        st = stat_of(stats, ocfg["hierarchical"])
        model.train()
        with torch.no_grad():                      # move every BatchNorm's running statistics off their initial values
            for s in range(2):
                model(to_ref_batch(O.synthetic_batch(bs, n_train, st, seed=400 + s)), ds)
        model.eval()
        batch = O.synthetic_batch(bs, n_train, st, seed=410)
        with torch.no_grad():
            losses, results = model(to_ref_batch(batch), ds)
        for k, v in losses.items():
            out["%s/loss/%s" % (tag, k)] = np.float64(v.item())
        out[tag + "/pred"] = results["poses_pred_batch"].numpy().astype(np.float32)
        for k in ("mu_pred", "mu_gt", "logvar_pred", "logvar_gt", "condition_code"):
            if results.get(k) is not None:
                out["%s/%s" % (tag, k)] = results[k].numpy()
        out[tag + "/batch_size"], out[tag + "/n_train"] = np.int64(bs), np.int64(n_train)
        for k, v in model.state_dict().items():
            if "running_" in k or "num_batches" in k:
                out["%s/buf/%s" % (tag, k)] = v.numpy()
        print("eval", tag, {k: float(v) for k, v in losses.items()})
    np.savez_compressed(os.path.join(HERE, "eval_golden.npz"), **out)


def golden_pose2pose(stats, batch_size=4, n_train=16, steps=2):
    from core.pipelines.pose2pose import Pose2PoseModel
    import core.networks.poses_reconstruction.autoencoder as ae_mod
    cfg = refshim.get_cfg("pose2pose")
    torch.manual_seed(0)
    model = Pose2PoseModel(cfg, num_train_samples=n_train)
    model.train()
    out = {"batch_size": np.int64(batch_size), "n_train": np.int64(n_train), "steps": np.int64(steps)}
    digest_dict("init", model.state_dict(), out)
    opt = torch.optim.Adam(model.ae.parameters(), lr=cfg.TRAIN.LR, weight_decay=cfg.TRAIN.WD)
    for s in range(steps):
        batch = O.synthetic_batch(batch_size, n_train, stat_of(stats, True), seed=200 + s)
        eps = torch.randn(batch_size, 32, generator=torch.Generator().manual_seed(300 + s))
        real_randn = torch.randn
        ae_mod.torch.randn = lambda *a, **k: eps.clone()        # inject the N(0,1) draw (autoencoder.py:86)
        try:
            losses, results = model(to_ref_batch(batch))
        finally:
            ae_mod.torch.randn = real_randn
        idx = batch["clip_index"]
        model.clip_code_mu[idx] = results["clip_code_mu"].detach()
        model.clip_code_logvar[idx] = results["clip_code_logvar"].detach()
        opt.zero_grad()
        losses["loss"].backward(retain_graph=True)
        grads = OrderedDict((k, p.grad.clone()) for k, p in model.named_parameters())
        opt.step()
        p = "step%d" % s
        for k, v in losses.items():
            out["%s/loss/%s" % (p, k)] = np.float64(v.item())
        out[p + "/eps"] = eps.numpy()
        out[p + "/pred"] = results["poses_pred_batch"].detach().numpy()
        out[p + "/mu"] = results["clip_code_mu"].detach().numpy()
        out[p + "/logvar"] = results["clip_code_logvar"].detach().numpy()
        digest_dict(p + "/grad", grads, out)
        digest_dict(p + "/state", model.state_dict(), out)
    np.savez_compressed(os.path.join(HERE, "pose2pose_step_golden.npz"), **out)
    print("pose2pose", {k: float(v) for k, v in out.items() if "/loss/" in k})


def main():
    torch.set_num_threads(8)
    stats = speaker_stats()
    golden_mel()
    golden_keypoints(stats)
    golden_s2g_forward(stats)
    run_voice2pose("voice2pose_sdt_bp", stats, 2, 16, 2, True, "sdt_bp_step_golden")
    run_voice2pose("voice2pose_sdt_bp", stats, 2, 16, 1, False, "sdt_bp_zero_code_golden")
    run_voice2pose("voice2pose_s2g", stats, 2, 16, 2, False, "s2g_step_golden")
    golden_pose2pose(stats)
    golden_eval(stats)
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print("%-32s %8.1f KB" % (f, os.path.getsize(os.path.join(HERE, f)) / 1024))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "eval":        # only the (round-2) eval fixture
        torch.set_num_threads(8)
        golden_eval(speaker_stats())
    else:
        main()
