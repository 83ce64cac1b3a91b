"""Checkpoints in the reference's .pth layout (speechdrivestemplates_b200/checkpoint.py): what the fused trainer writes is
accepted by the reference-side objects (drop-in model wrapped in DataParallel-style 'module.' keys, torch.optim.Adam), a resumed
trainer continues bit for bit, and the Adam state means the same thing on both sides."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from util import oliver_stat  # noqa: E402


def _batch(bs, n_train, seed):
    from oracle import sdt_oracle as O
    b = O.synthetic_batch(bs, n_train, oliver_stat(True), seed=seed)
    b["speaker_stat"] = {k: torch.from_numpy(np.asarray(v)) for k, v in b["speaker_stat"].items()}
    return b


def _trainer(n_train=16):
    from speechdrivestemplates_b200 import config, pipeline
    tr = pipeline.Voice2PoseTrainer(config.get_cfg("voice2pose_sdt_bp"), n_train, torch.device("cuda:0"), use_cuda_graph=False, seed=0)
    tr.model.clips_code.data.copy_(0.1 * torch.randn(n_train, 32, generator=torch.Generator().manual_seed(11)))
    return tr


def test_resume_continues_bit_for_bit(tmp_path):
    from speechdrivestemplates_b200 import checkpoint as C
    a = _trainer()
    for s in range(3):
        a.train_step(_batch(4, 16, 700 + s))
    path = tmp_path / "checkpoint_epoch-1_step-3.pth"
    C.save_voice2pose(a, str(path), epoch=1, step=3)
    ck = torch.load(str(path), map_location="cpu")
    assert set(ck) == {"epoch", "step", "model_state_dict", "optimizerG_state_dict", "optimizerClipCode_state_dict"}
    assert all(k.startswith("module.") for k in ck["model_state_dict"])
    assert "module.netG.audio_encoder.specgram_encoder_2d.0.0.conv.weight" in ck["model_state_dict"]
    assert "module.mel_transfm.spectrogram.window" in ck["model_state_dict"] and "module.clips_code" in ck["model_state_dict"]
    b = _trainer()
    assert C.load_voice2pose(b, str(path)) == (1, 3)
    for s in range(3, 5):
        a.train_step(_batch(4, 16, 700 + s))
        b.train_step(_batch(4, 16, 700 + s))
    assert a.losses_to_host() == b.losses_to_host()
    assert torch.equal(a.flat_p, b.flat_p) and torch.equal(a.exp_avg, b.exp_avg) and torch.equal(a.exp_avg_sq, b.exp_avg_sq)


def test_written_optimizer_state_is_what_torch_adam_expects():
    """Load the written optimizerG_state_dict into a real torch.optim.Adam over the same parameters and take one step with the
    same gradients: it must land on the fused trainer's next parameters (1e-7, the documented Adam tolerance)."""
    from speechdrivestemplates_b200 import checkpoint as C
    tr = _trainer()
    for s in range(2):
        tr.train_step(_batch(4, 16, 800 + s))
    ck = C.voice2pose_checkpoint(tr, epoch=0, step=2)
    names = [n for n, _ in tr.model.netG.named_parameters()]
    clones = [torch.nn.Parameter(p.detach().clone()) for _, p in tr.model.netG.named_parameters()]
    opt = torch.optim.Adam(clones, lr=tr.lr, weight_decay=float(tr.cfg.TRAIN.WD))
    opt.load_state_dict(ck["optimizerG_state_dict"])
    assert int(float(opt.state[clones[0]]["step"])) == 2
    tr.train_step(_batch(4, 16, 802))                      # fused step 3; its gradients are still in tr.grads
    for n, c in zip(names, clones):
        c.grad = tr.grads[n].detach().clone()
    opt.step()
    for (n, p), c in zip(tr.model.netG.named_parameters(), clones):
        assert float((p.detach() - c.detach()).abs().max()) <= 1e-7, n


def test_state_dict_keys_strip_and_strictness():
    from speechdrivestemplates_b200 import checkpoint as C
    tr = _trainer()
    ck = C.voice2pose_checkpoint(tr, 0, 0)
    assert ck["optimizerG_state_dict"]["state"] == {}      # torch creates Adam state lazily: nothing before the first step
    bad = dict(ck, model_state_dict={k: v for k, v in ck["model_state_dict"].items() if "clips_code" not in k})
    with pytest.raises(RuntimeError, match="missing"):
        C.load_voice2pose(tr, bad)
    C.load_voice2pose(tr, bad, strict=False)


def test_pose2pose_checkpoint_resume_and_handoff_to_sdt_vae(tmp_path):
    """Pose2PoseTrainer checkpoints in the reference layout (trainer.py:305-321 with the single 'optimizer', pose2pose.py:114):
    (a) a resumed trainer continues bit for bit; (b) the file is what voice2pose_sdt_vae reads as VOICE2POSE.POSE_ENCODER.AE_CHECKPOINT
    -- external clip codes from module.clip_code_mu (voice2pose.py:40-55) and the FGD encoder from module.ae.encoder.* (:234-242)."""
    from speechdrivestemplates_b200 import checkpoint as C, config, pipeline
    n_train = 16
    dev = torch.device("cuda:0")

    def p2p():
        tr = pipeline.Pose2PoseTrainer(config.get_cfg("pose2pose"), n_train, dev, use_cuda_graph=False, seed=3)
        return tr

    a = p2p()
    eps = [torch.randn(4, 32, generator=torch.Generator().manual_seed(50 + s)).to(dev) for s in range(5)]
    for s in range(3):
        a.eps_override = eps[s]
        a.train_step(_batch(4, n_train, 900 + s))
    path = str(tmp_path / "checkpoint_epoch-2_step-3.pth")
    C.save_pose2pose(a, path, epoch=2, step=3)
    ck = torch.load(path, map_location="cpu")
    assert set(ck) == {"epoch", "step", "model_state_dict", "optimizer_state_dict"}
    keys = set(ck["model_state_dict"])
    assert {"module.clip_code_mu", "module.clip_code_logvar", "module.mel_transfm.spectrogram.window",
            "module.ae.encoder.blocks.0.conv.weight", "module.ae.decoder.blocks.4.bias"} <= keys
    assert ck["optimizer_state_dict"]["state"][0]["exp_avg"].shape == ck["model_state_dict"]["module.ae.encoder.blocks.0.conv.weight"].shape
    b = p2p()
    assert C.load_pose2pose(b, path) == (2, 3)
    for s in range(3, 5):
        a.eps_override = b.eps_override = eps[s]
        oa = a.train_step(_batch(4, n_train, 900 + s))
        ob = b.train_step(_batch(4, n_train, 900 + s))
    assert a.losses_to_host(oa) == b.losses_to_host(ob)
    assert torch.equal(a.flat_p, b.flat_p) and torch.equal(a.exp_avg_sq, b.exp_avg_sq)
    assert torch.equal(a.model.clip_code_mu, b.model.clip_code_mu)
    # (b) hand-off
    cfg = config.get_cfg("voice2pose_sdt_vae", ["VOICE2POSE.POSE_ENCODER.AE_CHECKPOINT", path])
    v = pipeline.Voice2PoseTrainer(cfg, n_train, dev, use_cuda_graph=False, seed=0)
    assert not isinstance(v.model.clips_code, torch.nn.Parameter)
    assert torch.equal(v.model.clips_code.cpu(), ck["model_state_dict"]["module.clip_code_mu"])
    enc = ck["model_state_dict"]["module.ae.encoder.blocks.3.norm.running_var"]
    assert torch.equal(v.model.pose_encoder.state_dict()["blocks.3.norm.running_var"].cpu(), enc)
    out = v.train_step(_batch(4, n_train, 950))
    assert np.isfinite(v.losses_to_host(out)["G_loss"])
