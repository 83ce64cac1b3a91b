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
