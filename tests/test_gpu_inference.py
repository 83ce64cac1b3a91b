"""Long-audio (demo) inference, SURVEY §8f row 4: the time-tiled forward of speechdrivestemplates_b200/inference.py equals the
one-shot forward the reference runs (trainer.py:459-484, voice2pose.py:386-410), and the one-shot forward equals the oracle."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_chan_stats_window_matches_torch():
    from speechdrivestemplates_b200 import ops
    g = torch.Generator().manual_seed(3)
    for (B, H, W, C, w0, w1, rpp) in [(2, 5, 37, 64, 3, 29, 64), (1, 80, 50, 64, 0, 50, 64), (3, 10, 21, 256, 20, 21, 16), (2, 4, 9, 128, 5, 5, 8)]:
        x = torch.randn(B, H, W, C, generator=g).cuda()
        n_parts = -(-(H * (w1 - w0)) // rpp) + 2
        partial = torch.full((B, n_parts, 2, C), 7.0, device="cuda")
        ops.chan_stats(x, w0, w1, rpp, partial)
        win = x[:, :, w0:w1].double()
        s, q = win.sum((1, 2)), (win * win).sum((1, 2))
        got = partial.double().sum(1)
        assert torch.allclose(got[:, 0], s, rtol=1e-5, atol=1e-3), (B, H, W, C)
        assert torch.allclose(got[:, 1], q, rtol=1e-5, atol=1e-3)
        used = -(-(H * (w1 - w0)) // rpp)
        assert float(partial[:, used:].abs().max()) == 0.0          # parts past the window are written as zeros


@pytest.mark.parametrize("mode,store,tol", [(0, (3, 5), 2e-4), (0, (), 2e-4), (3, (3, 5), 3e-3), (3, (1, 4), 3e-3)])
def test_chunked_inference_equals_one_shot_on_60s(mode, store, tol):
    """60 s of audio (900 frames), 64-frame chunks -> 14 tiles with 64-column halos; InstanceNorm2d statistics are exact, so the
    stream equals the one-shot forward up to the summation order of the statistics (stated bound 2e-4 of the pose range in fp32 math).
    In the TF32 math mode the two differ by TF32 rounding noise instead: the one-shot path runs its first block in the fused
    single-pass kernel, the tiles run it as a convolution, and a 1e-7 difference in an activation can flip its rounding to 10
    mantissa bits (5e-4 relative) -- the same ~1e-3 floor mode 3 has against the fp32 oracle (stated bound 3e-3)."""
    from speechdrivestemplates_b200 import config, data, inference
    cfg = config.get_cfg("voice2pose_sdt_bp")
    alen, nf = data.parse_audio_length(60 * 16000, 16000, 15)
    audio = 0.1 * torch.randn(1, alen, generator=torch.Generator().manual_seed(8))
    code = 0.1 * torch.randn(1, 32, generator=torch.Generator().manual_seed(9))
    torch.manual_seed(0)
    one = inference.StreamingGenerator(cfg, "cuda:0", conv_math=mode, chunk_frames=0)
    torch.cuda.reset_peak_memory_stats()
    base = torch.cuda.memory_allocated()
    ref = one(audio, nf, code)
    assert one.last_chunks == 1
    one_shot_bytes = torch.cuda.max_memory_allocated() - base
    one.netG._eng.arena.bufs.clear()                     # drop the one-shot workspaces before measuring the tiled path
    torch.cuda.empty_cache()
    torch.cuda.reset_peak_memory_stats()
    base = torch.cuda.memory_allocated()
    chunked = inference.StreamingGenerator(cfg, "cuda:0", conv_math=mode, chunk_frames=64, netG=one.netG, mel=one.mel, store_layers=store)
    out = chunked(audio, nf, code)
    assert chunked.last_chunks >= 12
    assert out.shape == ref.shape == (1, nf, 2, cfg.DATASET.NUM_LANDMARKS)
    err = float((out - ref).abs().max() / ref.abs().max())
    assert err < tol, err
    # a second call reuses every buffer and gives the same stream
    again = chunked(audio, nf, code)
    assert torch.equal(again, out)
    assert torch.cuda.max_memory_allocated() - base < 0.5 * one_shot_bytes, (torch.cuda.max_memory_allocated() - base, one_shot_bytes)


def test_one_shot_inference_matches_the_oracle_on_8s():
    from oracle import sdt_oracle as O
    from speechdrivestemplates_b200 import config, data, inference
    cfg = config.get_cfg("voice2pose_sdt_bp")
    alen, nf = data.parse_audio_length(8 * 16000, 16000, 15)
    audio = 0.1 * torch.randn(1, alen, generator=torch.Generator().manual_seed(18))
    code = 0.1 * torch.randn(1, 32, generator=torch.Generator().manual_seed(19))
    torch.manual_seed(0)
    gen = inference.StreamingGenerator(cfg, "cuda:0", conv_math=0, chunk_frames=0)
    sd = {"netG." + k: v.detach().cpu() for k, v in gen.netG.state_dict().items()}
    ocfg = O.make_cfg("voice2pose_sdt_bp")
    with torch.no_grad():
        mel = O.mel_spectrogram(audio)
        ref = O.generator_forward(mel, nf, code, sd, ocfg, False, "netG.")
    for chunk in (0, 40):
        gen.chunk_frames = chunk
        out = gen(audio, nf, code)
        err = float((out - ref.view_as(out)).abs().max() / ref.abs().max())
        assert err < 2e-4, (chunk, err)
    assert gen.last_chunks >= 2


def test_dropin_model_demo_branch_streams_long_audio(monkeypatch):
    """Voice2PoseModel.forward(return_loss=False) -- what Voice2Pose.demo_step calls (voice2pose.py:386-410) -- switches to the
    time-tiled forward for long utterances when SDT_DEMO_CHUNK_FRAMES is set, with the same result as the one-shot forward."""
    from speechdrivestemplates_b200 import config, data, pipeline
    cfg = config.get_cfg("voice2pose_sdt_bp", ["DEMO.CODE_INDEX", 3])
    torch.manual_seed(0)
    model = pipeline.Voice2PoseModel(cfg, num_train_samples=8).cuda().eval()
    model.clips_code.data.copy_(0.1 * torch.randn(8, 32, generator=torch.Generator().manual_seed(2)))
    alen, nf = data.parse_audio_length(30 * 16000, 16000, 15)
    batch = {"audio": 0.1 * torch.randn(1, alen, generator=torch.Generator().manual_seed(4)), "clip_index": torch.zeros(1, dtype=torch.long),
             "num_frames": torch.tensor([nf])}
    with torch.no_grad():
        ref = model(batch, None, return_loss=False)["poses_pred_batch"].clone()
    monkeypatch.setenv("SDT_DEMO_CHUNK_FRAMES", "64")
    with torch.no_grad():
        out = model(batch, None, return_loss=False)["poses_pred_batch"]
    assert model._streaming.last_chunks >= 5 and out.shape == ref.shape == (1, nf, 2, 121)
    assert float((out - ref).abs().max() / ref.abs().max()) < 2e-4        # conftest pins the fp32 math mode


def test_graphed_tiled_forward_equals_the_eager_one():
    """graph=True: the tiled forward of a given utterance length is captured once and replayed; same stream, bit for bit, also for a
    different utterance of the same length (the captured graph reads the static input buffers)."""
    from speechdrivestemplates_b200 import config, data, inference
    cfg = config.get_cfg("voice2pose_sdt_bp")
    alen, nf = data.parse_audio_length(40 * 16000, 16000, 15)
    code = 0.1 * torch.randn(1, 32, generator=torch.Generator().manual_seed(9))
    torch.manual_seed(0)
    eager = inference.StreamingGenerator(cfg, "cuda:0", conv_math=3, chunk_frames=96)
    graphed = inference.StreamingGenerator(cfg, "cuda:0", conv_math=3, chunk_frames=96, netG=eager.netG, mel=eager.mel, graph=True)
    for seed in (1, 2, 3):
        audio = 0.1 * torch.randn(1, alen, generator=torch.Generator().manual_seed(seed))
        ref = eager(audio, nf, code)
        out = graphed(audio, nf, code)
        assert torch.equal(out, ref), seed
    assert len(graphed._graph_cache) == 1 and graphed.last_chunks >= 4
