"""Raw clip -> device batch (speechdrivestemplates_b200/data.py) against the CPU oracle's restatement of
GestureDataset.__getitem__ (gesture_dataset.py:86-105, audio_processing.py:5-19): bit-exact keypoints, exact audio crop/pad,
and a train step fed from the builder equals a train step fed from the host batch."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from util import oliver_stat  # noqa: E402


class _DS:
    AUDIO_LENGTH, AUDIO_SR, FPS, NUM_FRAMES, HIERARCHICAL_POSE = 68267, 16000, 15, 64, True


def _clips(B, seed, lengths):
    g = np.random.default_rng(seed)
    out = []
    for i in range(B):
        pose = np.empty((70, 3, 137), np.float32)
        pose[:, 0] = g.uniform(0, 1280, (70, 137))
        pose[:, 1] = g.uniform(0, 720, (70, 137))
        pose[:, 2] = g.uniform(0, 1, (70, 137))
        out.append(((0.1 * g.standard_normal(lengths[i % len(lengths)])).astype(np.float32), pose, 7 * i + 3))
    return out


def test_builder_matches_the_reference_item_pipeline_bit_for_bit():
    from oracle import sdt_oracle as O
    from speechdrivestemplates_b200 import data
    st = oliver_stat()
    B = 5
    bld = data.DeviceBatchBuilder(_DS, st, B, "cuda:0")
    assert (bld.audio_len, bld.num_frames) == O.parse_audio_length(_DS.AUDIO_LENGTH, _DS.AUDIO_SR, _DS.FPS) == (68266, 64)
    for rep in range(3):                                  # three batches: both staging slots get reused
        clips = _clips(B, 50 + rep, [68266, 70000, 12345, 68265, 90000])
        batch = bld(clips)
        torch.cuda.synchronize()
        for i, (audio, pose, idx) in enumerate(clips):
            ref_a = O.crop_pad_audio(audio, bld.audio_len)
            assert np.array_equal(batch["audio"][i].cpu().numpy(), np.asarray(ref_a, np.float32))
            ref_p = O.preprocess_pose(pose[:64], st["mean"], st["std"], True)
            assert np.array_equal(batch["poses"][i].cpu().numpy(), np.asarray(ref_p)), "keypoint pipeline is not bit-exact"
            assert int(batch["clip_index"][i]) == idx
        assert batch["speaker_stat"]["mean"].dtype == torch.float64
        assert np.array_equal(batch["speaker_stat"]["mean"][0].cpu().numpy(), np.asarray(st["mean"], np.float64).reshape(242))


def test_train_step_from_the_builder_equals_train_step_from_a_host_batch():
    from oracle import sdt_oracle as O
    from speechdrivestemplates_b200 import config, data, pipeline
    st = oliver_stat()
    B, n_train = 4, 64
    clips = [(a, p, i % n_train) for a, p, i in _clips(B, 77, [68266])]
    code0 = 0.1 * torch.randn(n_train, 32, generator=torch.Generator().manual_seed(11))
    outs = []
    for mode in ("builder", "host"):
        tr = pipeline.Voice2PoseTrainer(config.get_cfg("voice2pose_sdt_bp"), n_train, torch.device("cuda:0"), use_cuda_graph=False, seed=0)
        tr.model.clips_code.data.copy_(code0)
        if mode == "builder":
            batch = data.DeviceBatchBuilder(_DS, st, B, "cuda:0")(clips)
        else:
            poses = torch.stack([torch.as_tensor(np.asarray(O.preprocess_pose(p[:64], st["mean"], st["std"], True))) for _, p, _ in clips])
            batch = {"audio": torch.stack([torch.from_numpy(a[:68266]) for a, _, _ in clips]), "poses": poses,
                     "clip_index": torch.tensor([i for _, _, i in clips]), "num_frames": torch.full((B,), 64),
                     "speaker_stat": {"mean": torch.from_numpy(np.tile(np.asarray(st["mean"], np.float64).reshape(1, 242), (B, 1))),
                                      "std": torch.from_numpy(np.tile(np.asarray(st["std"], np.float64).reshape(1, 242), (B, 1))),
                                      "scale_factor": torch.full((B,), float(st["scale_factor"]), dtype=torch.float64)}}
        tr.train_step(batch)
        outs.append(tr.losses_to_host())
    assert outs[0] == outs[1]
