"""Raw clip -> device batch: the numeric part of ``GestureDataset.__getitem__`` + default collate
(core/datasets/gesture_dataset.py:86-119, core/utils/audio_processing.py:5-19) with the keypoint work on the GPU.

The reference's DataLoader workers load a clip's ``.npz`` (``audio`` f32 samples, ``pose`` (frames,3,137) f32), crop / pad the
audio to ``AUDIO_LENGTH`` rounded down to whole video frames, select 122 of the 137 keypoints, make them relative to the neck,
split them into the parted hierarchy and normalise with the speaker statistics -- per clip, on CPU.  At ~9 k clips/s per
GPU that is the next bottleneck, so here the host only packs the raw arrays into pinned staging buffers
(audio crop / pad is a memcpy with a zero tail) and the keypoint pipeline runs as ONE launch of ``sdt_pose_preprocess`` over the
whole batch, bit-exact with the reference (IEEE f32 subtract / divide, tests/test_gpu_data.py).  The result is the batch dict
``Voice2PoseTrainer.train_step`` / ``Voice2PoseModel.forward`` take (device tensors are handed over device-to-device).
"""
import numpy as np
import torch

from . import ops


def parse_audio_length(audio_length, sr, fps):
    """audio_processing.py:5-11: whole video frames only."""
    bit_per_frames = sr / fps
    num_frames = int(audio_length / bit_per_frames)
    return int(num_frames * bit_per_frames), num_frames


class DeviceBatchBuilder:
    """Packs raw clips into pinned staging buffers, uploads them on a copy stream and preprocesses the poses on the device.

    cfg: the reference's DATASET node (AUDIO_LENGTH, AUDIO_SR, FPS, NUM_FRAMES, HIERARCHICAL_POSE).
    speaker_stat: {'mean': (242), 'std': (242), 'scale_factor': float} of the speaker (speakers_stat.py), as returned by
    ``GestureDataset.get_speaker_stat``; the f64 values go into the batch untouched (``get_final_results`` needs them in
    f64), the f32 casts used by ``normalize_poses`` (gesture_dataset.py:175-177) are made once.
    """

    def __init__(self, cfg, speaker_stat, batch_size, device):
        self.device = torch.device(device)
        self.B = int(batch_size)
        self.audio_len, self.num_frames = parse_audio_length(cfg.AUDIO_LENGTH, cfg.AUDIO_SR, cfg.FPS)
        self.T = int(cfg.NUM_FRAMES)
        self.hier = bool(cfg.HIERARCHICAL_POSE)
        mean = np.asarray(speaker_stat["mean"], np.float64).reshape(242)
        std = np.asarray(speaker_stat["std"], np.float64).reshape(242)
        self.mean32 = torch.from_numpy(mean.astype(np.float32)).to(self.device)
        self.std32 = torch.from_numpy(std.astype(np.float32)).to(self.device)
        self.stat_dev = {"mean": torch.from_numpy(mean).to(self.device).expand(self.B, 242).contiguous(),
                         "std": torch.from_numpy(std).to(self.device).expand(self.B, 242).contiguous(),
                         "scale_factor": torch.full((self.B,), float(speaker_stat["scale_factor"]), dtype=torch.float64,
                                                    device=self.device)}
        # two pinned staging sets: the host fills one while the other is in flight
        self._host = [dict(audio=torch.zeros(self.B, self.audio_len).pin_memory(),
                           pose=torch.zeros(self.B, self.T, 3, 137).pin_memory(),
                           idx=torch.zeros(self.B, dtype=torch.long).pin_memory()) for _ in range(2)]
        self._dev = [dict(audio=torch.empty(self.B, self.audio_len, device=self.device),
                          pose=torch.empty(self.B, self.T, 3, 137, device=self.device),
                          idx=torch.empty(self.B, dtype=torch.long, device=self.device),
                          poses=torch.empty(self.B, self.T, 2, 121, device=self.device)) for _ in range(2)]
        self._copy = torch.cuda.Stream(device=self.device)
        self._done = [torch.cuda.Event(), torch.cuda.Event()]       # upload + preprocessing of slot i finished
        self._free = [torch.cuda.Event(), torch.cuda.Event()]       # the H2D copies that read pinned set i have finished
        self._consumed = [torch.cuda.Event(), torch.cuda.Event()]   # the consumer's last read of device set i (release())
        for e in self._free + self._consumed:
            e.record(torch.cuda.current_stream(self.device))
        self._handed = [False, False]                               # batch(i) handed out and not released yet
        self._slot = 0

    def pack(self, clips):
        """clips: sequence of B (audio 1-D f32 array, pose (frames,3,137) f32 array, clip_index).  Host work only: crop / pad
        (audio_processing.py:14-19) into the pinned buffers of the next slot.  Returns the slot."""
        assert len(clips) == self.B, "expected %d clips, got %d" % (self.B, len(clips))
        s = self._slot
        self._slot ^= 1
        self._free[s].synchronize()                       # the device copy that read this pinned set has completed
        h = self._host[s]
        for i, (audio, pose, idx) in enumerate(clips):
            a = np.asarray(audio, np.float32).reshape(-1)
            n = min(a.shape[0], self.audio_len)
            h["audio"][i, :n] = torch.from_numpy(a[:n])
            if n < self.audio_len:
                h["audio"][i, n:] = 0.0
            p = np.asarray(pose, np.float32)[:self.T]
            assert p.shape == (self.T, 3, 137), "pose must hold at least NUM_FRAMES frames of (3,137)"
            h["pose"][i] = torch.from_numpy(p)
            h["idx"][i] = int(idx)
        return s

    def upload(self, slot):
        """H2D on the copy stream + the keypoint pipeline (gather 122 of 137, neck-relative, parted, normalise) in one launch."""
        h, d = self._host[slot], self._dev[slot]
        cs = self._copy
        if self._handed[slot]:
            # the previous batch of this slot was never released: fall back to "everything enqueued on the consumer's
            # stream so far" (correct for any consumer on the current stream, but serialises the upload behind the step in flight)
            self.release(slot)
        cs.wait_event(self._consumed[slot])                # the device set is overwritten only after its last reader
        with torch.cuda.stream(cs):
            for k in ("audio", "pose", "idx"):
                d[k].copy_(h[k], non_blocking=True)
            self._free[slot].record(cs)                    # pinned set reusable once the copies are done
            ops.pose_preprocess(d["pose"].view(self.B * self.T, 3, 137), self.mean32, self.std32, self.hier,
                                out=d["poses"].view(self.B * self.T, 2, 121))
            self._done[slot].record(cs)

    def batch(self, slot):
        """The reference's collated batch dict with device tensors.  The consumer calls ``batch["_release"]()`` (or
        ``builder.release(slot)``) on the stream of its LAST read of these tensors -- ``Voice2PoseTrainer`` does so right after
        its device-to-device staging copies -- and the next upload into this slot (two calls later) waits for exactly that
        point.  Without a release the next upload waits for everything enqueued on the current stream at that time."""
        torch.cuda.current_stream(self.device).wait_event(self._done[slot])
        d = self._dev[slot]
        self._handed[slot] = True
        return {"audio": d["audio"], "poses": d["poses"], "clip_index": d["idx"],
                "num_frames": torch.full((self.B,), self.num_frames, dtype=torch.long), "speaker_stat": self.stat_dev,
                "_release": lambda stream=None, _s=slot: self.release(_s, stream)}

    def release(self, slot, stream=None):
        """The consumer is done reading the device tensors of ``slot`` once the work enqueued so far on ``stream`` (default: the
        current stream) has run."""
        self._consumed[slot].record(stream if stream is not None else torch.cuda.current_stream(self.device))
        self._handed[slot] = False

    def __call__(self, clips):
        s = self.pack(clips)
        self.upload(s)
        return self.batch(s)
