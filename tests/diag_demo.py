"""Diagnostic (not a test): BASELINE configs[4] -- demo inference over a long wav (10 min @ 16 kHz -> 9000 frames x 137 kpts),
one fully-convolutional forward (mel + generator, eval mode, fixed clip code) timed with CUDA events.
    python tests/diag_demo.py [seconds] [math mode]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from speechdrivestemplates_b200 import config, data, networks, ops, pipeline  # noqa: E402


def main():
    seconds = int(sys.argv[1]) if len(sys.argv) > 1 else 600
    mode = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    sr, fps = 16000, 15
    alen, nf = data.parse_audio_length(seconds * sr, sr, fps)
    dev = torch.device("cuda:0")
    ops.set_conv_math(mode)
    torch.manual_seed(0)
    net = networks.SequenceGeneratorCNN(config.get_cfg("voice2pose_sdt_bp")).to(dev).eval()
    mel = pipeline.MelSpectrogram().to(dev)
    audio = (0.1 * torch.randn(1, alen, generator=torch.Generator().manual_seed(8))).pin_memory()
    code = 0.1 * torch.randn(1, 32, generator=torch.Generator().manual_seed(9)).to(dev)

    def run():
        with torch.no_grad():
            a = audio.to(dev, non_blocking=True)
            pred = net(mel(a), nf, code)
            return pred.cpu()                       # the pose stream back on the host, as the demo writes it out

    for _ in range(2):
        out = run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    e0.record()
    for _ in range(reps):
        out = run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    flops = 7356.0e6 * (nf / 64.0)                  # generator forward, SURVEY 8d: 7,356 MFLOP per 64-frame clip
    print(json.dumps({"workload": "demo inference, %d s of 16 kHz audio -> %d frames x 121 keypoints, host wav in / host poses out" % (seconds, nf),
                      "math_mode": mode, "ms": ms, "frames_per_s": nf / ms * 1e3, "x_realtime": seconds / (ms * 1e-3),
                      "generator_TFLOPs": flops / ms / 1e9, "peak_mem_GB": torch.cuda.max_memory_allocated() / 1e9,
                      "out_shape": list(out.shape)}))
    ops.set_conv_math(0)


if __name__ == "__main__":
    main()
