"""front-end throughput: one hour of 16 kHz audio -> normalised 80-bin mel spectrogram on the GPU"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lcasr_b200.frontend import to_spectogram
dev = torch.device("cuda", 0)
for minutes in (20, 60):
    wav = (0.1 * torch.randn(1, int(minutes * 60 * 16000))).to(dev)
    for _ in range(2):
        to_spectogram(wav)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        out = to_spectogram(wav)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    frames = out.shape[-1]
    print(json.dumps({"minutes": minutes, "frames": frames, "ms": ms, "audio_s_per_s": minutes * 60 / (ms / 1e3),
                      "dft_tflops_fp32": 2.0 * frames * 512 * 257 * 2 / (ms / 1e3) / 1e12}))
