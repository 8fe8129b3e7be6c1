import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
RAGGED_CASES = ["tiny_lengths", "tiny_lengths_dh128"]  # batches with per-recording lengths (key-padding mask path)
GOLDEN_CASES = ["tiny_dh32_ragged", "tiny_dh128", "tiny_rms_nosc", "tiny_norotary", "cfg1_6L256D8H", "cfg1_peaky"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200, sm_100a); run with -m gpu on the GPU box")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
    g = {k: z[k] for k in z.files}
    g["config"] = json.loads(str(g["config"]))
    g["greedy"] = [json.loads(s) for s in g["greedy"].tolist()]
    if "state_dict_shapes" in g:
        g["state_dict_shapes"] = json.loads(str(g["state_dict_shapes"]))
    for k in ("batch", "frames", "weight_seed", "input_seed", "target_seed", "init_seed"):
        if k in g:
            g[k] = int(g[k])
    if "peak" in g:
        g["peak"] = float(g["peak"])
    g["frame_lengths"] = g["frame_lengths"].tolist() if "frame_lengths" in g else [g["frames"]] * g["batch"]
    return g


@pytest.fixture(scope="session")
def cuda_device():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda", 0)
