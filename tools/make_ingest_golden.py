"""Golden vectors for the segmentation ingest: the reference's own torch ops (FlipAndPad.forward,
tools/CNN_training/models/wrappers.py:50-60, minus the hard-wired 98-row index and .cuda()) on seeded inputs.
  python tools/make_ingest_golden.py -> tests/golden/ingest_*.npz"""
import math
import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def make_input(seed, hs, ws):
    rng = np.random.default_rng(seed)
    x = np.empty((21, hs, ws), dtype=np.float32)
    logits = rng.normal(0, 3, size=(19, hs, ws)).astype(np.float32)
    x[:19] = -torch.log_softmax(torch.from_numpy(logits), dim=0).numpy()
    x[19:] = rng.normal(0, 40, size=(2, hs, ws)).astype(np.float32)          # offsets: both signs, fractional
    return x


def flip_and_pad_torch(x):
    hs = x.shape[1]
    pad_rows = 2 ** (math.ceil(math.log2(hs + 1))) - hs
    t = torch.from_numpy(x).unsqueeze(0)
    t = t.permute(0, 3, 1, 2)
    t = torch.index_select(t, 3, torch.arange(hs - 1, -1, -1))
    t = torch.nn.functional.pad(t, value=0, pad=(0, pad_rows))
    t = t * 8
    return t.int()[0].numpy()


if __name__ == "__main__":
    for seed, hs, ws in ((0, 16, 24), (1, 25, 41), (2, 98, 224 // 4)):
        x = make_input(seed, hs, ws)
        path = os.path.join(ROOT, "tests", "golden", f"ingest_s{seed}.npz")
        np.savez_compressed(path, seed=seed, hs=hs, ws=ws, out=flip_and_pad_torch(x))
        print(path, x.shape)
