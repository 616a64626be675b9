"""CPU restatement of the reference's segmentation layout transform `FlipAndPad`
(tools/CNN_training/models/wrappers.py:35-61).  Test infrastructure: only tests/ may import it.
Pinned by tests/golden/ingest_*.npz, produced with the reference's own torch ops (tools/make_ingest_golden.py)."""
from __future__ import annotations

import math

import numpy as np


def rows_power2(hs: int) -> int:
    return 2 ** math.ceil(math.log2(hs + 1))   # wrappers.py:40-41, Stixels.cu:132-133


def flip_and_pad(cnn: np.ndarray, column_step: int = 8) -> np.ndarray:
    """cnn float32 [21][Hs][Ws] (rows top-down) -> int32 [C][21][Hs2], C = Ws * 8 / column_step:
    permute(0,3,1,2), index_select(rows reversed), pad(0, pad_rows), x *= 8, x.int()  (wrappers.py:50-60)."""
    ch, hs, ws = cnn.shape
    hs2 = rows_power2(hs)
    x = np.transpose(cnn.astype(np.float32), (2, 0, 1))[:, :, ::-1]          # [Ws][21][Hs], rows flipped
    out = np.zeros((ws, ch, hs2), dtype=np.float32)
    out[:, :, :hs] = x
    out = np.trunc(out * np.float32(8.0)).astype(np.int32)                    # fp32 multiply, truncation
    per = 8 // column_step                                                    # width 4
    return np.repeat(out, per, axis=0) if per > 1 else out
