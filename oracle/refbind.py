"""TEST INFRASTRUCTURE -- ctypes driver of oracle/_ref/libref_stixels.so, the
UNMODIFIED reference CUDA sources compiled for sm_100a by oracle/Makefile
(plus oracle/ref_capi.cu and the DBSCAN stand-in).  Only tests/, the smoke
check and bench.py's reference arm import this; the product never does.
"""
from __future__ import annotations

import ctypes as C
import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(_HERE))
from instance_stixels_b200 import _lib as L  # noqa: E402  (struct layouts only)

REF_PATH = os.path.join(_HERE, "_ref", "libref_stixels.so")
_ref = None


def available() -> bool:
    return os.path.exists(REF_PATH)


def load():
    global _ref
    if _ref is None:
        lib = C.CDLL(REF_PATH)
        V = C.c_void_p
        lib.ref_create.restype = V
        lib.ref_destroy.argtypes = [V]
        lib.ref_config_init.argtypes = [C.POINTER(L.Config)]
        lib.ref_config_init.restype = None
        lib.ref_configure.argtypes = [V, C.POINTER(L.Config)]
        lib.ref_real_cols.argtypes = [V]
        lib.ref_max_sections.argtypes = [V]
        lib.ref_segmentation_elems.argtypes = [V]
        lib.ref_segmentation_elems.restype = C.c_size_t
        lib.ref_compute_frame.argtypes = [V, C.c_int, V, C.c_size_t, V, C.c_size_t, C.c_int, C.c_float,
                                          C.c_float, C.c_float, V, C.POINTER(L.FrameMeta)]
        lib.ref_time_frames.argtypes = [V, C.c_int, C.c_int, V, C.c_size_t, V, C.c_size_t, C.c_int, C.c_float,
                                        C.c_float, C.c_float]
        lib.ref_time_frames.restype = C.c_double
        lib.ref_time_frames_split.argtypes = lib.ref_time_frames.argtypes + [C.POINTER(C.c_double)]
        lib.ref_time_frames_split.restype = C.c_double
        lib.ref_num_instances.argtypes = [V]
        lib.ref_get_instances.argtypes = [V, V, C.c_int]
        lib.ref_tensor_elems.argtypes = [V, C.c_int]
        lib.ref_tensor_elems.restype = C.c_size_t
        lib.ref_read_tensor.argtypes = [V, C.c_int, V, C.c_size_t]
        lib.ref_read_ground_tables.argtypes = [V, V, V, V]
        lib.ref_read_init_tables.argtypes = [V, V, V, V]
        lib.ref_params_bytes.restype = C.c_size_t
        _ref = lib
    return _ref


class RefStixels:
    """SetConfig+Initialize once, then one frame per call like apps/run_cityscapes.cu:335-449."""

    def __init__(self, config: L.Config):
        self.lib = load()
        self.h = C.c_void_p(self.lib.ref_create())
        if self.lib.ref_configure(self.h, C.byref(config)) != 0:
            raise ValueError("reference SetConfig threw std::invalid_argument")
        self.realcols = self.lib.ref_real_cols(self.h)
        self.max_sections = self.lib.ref_max_sections(self.h)
        self.rows = int(config.rows)
        self.max_dis = int(config.max_dis)

    def close(self):
        if self.h:
            self.lib.ref_destroy(self.h)
            self.h = None

    def compute(self, pairwise: bool, disparity: np.ndarray, segmentation: np.ndarray, road: dict):
        d = np.ascontiguousarray(disparity, dtype=np.float32)
        s = np.ascontiguousarray(segmentation, dtype=np.int32)
        sections = np.zeros(self.realcols * self.max_sections, dtype=L.SECTION_DTYPE)
        meta = L.FrameMeta()
        rc = self.lib.ref_compute_frame(self.h, int(pairwise), d.ctypes.data, d.size, s.ctypes.data, s.size,
                                        int(road["vhor"]), road["camera_tilt"], road["camera_height"],
                                        road["alpha_ground"], sections.ctypes.data, C.byref(meta))
        if rc != 0:
            raise RuntimeError(f"reference Compute failed ({rc})")
        n = self.lib.ref_num_instances(self.h)
        inst = np.zeros(max(n, 1), dtype=L.INSTANCE_DTYPE)
        n = self.lib.ref_get_instances(self.h, inst.ctypes.data, n)
        return sections.reshape(self.realcols, self.max_sections), inst[:n], meta

    def time_frames(self, pairwise: bool, disparity: np.ndarray, segmentation: np.ndarray, road: dict) -> float:
        """Seconds for len(disparity) frames through the reference's public API (CUDA events)."""
        d = np.ascontiguousarray(disparity, dtype=np.float32)
        s = np.ascontiguousarray(segmentation, dtype=np.int32)
        n = d.shape[0]
        return self.lib.ref_time_frames(self.h, int(pairwise), n, d.ctypes.data, d[0].size, s.ctypes.data,
                                        s[0].size, int(road["vhor"]), road["camera_tilt"],
                                        road["camera_height"], road["alpha_ground"])

    def time_frames_split(self, pairwise: bool, disparity: np.ndarray, segmentation: np.ndarray, road: dict) -> dict:
        """Host-clock split of the same loop (every call of the sequence blocks): seconds over all frames."""
        d = np.ascontiguousarray(disparity, dtype=np.float32)
        s = np.ascontiguousarray(segmentation, dtype=np.int32)
        out = (C.c_double * 4)()
        total = self.lib.ref_time_frames_split(self.h, int(pairwise), d.shape[0], d.ctypes.data, d[0].size,
                                               s.ctypes.data, s[0].size, int(road["vhor"]), road["camera_tilt"],
                                               road["camera_height"], road["alpha_ground"], out)
        return dict(total=total, set_inputs=out[0], compute=out[1], get_instance_stixels=out[2],
                    dbscan_standin=out[3], frames=int(d.shape[0]))

    def read_tensor(self, tensor: int) -> np.ndarray:
        n = self.lib.ref_tensor_elems(self.h, tensor)
        out = np.zeros(n, dtype=np.float32)
        if self.lib.ref_read_tensor(self.h, tensor, out.ctypes.data, out.nbytes) != 0:
            raise RuntimeError("ref_read_tensor failed")
        return out

    def ground_tables(self):
        g = [np.zeros(self.rows, dtype=np.float32) for _ in range(3)]
        self.lib.ref_read_ground_tables(self.h, *[a.ctypes.data for a in g])
        return g

    def init_tables(self):
        D = self.max_dis
        lut = np.zeros(D * D, dtype=np.float32)
        rng = np.zeros(D, dtype=np.float32)
        params = np.zeros(self.lib.ref_params_bytes() // 4, dtype=np.float32)
        self.lib.ref_read_init_tables(self.h, lut.ctypes.data, rng.ctypes.data, params.ctypes.data)
        return lut.reshape(D, D), rng, params
