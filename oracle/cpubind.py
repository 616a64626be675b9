"""TEST INFRASTRUCTURE -- ctypes driver of oracle/liboracle_cpu.so (the host
C++/OpenMP restatement in stixels_cpu.cpp).  Imported only by tests/, the smoke
check and bench.py's cpu_baseline leg."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(_HERE))
from instance_stixels_b200 import _lib as L  # noqa: E402  (struct layouts only)

CPU_PATH = os.path.join(_HERE, "liboracle_cpu.so")
_cpu = None


def build():
    subprocess.check_call(["make", "-C", _HERE, "cpu"], stdout=subprocess.DEVNULL)


def load():
    global _cpu
    if _cpu is None:
        if not os.path.exists(CPU_PATH):
            build()
        lib = C.CDLL(CPU_PATH)
        V = C.c_void_p
        lib.orc_compute_frame.argtypes = [C.POINTER(L.Config), C.c_int, V, V, C.POINTER(L.Road), C.c_int, V, V,
                                          C.c_int, C.POINTER(C.c_int), V, V, V, V]
        lib.orc_max_threads.restype = C.c_int
        lib.orc_dbscan.argtypes = [V, C.c_int, C.c_float, C.c_int, V, V]
        lib.orc_dbscan.restype = None
        _cpu = lib
    return _cpu


def default_config(**fields) -> L.Config:
    """StixelConfig defaults (types.h:30-141) without touching the CUDA library."""
    c = L.Config()
    for n, _ in L.Config._fields_:
        setattr(c, n, -1)
    c.invalid_disparity = -1.0
    c.pairwise = 0
    c.sigma_disparity_object, c.sigma_disparity_ground, c.sigma_sky = 1.0, 2.0, 0.1
    c.pout, c.pout_sky, c.pord, c.pgrav, c.pblg = 0.15, 0.4, 0.2, 0.1, 0.04
    c.pground_given_nexist, c.pobject_given_nexist, c.psky_given_nexist = 0.28, 0.44, 0.28
    c.pnexist_dis = 0.25
    c.pground = c.pobject = c.psky = float(np.float32(1.0) / np.float32(3.0))
    c.width_margin = 0
    c.sigma_camera_tilt = c.sigma_camera_height = 0.05
    c.median_join = 0
    c.epsilon, c.range_objects_z, c.road_vdisparity_threshold = 3.0, 10.20, 0.2
    for k, v in fields.items():
        setattr(c, k, int(v) if isinstance(v, bool) else v)
    return c


def compute_frame(config: L.Config, pairwise: bool, disparity: np.ndarray, segmentation: np.ndarray, road: dict,
                  nthreads: int = 0, tables: bool = False):
    """Returns (sections [C][200], instances, extras dict)."""
    lib = load()
    H, W = int(config.rows), int(config.cols)
    Cc = (W - config.width_margin) // config.column_step
    D = config.max_dis
    d = np.ascontiguousarray(disparity, dtype=np.float32)
    s = np.ascontiguousarray(segmentation, dtype=np.int32)
    sections = np.zeros((Cc, 200), dtype=L.SECTION_DTYPE)
    inst = np.zeros(Cc * 200, dtype=L.INSTANCE_DTYPE)
    n = C.c_int(0)
    r = L.Road(int(road["vhor"]), road["camera_tilt"], road["camera_height"], road["alpha_ground"])
    ex = {}
    ptr = [None] * 4
    if tables:
        ex["cost_table"] = np.zeros((Cc, H, 3), np.float32)
        ex["index_table"] = np.zeros((Cc, H, 3), np.int32)
        ex["joined"] = np.zeros((Cc, H), np.float32)
        ex["object_lut"] = np.zeros((Cc, D, H + 1), np.float32)
        ptr = [ex[k].ctypes.data for k in ("cost_table", "index_table", "joined", "object_lut")]
    lib.orc_compute_frame(C.byref(config), int(pairwise), d.ctypes.data, s.ctypes.data, C.byref(r), nthreads,
                          sections.ctypes.data, inst.ctypes.data, inst.size, C.byref(n), *ptr)
    return sections, inst[:n.value], ex


def dbscan(xy: np.ndarray, eps: float, min_pts: int, core_candidate: np.ndarray) -> np.ndarray:
    lib = load()
    xy = np.ascontiguousarray(xy, dtype=np.float32)
    cand = np.ascontiguousarray(core_candidate, dtype=np.uint8)
    labels = np.zeros(len(xy), dtype=np.int32)
    lib.orc_dbscan(xy.ctypes.data, len(xy), eps, min_pts, cand.ctypes.data, labels.ctypes.data)
    return labels
