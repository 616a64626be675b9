"""CPU restatement of the reference's road estimation (test infrastructure, not a product path).

Follows InstanceStixels/src/RoadEstimationKernels.cu:25-60 (v-disparity histogram, maximum, binary image),
RoadEstimation.cu:93-193 (Compute, ComputeHough, ComputeCameraProperties) and -- for the one third-party
step -- OpenCV's cv::HoughLines (standard transform; modules/imgproc/src/hough.cpp HoughLinesStandard,
createTrigTable, findLocalMaximums, hough_cmp_gt; the reference pins no OpenCV version, the restatement is
validated against the cv2 4.13 wheel of this image in tests/test_road_oracle.py, fixtures under tests/golden).

Only tests/, __graft_entry__.smoke() and bench.py's baseline legs may import this module.
"""
from __future__ import annotations

import ctypes
import ctypes.util
import math

import numpy as np

# the reference's host code calls glibc's sinf / cosf / atanf (RoadEstimation.cu:176-192)
_libm = ctypes.CDLL(ctypes.util.find_library("m") or "libm.so.6")
for _n in ("sinf", "cosf", "atanf"):
    getattr(_libm, _n).restype = ctypes.c_float
    getattr(_libm, _n).argtypes = [ctypes.c_float]

HOUGH_ACCUM_THRESHOLD = 25            # RoadEstimation.cu:45
MAX_PITCH = np.float32(np.float32(50) * np.float32(math.pi) / np.float32(180.0))   # :47-58
MIN_PITCH = np.float32(np.float32(-50) * np.float32(math.pi) / np.float32(180.0))


def vdisparity(disp: np.ndarray, max_dis: int) -> np.ndarray:
    """ComputeHistogram (RoadEstimationKernels.cu:25-38): per image row, histogram of (int)d over d != 0.
    Values outside [0, max_dis) would write out of bounds in the reference; they are ignored here."""
    rows, _ = disp.shape
    d = disp.astype(np.float32)
    col = np.trunc(d).astype(np.int64)
    ok = (d != 0) & (col >= 0) & (col < max_dis)
    out = np.zeros((rows, max_dis), dtype=np.int32)
    r = np.broadcast_to(np.arange(rows)[:, None], d.shape)
    np.add.at(out, (r[ok], col[ok]), 1)
    return out


def binary_image(vdisp: np.ndarray, threshold: float) -> np.ndarray:
    """ComputeMaximum + ComputeBinaryImage (RoadEstimationKernels.cu:41-60): 255 where (float)p > max * thr."""
    mx = np.float32(vdisp.max())
    return np.where(vdisp.astype(np.float32) > mx * np.float32(threshold), 255, 0).astype(np.uint8)


def trig_tables(numangle: int, theta: np.float32, irho: np.float32):
    """createTrigTable: the angle is accumulated in float, sin/cos evaluated in double."""
    tab_sin = np.empty(numangle, dtype=np.float32)
    tab_cos = np.empty(numangle, dtype=np.float32)
    ang = np.float32(0.0)
    for n in range(numangle):
        tab_sin[n] = np.float32(math.sin(float(ang)) * float(irho))
        tab_cos[n] = np.float32(math.cos(float(ang)) * float(irho))
        ang = np.float32(ang + theta)
    return tab_sin, tab_cos


def hough_numangle(theta: float) -> int:
    """computeNumangle(0, CV_PI, theta) of OpenCV 4.x; 180 for theta = CV_PI/180 in every version."""
    n = int(math.floor(math.pi / theta)) + 1
    if n > 1 and abs(math.pi - (n - 1) * theta) < theta / 2:
        n -= 1
    return n


def hough_lines(binary: np.ndarray, threshold: int = HOUGH_ACCUM_THRESHOLD):
    """cv::HoughLines(img, lines, 1.0, CV_PI/180, threshold) -> [(rho, theta)] sorted like OpenCV
    (votes descending, accumulator index ascending), plus the accumulator for the tests."""
    height, width = binary.shape
    theta = np.float32(math.pi / 180)
    numangle = hough_numangle(float(theta))
    max_rho = width + height
    numrho = int(round(((2 * max_rho) + 1) / 1.0))
    tab_sin, tab_cos = trig_tables(numangle, theta, np.float32(1.0))
    accum = np.zeros((numangle + 2, numrho + 2), dtype=np.int32)
    ii, jj = np.nonzero(binary)
    if len(ii):
        # r = cvRound(j * tabCos[n] + i * tabSin[n]): float products, float sum, round half to even
        v = (jj.astype(np.float32)[:, None] * tab_cos[None, :]).astype(np.float32) + \
            (ii.astype(np.float32)[:, None] * tab_sin[None, :]).astype(np.float32)
        r = np.rint(v.astype(np.float32)).astype(np.int64) + (numrho - 1) // 2
        n_idx = np.broadcast_to(np.arange(numangle)[None, :], r.shape)
        np.add.at(accum, (n_idx + 1, r + 1), 1)
    # findLocalMaximums: > threshold, > left, >= right, > previous angle, >= next angle
    c = accum[1:-1, 1:-1]
    cand = (c > threshold) & (c > accum[1:-1, :-2]) & (c >= accum[1:-1, 2:]) & (c > accum[:-2, 1:-1]) & \
           (c >= accum[2:, 1:-1])
    nn, rr = np.nonzero(cand)
    base = (nn + 1) * (numrho + 2) + rr + 1
    order = np.lexsort((base, -c[nn, rr].astype(np.int64)))
    lines = []
    for k in order:
        n, r = int(nn[k]), int(rr[k])
        rho = np.float32((np.float32(r) - np.float32(numrho - 1) * np.float32(0.5)) * np.float32(1.0))
        ang = np.float32(np.float32(0.0) + np.float32(n) * theta)
        lines.append((rho, ang, int(c[n, r])))
    return lines, accum


def camera_properties(rho, theta, rows, cy, baseline, focal):
    """ComputeCameraProperties (RoadEstimation.cu:172-193), fp32."""
    f = np.float32
    sinf = lambda x: f(_libm.sinf(float(x)))
    cosf = lambda x: f(_libm.cosf(float(x)))
    atanf = lambda x: f(_libm.atanf(float(x)))
    horizon = f(f(rho) / sinf(theta))
    pitch = f(-atanf(f(f(f(cy) - horizon) / f(focal))))
    last_row = f(rows - 1)
    vdisp_down = f(f(f(rho) - f(last_row * sinf(theta))) / cosf(theta))
    slope = f(f(f(0) - vdisp_down) / f(horizon - last_row))
    height = f(f(f(baseline) * cosf(pitch)) / slope)
    return horizon, pitch, height, slope


def estimate(disp: np.ndarray, max_dis: int, cy: float, baseline: float, focal: float,
             vdisparity_threshold: float = 0.2):
    """RoadEstimation::Compute (RoadEstimation.cu:107-137).  Returns dict(ok, horizon_point, pitch,
    camera_height, slope, rho, theta) and the intermediates."""
    rows = disp.shape[0]
    vd = vdisparity(disp, max_dis)
    binary = binary_image(vd, vdisparity_threshold)
    lines, accum = hough_lines(binary)
    out = dict(ok=False, horizon_point=0, pitch=np.float32(0), camera_height=np.float32(0), slope=np.float32(0),
               rho=np.float32(0), theta=np.float32(0))
    with np.errstate(divide="ignore", invalid="ignore"):
        for rho, theta, _votes in lines:
            rho = np.float32(abs(rho))
            horizon, pitch, height, slope = camera_properties(rho, theta, rows, cy, baseline, focal)
            if pitch >= MIN_PITCH and pitch <= MAX_PITCH:       # RoadEstimation.cu:163
                out = dict(ok=True, horizon_point=int(math.ceil(float(horizon))), pitch=pitch, camera_height=height,
                           slope=slope, rho=rho, theta=theta)
                break
    return out, dict(vdisp=vd, binary=binary, accum=accum, lines=lines)
