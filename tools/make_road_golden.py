"""Golden vectors for the road estimation: the reference pipeline with the REAL cv2.HoughLines of this image
(OpenCV 4.13) on seeded synthetic frames.  Run in the build container (cv2 is not assumed on the GPU box):
  python tools/make_road_golden.py      -> tests/golden/road_*.npz
Each file holds the frame recipe (the disparity image is regenerated from the seed), the first 64 lines
cv2.HoughLines returned for the binary v-disparity image, and the resulting camera properties."""
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from instance_stixels_b200 import synth  # noqa: E402
from oracle import road_cpu  # noqa: E402

CASES = [("small", 256, 512, 0), ("small", 256, 512, 7), ("ragged", 200, 328, 3), ("full", 1024, 2048, 1),
         ("tilted", 512, 1024, 5)]


def frame_disparity(name, rows, cols, frame):
    disp = synth.make_frame(frame, rows=rows, cols=cols).disparity.copy()
    if name == "tilted":   # a wall: many rows with one disparity -> a vertical v-disparity line (theta = 0) wins first
        disp[: rows // 2, : cols // 2] = 37.25
    return disp


def main():
    for name, rows, cols, frame in CASES:
        disp = frame_disparity(name, rows, cols, frame)
        vd = road_cpu.vdisparity(disp, 128)
        binary = road_cpu.binary_image(vd, 0.2)
        lines = cv2.HoughLines(binary, 1.0, np.pi / 180, road_cpu.HOUGH_ACCUM_THRESHOLD)
        lines = np.zeros((0, 2), np.float32) if lines is None else lines.reshape(-1, 2)
        est = dict(ok=False, horizon_point=0, pitch=0.0, camera_height=0.0, slope=0.0, rho=0.0, theta=0.0)
        with np.errstate(divide="ignore", invalid="ignore"):
            for rho, theta in lines:
                rho = np.float32(abs(rho))
                h, p, ch, s = road_cpu.camera_properties(rho, theta, rows, 512.0, 0.209313, 2262.52)
                if road_cpu.MIN_PITCH <= p <= road_cpu.MAX_PITCH:
                    est = dict(ok=True, horizon_point=int(np.ceil(h)), pitch=p, camera_height=ch, slope=s, rho=rho,
                               theta=theta)
                    break
        path = os.path.join(ROOT, "tests", "golden", f"road_{name}_f{frame}.npz")
        np.savez_compressed(path, name=name, rows=rows, cols=cols, frame=frame, max_dis=128,
                            n_lines=len(lines), lines=lines[:64], vdisp_sum=int(vd.sum()), binary_count=int((binary > 0).sum()),
                            ok=est["ok"], horizon_point=est["horizon_point"],
                            floats=np.array([est["pitch"], est["camera_height"], est["slope"], est["rho"], est["theta"]],
                                            dtype=np.float32),
                            cv2_version=cv2.__version__)
        print(path, len(lines), est)


if __name__ == "__main__":
    main()
