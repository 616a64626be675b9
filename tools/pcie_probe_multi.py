"""Concurrent host <-> device copy rates of the box: N processes (one per GPU, bound to the CPUs next to their GPU
like bench.py) copy pinned buffers at the same time, for N = 1, 2, 4, 8 (as many as the box has).  Names the limit of
the end-to-end arm: per-GPU link rate vs what the host side sustains when every GPU copies at once.

  python tools/pcie_probe_multi.py [--seconds 1.5] [--mb 512] > gpurun_out/pcie_probe_multi.json
"""
import argparse
import json
import os
import subprocess
import sys
import time


def worker(idx, start_at, seconds, mb, total):
    """All workers start once; the phases (direction, N) follow a wall-clock schedule, worker idx takes part in a
    phase when idx < N."""
    import torch
    torch.cuda.set_device(idx)
    try:
        import pynvml
        pynvml.nvmlInit()
        try:
            h = pynvml.nvmlDeviceGetHandleByUUID("GPU-" + str(torch.cuda.get_device_properties(idx).uuid))
        except Exception:
            h = pynvml.nvmlDeviceGetHandleByIndex(idx)
        pynvml.nvmlDeviceSetCpuAffinity(h)
    except Exception:
        pass
    n = mb * 1024 * 1024
    host = torch.empty(n, dtype=torch.uint8).pin_memory()
    host.fill_(1)
    dev = torch.empty(n, dtype=torch.uint8, device="cuda")
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        dev.copy_(host, non_blocking=True)
        host.copy_(dev, non_blocking=True)
        s.synchronize()
        for k, (direction, count) in enumerate(schedule(total)):
            t_phase = start_at + k * (seconds + 1.0)
            if idx >= count:
                continue
            while time.time() < t_phase:
                time.sleep(0.0005)
            t0 = time.perf_counter()
            reps = 0
            while time.perf_counter() - t0 < seconds:
                for _ in range(2):
                    if direction == "h2d":
                        dev.copy_(host, non_blocking=True)
                    else:
                        host.copy_(dev, non_blocking=True)
                s.synchronize()
                reps += 2
            dt = time.perf_counter() - t0
            print(json.dumps(dict(phase=k, gpu=idx, gbs=reps * n / dt / 1e9, cpus=len(os.sched_getaffinity(0)))), flush=True)


def schedule(total):
    counts = [k for k in (1, 2, 4, 8) if k <= total]
    return [("h2d", k) for k in counts] + [("d2h", k) for k in counts]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--worker", type=int, default=-1)
    ap.add_argument("--start-at", type=float, default=0.0)
    ap.add_argument("--seconds", type=float, default=1.0)
    ap.add_argument("--mb", type=int, default=512)
    ap.add_argument("--total", type=int, default=0)
    a = ap.parse_args()
    if a.worker >= 0:
        worker(a.worker, a.start_at, a.seconds, a.mb, a.total)
        return
    import torch
    total = torch.cuda.device_count()
    start_at = time.time() + 20.0      # CUDA context creation + pinning of every worker
    procs = [subprocess.Popen([sys.executable, os.path.abspath(__file__), "--worker", str(i), "--start-at", repr(start_at),
                               "--seconds", str(a.seconds), "--mb", str(a.mb), "--total", str(total)],
                              stdout=subprocess.PIPE, text=True) for i in range(total)]
    res = []
    for p in procs:
        o, _ = p.communicate()
        res += [json.loads(line) for line in o.splitlines() if line.startswith("{")]
    out = dict(gpus_in_box=total, cpus=os.cpu_count(), buffer_mb=a.mb, seconds_per_phase=a.seconds, runs=[])
    for k, (direction, count) in enumerate(schedule(total)):
        rates = [r["gbs"] for r in sorted(res, key=lambda r: r["gpu"]) if r["phase"] == k]
        out["runs"].append(dict(direction=direction, gpus=count, per_gpu_gbs=[round(x, 1) for x in rates],
                                aggregate_gbs=round(sum(rates), 1)))
    out["cpus_per_worker"] = sorted({r["cpus"] for r in res})
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
