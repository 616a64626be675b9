#!/usr/bin/env python
"""Benchmark of the stixel hot path (BASELINE.json metric: stixel frames/sec @1024x2048).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload unary_b64|pairwise_b64|pairwise_w4_b64]
  python bench.py --impl reference ...      # the reference's own implementation, same workload

A "step" is one pass of the whole path (column join -> tables -> DP -> backtracking -> instance
grouping) over one batch of 64 synthetic Cityscapes-shaped frames per GPU.  Frames are independent,
so N GPUs run N shards with no collective in the data path (weak scaling).

  value : frames/s with inputs already resident in HBM (isx_compute_batch_device), timed with CUDA
          events on the stream the kernels are launched on, max over ranks.
  e2e   : the same batch through the host-buffer entry point (isx_compute_batch_host): pinned host
          inputs -> H2D -> kernels -> D2H of all Sections + instance records, every step.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # BASELINE.json configs[1]: unary model, batch of 64 frames on 1 B200
    "unary_b64": dict(mode="unary", step=8, batch=64),
    # configs[2]: pairwise model + instance grouping, batch 64
    "pairwise_b64": dict(mode="pairwise", step=8, batch=64),
    # configs[3]: pairwise at stixel width 4
    "pairwise_w4_b64": dict(mode="pairwise", step=4, batch=64),
    # configs[4]: the 4096-frame stream at the largest batch; --steps 4096 / (256 * N) covers the stream
    "pairwise_stream_b256": dict(mode="pairwise", step=8, batch=256),
}
ROWS, COLS = 1024, 2048
OPS_PER_CELL = {"unary": 103, "pairwise": 128}  # SURVEY.md 8d minimal op budget
# dram__bytes_read.sum + dram__bytes_write.sum per launch of a 32-frame chunk, from the `ncu --set full` captures
# summarised in profiles/r1j_{unary,pairwise}_b32.txt (width 8; no capture for width 4).
# "tables" = join_columns + column_tables + object_lut kernels.
NCU_CHUNK = 32
NCU_TRAFFIC = {
    ("unary", 8): dict(dp=(3.4238 + 0.1345) * 1e9,   # dp_unary_pruned_kernel
                       tables=(0.2685 + 0.0267 + 0.1555 + 2.1040 + 0.0338 + 4.2369) * 1e9),
    ("pairwise", 8): dict(dp=(12.9455 + 0.5673) * 1e9,   # dp_pairwise_walk_kernel (r1j): 2368 columns in flight, L2 hit 31 %
                          tables=(0.2685 + 0.0273 + 0.1559 + 2.1009 + 0.0338 + 4.2363) * 1e9),
}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return dict(hbm_gbs=d.get("hbm_gbs", 6650.0), sm_max_mhz=d.get("sm_max_mhz", 1965.0), source="measured")
    return dict(hbm_gbs=6650.0, sm_max_mhz=1965.0, source="fallback")


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (B200_PROFILING.md clocks line):
    NVML in-process every 10 ms, nvidia-smi as the fallback."""

    _REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index: int):
        self.index = index
        self.sm = []
        self.reasons = set()
        self.sm_max = None
        self.source = "nvml"
        self._stop = threading.Event()
        self._t = None
        self._h = None
        try:
            import pynvml
            import torch
            pynvml.nvmlInit()
            try:
                uuid = "GPU-" + str(torch.cuda.get_device_properties(index).uuid)
                self._h = pynvml.nvmlDeviceGetHandleByUUID(uuid)
            except Exception:
                self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self._nv = pynvml
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self._h = None
            self.source = "nvidia-smi"

    def _sample_nvml(self):
        nv = self._nv
        self.sm.append(float(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)))
        try:
            bits = nv.nvmlDeviceGetCurrentClocksEventReasons(self._h)
        except Exception:
            bits = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
        for bit, name in self._REASONS.items():
            if bits & bit:
                self.reasons.add(name)

    def _sample_smi(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                             capture_output=True, text=True, timeout=5).stdout
        parts = [x.strip() for x in out.strip().split(",")]
        if len(parts) >= 6:
            self.sm.append(float(parts[0]))
            self.sm_max = float(parts[1])
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], parts[2:6]):
                if v.lower().startswith("active"):
                    self.reasons.add(name)

    def _loop(self):
        while not self._stop.is_set():
            try:
                if self._h is not None:
                    self._sample_nvml()
                else:
                    self._sample_smi()
            except Exception:
                pass
            self._stop.wait(0.01 if self._h is not None else 0.2)

    def __enter__(self):
        self._t = threading.Thread(target=self._loop, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        if not self.sm:
            return dict(sm_mhz=None, sm_max_mhz=self.sm_max, reasons=["unavailable"], samples=0, source=self.source)
        sm = sorted(self.sm)
        return dict(sm_mhz=sm[len(sm) // 2], sm_min_mhz=sm[0], sm_max_mhz=self.sm_max, reasons=sorted(self.reasons),
                    samples=len(sm), source=self.source)


def cpu_baseline(wl, nframes=1):
    """The oracle port on the host cores, bounded sample of the same workload."""
    from instance_stixels_b200 import synth
    from oracle import cpubind
    pre = synth.preset(wl["mode"], ROWS, COLS, wl["step"])
    cfg = cpubind.default_config(**pre)
    lib = cpubind.load()
    cores = lib.orc_max_threads()
    frames = [synth.make_frame(i, rows=ROWS, cols=COLS, column_step=wl["step"]) for i in range(nframes)]
    t0 = time.perf_counter()
    for fr in frames:
        cpubind.compute_frame(cfg, wl["mode"] == "pairwise", fr.disparity, fr.segmentation, fr.road)
    dt = time.perf_counter() - t0
    return dict(value=nframes / dt, unit="frames/s", cores=cores, kind="port",
                sample=f"{nframes} frame(s) of the workload, oracle/stixels_cpu.cpp, OpenMP over columns")


def run_reference(args, wl, rank, world):
    """--impl reference: the reference's own implementation of the path.  The reference has no CPU
    path (all stages are __global__ kernels): its implementation IS the CUDA build, compiled
    unmodified for sm_100a into oracle/_ref and driven one frame per Compute() like
    apps/run_cityscapes.cu:346-431.  Falls back to the CPU port when oracle/_ref cannot be used."""
    if rank != 0:
        return
    from instance_stixels_b200 import api, synth
    from oracle import refbind
    pairwise = wl["mode"] == "pairwise"
    sample = min(wl["batch"], 16)
    line = dict(impl="reference", metric="stixel frames/sec @1024x2048", unit="frames/s", n_gpus=args.gpus,
                steps=args.steps, warmup=args.warmup, higher_is_better=True, scaling="weak", vs_baseline=None,
                dtype="f32", data="synthetic",
                config=dict(workload=args.workload, rows=ROWS, cols=COLS, column_step=wl["step"],
                            mode=wl["mode"], frames_per_step=sample))
    try:
        import torch
        if not (refbind.available() and torch.cuda.is_available()):
            raise RuntimeError("oracle/_ref or CUDA device not available")
        disp, seg, roads = synth.make_batch(sample, rows=ROWS, cols=COLS, column_step=wl["step"])
        ref = refbind.RefStixels(api.StixelConfig(**synth.preset(wl["mode"], ROWS, COLS, wl["step"])))
        for _ in range(max(args.warmup, 1)):
            ref.time_frames(pairwise, disp[:2], seg[:2], roads[0])
        t = 0.0
        for _ in range(args.steps):
            t += ref.time_frames(pairwise, disp, seg, roads[0])
        ref.close()
        fps = sample * args.steps / t
        line.update(value=fps, ms_per_step=1e3 * t / args.steps,
                    cpu_baseline=dict(value=fps, unit="frames/s", cores=0, kind="reference",
                                      sample=f"{sample} frames/step through the reference CUDA build "
                                             "(oracle/_ref, sm_100a) on GPU 0, one frame per Compute()"))
    except Exception as e:  # no GPU / no _ref: time the CPU port instead
        cb = cpu_baseline(wl, 1)
        line.update(value=cb["value"], ms_per_step=1e3 / cb["value"], cpu_baseline=cb, note=f"reference CUDA build unusable: {e!r}")
    line["e2e"] = dict(value=line["value"], unit="frames/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0)
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="unary_b64", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the latency / other-workload extras")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, wl, rank, world)
        return

    import torch
    import torch.distributed as dist
    from instance_stixels_b200 import api, synth

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the stixel path has no CPU fallback")
    torch.cuda.set_device(local)
    # Bind this process to the CPUs next to its GPU before any host buffer is allocated or pinned: the host <-> device
    # copies of the e2e arm then stay on the local NUMA node (what a multi-GPU launcher does per rank).
    host_cpus = None
    if not os.environ.get("ISX_BENCH_NO_AFFINITY"):
        try:
            import pynvml
            pynvml.nvmlInit()
            try:
                nvh = pynvml.nvmlDeviceGetHandleByUUID("GPU-" + str(torch.cuda.get_device_properties(local).uuid))
            except Exception:
                nvh = pynvml.nvmlDeviceGetHandleByIndex(local)
            pynvml.nvmlDeviceSetCpuAffinity(nvh)
            host_cpus = len(os.sched_getaffinity(0))
        except Exception:
            host_cpus = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    B = wl["batch"]
    pairwise = wl["mode"] == "pairwise"
    pre = synth.preset(wl["mode"], ROWS, COLS, wl["step"])
    st = api.make_stixels(pre, max_batch=B, device=local)
    C_ = st.GetRealCols()

    # synthetic frames: every rank gets its own shard of the stream (frame ids rank*B ...)
    disp, seg, roads = synth.make_batch(B, start=rank * B, rows=ROWS, cols=COLS, column_step=wl["step"])
    h_disp = torch.from_numpy(disp).pin_memory()
    h_seg = torch.from_numpy(seg).pin_memory()
    d_disp = h_disp.cuda(non_blocking=True)
    d_seg = h_seg.cuda(non_blocking=True)
    torch.cuda.synchronize()
    stream = torch.cuda.ExternalStream(st.stream(), device=local)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def device_step():
        st.ComputeBatchDevice(pairwise, B, d_disp.data_ptr(), d_seg.data_ptr(), roads)

    # ---- value: inputs resident in HBM ----
    for _ in range(max(args.warmup, 3)):
        device_step()
    st.Synchronize()
    st.set_profiling(True)
    st.stage_times(reset=True)
    launches0 = st._lib.isx_kernel_launch_count()
    units0 = st.dp_units()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    with ClockSampler(local) as clk:
        e0.record(stream)
        for _ in range(args.steps):
            device_step()
        st.Flush()  # the emission stream's last results are ordered before the closing event
        e1.record(stream)
        st.Synchronize()
        barrier()
    ms = e0.elapsed_time(e1)
    launches = st._lib.isx_kernel_launch_count() - launches0
    units1 = st.dp_units()
    units_eval, units_total = units1[0] - units0[0], units1[1] - units0[1]
    stages = st.stage_times(reset=True)
    st.set_profiling(False)
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())

    # ---- e2e: host buffers through the public batch API, copies inside the timed region ----
    # One host thread, one context, the streaming form of the batch call: isx_submit_batch_host enqueues a batch
    # (H2D of its inputs from pinned memory, kernels, D2H of all Sections and instance records) and
    # isx_wait_batch_host delivers the oldest one; two batches are in flight, so the first copy and the last
    # emission + copy-out of one batch hide behind the kernels of the other.  Every step moves its own inputs and
    # reads its own results back.  e2e_single is the synchronous call (isx_compute_batch_host), one batch at a time.
    sections_host = [torch.empty((B, C_, 200, 32), dtype=torch.uint8).pin_memory() for _ in range(2)]
    sec_np = [t.numpy().view(api.L.SECTION_DTYPE).reshape(B, C_, 200) for t in sections_host]
    n_inst_seen = [0]

    def run_e2e(pipelined):
        for w in range(2):  # warm-up of the path that is timed (the second result set is allocated on first use)
            if pipelined:
                st.SubmitBatch(pairwise, h_disp.numpy(), h_seg.numpy(), roads, sec_np[w])
                st.WaitBatch()
            else:
                st.ComputeBatch(pairwise, h_disp.numpy(), h_seg.numpy(), roads, sections_out=sec_np[w])
        barrier()
        t0 = time.perf_counter()
        if not pipelined:
            for _ in range(args.steps):
                _, inst, _ = st.ComputeBatch(pairwise, h_disp.numpy(), h_seg.numpy(), roads, sections_out=sec_np[0])
                n_inst_seen[0] = len(inst)
        else:
            for i in range(args.steps):
                st.SubmitBatch(pairwise, h_disp.numpy(), h_seg.numpy(), roads, sec_np[i & 1])
                if i > 0:
                    _, inst, _ = st.WaitBatch()
                    n_inst_seen[0] = len(inst)
            _, inst, _ = st.WaitBatch()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item())

    e2e_single_s = run_e2e(False)
    e2e_s = run_e2e(True)
    n_inst = n_inst_seen[0]

    if rank == 0:
        peaks = measured_peaks()
        H = ROWS
        cells_per_frame = C_ * H * (H + 1) // 2
        dp_ms, dp_launches = stages["dp"]
        chunk = st.chunk_frames()
        # The unary DP prunes (tile, chunk) units that provably cannot win (exact branch and bound): the roofline
        # counts the cells the kernel evaluated (32 x 32 per unit), not the cells of the exhaustive scan.
        eval_frac = units_eval / max(units_total, 1)
        cells_eval_per_launch = units_eval * 1024 / max(dp_launches, 1)
        ops_per_launch = cells_eval_per_launch * OPS_PER_CELL[wl["mode"]]
        dp_avg_s = dp_ms * 1e-3 / max(dp_launches, 1)
        achieved = ops_per_launch / dp_avg_s / 1e12
        peak = 148 * 128 * peaks["sm_max_mhz"] * 1e6 / 1e12
        # table build (join + column tables + object LUT).  Bytes the three kernels must move per frame in this
        # design: the inputs (disparity + unpadded segmentation), the joined disparity (written once,
        # read by two kernels), and the tables the DP consumes (prefix records in both layouts, object LUT).
        seg_bytes = C_ * 21 * (H // 8) * 4
        in_bytes = H * COLS * 4 + seg_bytes
        rec_stride = 1056  # kRecStride (common.cuh)
        table_bytes = C_ * H * 4 * 3 + C_ * (30 + 32) * rec_stride * 4 + C_ * 128 * H * 4
        tab_bytes = (in_bytes + table_bytes) * chunk
        tab_ms = stages["join"][0] + stages["column_tables"][0]
        tab_launches = max(stages["join"][1], 1)
        ncu = NCU_TRAFFIC.get((wl["mode"], wl["step"])) if chunk == NCU_CHUNK else None
        # DRAM bytes one DP launch has to move once: object LUT + records in both layouts + the (cost, vB) rows
        dp_alg_bytes = (C_ * 128 * H * 4 + C_ * (30 + 32) * rec_stride * 4 + C_ * H * 16) * chunk
        line = dict(
            metric="stixel frames/sec @1024x2048", value=world * B * args.steps / (ms_max * 1e-3), unit="frames/s",
            n_gpus=world, steps=args.steps, warmup=max(args.warmup, 3), ms_per_step=ms_max / args.steps,
            higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32", data="synthetic",
            config=dict(workload=args.workload, rows=ROWS, cols=COLS, column_step=wl["step"], mode=wl["mode"],
                        frames_per_step_per_gpu=B, chunk_frames=chunk,
                        l2="inputs (890 MB/step) and tables (>2 GB/chunk) exceed the 126 MB L2"),
            e2e=dict(value=world * B * args.steps / e2e_s, unit="frames/s",
                     # the zero padding of the segmentation tensor (rows/8 of rows_power2_segmentation entries
                     # per channel are used) does not travel
                     h2d_bytes_per_step=int(h_disp.numel() * 4 + h_seg.numel() * 4 * ((ROWS + 7) // 8) // h_seg.shape[-1]),
                     d2h_bytes_per_step=int(sections_host[0].numel() + B * st.instance_capacity() * 16 + B * 4 + 4),
                     pipeline="1 host thread, isx_submit_batch_host / isx_wait_batch_host, 2 batches in flight"),
            e2e_single=dict(value=world * B * args.steps / e2e_single_s, unit="frames/s",
                            pipeline="1 host thread, 1 context, synchronous batches"),
            gpu_launches=int(launches),
            host=dict(cpus_bound_to_gpu=host_cpus),
            clocks=clk.summary(),
            roofline=dict(bound="alu", kernel="dp_kernel", achieved=achieved, peak=peak, unit="Tlane-op/s",
                          frac=achieved / peak, traffic=ncu["dp"] if ncu else None,
                          traffic_note=f"DRAM bytes per launch (ncu, profiles/r1j_*_b32.txt); the tables one launch reads "
                                       f"once are {dp_alg_bytes} bytes",
                          units_evaluated_frac=eval_frac,
                          note=f"{OPS_PER_CELL[wl['mode']]} lane-ops per DP cell x {cells_eval_per_launch:.0f} cells "
                               f"evaluated per launch ({100 * eval_frac:.1f} % of the {cells_per_frame} x {chunk} cells "
                               f"of the exhaustive scan incl. the dead half of the diagonal units) / "
                               f"{dp_avg_s * 1e3:.2f} ms avg launch (CUDA events, {dp_launches} launches); "
                               f"peak = 148 SM x 128 lanes x {peaks['sm_max_mhz']:.0f} MHz ({peaks['source']})"),
            roofline_tables=dict(bound="hbm", kernel="join_columns+column_tables",
                                 achieved=tab_bytes / (tab_ms * 1e-3 / tab_launches) / 1e9,
                                 peak=peaks["hbm_gbs"], unit="GB/s",
                                 frac=tab_bytes / (tab_ms * 1e-3 / tab_launches) / 1e9 / peaks["hbm_gbs"],
                                 traffic=ncu["tables"] if ncu else None,
                                 note=f"bytes per launch {tab_bytes}; per frame: inputs {in_bytes} + joined/records/object LUT {table_bytes}; "
                                      f"inputs alone are {in_bytes * chunk / (tab_ms * 1e-3 / tab_launches) / 1e9:.0f} GB/s"),
            stage_ms_per_step={k: v[0] / args.steps for k, v in stages.items()},
        )
        if not args.no_extra and world == 1:
            # p50 latency of one frame through the reference call sequence (host in -> host out), with the
            # caller's buffers in pinned memory and -- like the reference's callers -- in pageable memory
            fr = synth.make_frame(0, rows=ROWS, cols=COLS, column_step=wl["step"])
            pin = [torch.from_numpy(fr.disparity).pin_memory(), torch.from_numpy(fr.segmentation).pin_memory(),
                   torch.empty(C_ * 200 * 32, dtype=torch.uint8).pin_memory()]

            def one_frame(disp, seg, out):
                t0 = time.perf_counter()
                st.SetDisparityImage(disp)
                st.SetSegmentation(seg)
                st.SetRoadParameters(**fr.road)
                st.Compute(pairwise, sections_out=out)
                st.GetInstanceStixels()
                return 1e3 * (time.perf_counter() - t0)

            torch.cuda.synchronize()
            lat = sorted([one_frame(pin[0].numpy(), pin[1].numpy(), pin[2].numpy().view(api.L.SECTION_DTYPE))
                          for _ in range(120)][40:])   # 40 warm-up frames: clocks and caches settle after the batches
            lat_pageable = sorted([one_frame(fr.disparity, fr.segmentation, None) for _ in range(50)][10:])
            line["latency_ms_batch1"] = dict(p50=lat[len(lat) // 2], p99=lat[-1], buffers="pinned",
                                             p50_pageable=lat_pageable[len(lat_pageable) // 2])
        if not args.no_cpu_baseline and world == 1:
            try:   # the CPU baseline gets every core of the box again
                os.sched_setaffinity(0, range(os.cpu_count()))
            except Exception:
                pass
            line["cpu_baseline"] = cpu_baseline(wl, 24)
        print(json.dumps(line), flush=True)
    st.Finish()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
