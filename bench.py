#!/usr/bin/env python
"""bench.py -- frames/s of the aruco_detect hot path (detect + pose) on BASELINE.json's config C2.

  python bench.py --gpus N --steps K --warmup W            our arm (CUDA path through the C-ABI)
  python bench.py --impl reference --gpus N --steps K ...   the reference's CPU path (cv2) on host cores

A "step" is one pass of the hot path over one batch of synthetic 1920x1080 frames (16 markers of
DICT_6X6_250 each): BGR8 frames -> ids, corners, rvec/tvec, quaternion, image/object error, area,
followed by the per-camera fiducial_slam map update of those messages (and, for N > 1, one NCCL
all-gather + deterministic merge of the per-rank map tables).  Frames shard one stream per GPU,
weak scaling, no collective on the detect/pose path.

  value  : frames/s, whole job, frames already resident in HBM when the timed region starts
  e2e    : frames/s through the same C-ABI call with pinned HOST frames (H2D of every frame and D2H
           of every result inside the timed region)
  roofline: the threshold kernel (the HBM-bound stage BASELINE.json's metric names): algorithmic
           bytes 3*W*H + n_scales*W*H/8 per frame / its CUDA-event time, vs the measured
           copy bandwidth in MEASURED_PEAKS.json
  cpu_baseline: the same frames through cv2's ArucoDetector + solvePnP + projectPoints (the OpenCV
           calls of aruco_detect.cpp:350,247,210) on this box's host cores, bounded sample
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = "C2"
FIDUCIAL_LEN = 0.14
FRAMES_PER_STEP = int(os.environ.get("FID_BENCH_FRAMES", "128"))  # distinct frames per step; 128 x 6.2 MB = 796 MB >> 126 MB L2
DEPTH = int(os.environ.get("FID_BENCH_DEPTH", "2"))            # batches in flight (each takes FRAMES/SLOT of the library's FID_SLOTS chunk slots)
SLOT_FRAMES = int(os.environ.get("FID_BENCH_SLOT", "64"))     # frames per in-flight chunk inside the library (two chunks pipeline)
HBM_FALLBACK_GBS = 6650.0


# ------------------------------------------------------------------------------------------------
def rank_info():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


class ClockSampler:
    """nvidia-smi clocks/throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                      stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=2)
        except Exception:
            self.p.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


def measured_hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)"
    except Exception:
        return HBM_FALLBACK_GBS, "fallback 6.65 TB/s (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------
# CPU reference arm (cv2): exactly the per-frame calls of the reference node
# ------------------------------------------------------------------------------------------------
_W = {}


def _cpu_worker_init(nthreads, npy_path, dict_id, K, D):
    """Worker of the throughput-mode CPU baseline (spawned, so OpenCV's thread pool is never forked)."""
    sys.path.insert(0, ROOT)
    import cv2

    cv2.setNumThreads(nthreads)
    _W["frames"] = np.load(npy_path, mmap_mode="r")
    _W["cfg"] = (dict_id, K, D)


def _cpu_worker_frame(i):
    from oracle import aruco_oracle as ao

    dict_id, K, D = _W["cfg"]
    fr = _W["frames"]
    ids, corners, rv, tv, fields = ao.detect_and_pose(np.ascontiguousarray(fr[i % len(fr)]), dict_id, K, D, FIDUCIAL_LEN)
    return len(ids)


class CpuArm:
    """The reference's CPU path on a bounded sample, in both modes:
    reference mode  = one process, OpenCV threads = all cores (what the single-threaded node does);
    throughput mode = single-threaded worker processes, frames round-robin: one per logical core and one per
                      physical core (half of them on an SMT box -- the detector is memory bound and ran 30 % faster
                      that way on the 128-thread host of the B200 box).
    The worker pools are spawned once (spawn, so OpenCV's thread pool is never forked) and reused by every
    measure() call: --impl reference times K steps without paying K pool start-ups."""

    def __init__(self, frames, dict_id, K, D):
        import multiprocessing as mp
        import tempfile

        self.frames, self.dict_id, self.K, self.D = frames, dict_id, K, D
        self.ncores = os.cpu_count() or 1
        self.pools = []
        self.tmp = tempfile.NamedTemporaryFile(suffix=".npy", delete=False)
        np.save(self.tmp, np.ascontiguousarray(frames))
        self.tmp.close()
        try:
            ctx = mp.get_context("spawn")
            for npr in [max(1, self.ncores)] + ([self.ncores // 2] if self.ncores >= 8 else []):
                pool = ctx.Pool(npr, initializer=_cpu_worker_init, initargs=(1, self.tmp.name, dict_id, K, D))
                pool.map(_cpu_worker_frame, range(npr), chunksize=1)  # warm-up (imports, first-call setup)
                self.pools.append((npr, pool))
        except Exception as e:  # pragma: no cover
            print("throughput-mode CPU baseline unavailable: %r" % (e,), file=sys.stderr)

    def close(self):
        for _, pool in self.pools:
            pool.terminate()
            pool.join()
        self.pools = []
        try:
            os.unlink(self.tmp.name)
        except OSError:
            pass

    def measure(self, budget_s=12.0):
        """dict(value=best fps, cores, kind, sample)."""
        import cv2

        from oracle import aruco_oracle as ao

        frames, ncores = self.frames, self.ncores

        def one(i):
            return ao.detect_and_pose(frames[i % len(frames)], self.dict_id, self.K, self.D, FIDUCIAL_LEN)

        n_modes = 1 + len(self.pools)
        # reference mode
        cv2.setNumThreads(ncores)
        one(0)
        t0 = time.perf_counter()
        n_ref = 0
        while (time.perf_counter() - t0 < budget_s / n_modes and n_ref < 4 * len(frames)) or n_ref < 3:
            one(n_ref)
            n_ref += 1
        fps_ref = n_ref / (time.perf_counter() - t0)
        # throughput mode
        fps_thr, n_thr, nproc, tried = 0.0, 0, max(1, ncores), []
        for npr, pool in self.pools:
            try:
                t0 = time.perf_counter()
                pool.map(_cpu_worker_frame, range(npr), chunksize=1)
                per_round = time.perf_counter() - t0
                rounds = int(max(1, min(8, (budget_s / n_modes) / max(per_round, 1e-3))))
                n = npr * rounds
                t0 = time.perf_counter()
                pool.map(_cpu_worker_frame, range(n), chunksize=1)
                fps = n / (time.perf_counter() - t0)
            except Exception as e:  # pragma: no cover
                print("throughput-mode CPU baseline failed: %r" % (e,), file=sys.stderr)
                continue
            tried.append("%d procs %.2f fps" % (npr, fps))
            if fps > fps_thr:
                fps_thr, n_thr, nproc = fps, n, npr
        best = max(fps_ref, fps_thr)
        return {
            "value": best,
            "unit": "frames/s",
            "cores": ncores if fps_ref >= fps_thr else nproc,
            "kind": "reference",
            "sample": "cv2 %s ArucoDetector(reference params)+solvePnP+projectPoints on %s frames: reference mode (1 proc, %d OpenCV threads) %d frames %.2f fps; "
                      "throughput mode (%d procs x 1 thread) %d frames %.2f fps [%s]; value = best of all; os.cpu_count=%d"
                      % (cv2.__version__, WORKLOAD, ncores, n_ref, fps_ref, nproc, n_thr, fps_thr, ", ".join(tried), ncores),
        }


def cpu_reference_fps(frames, dict_id, K, D, budget_s=12.0):
    arm = CpuArm(frames, dict_id, K, D)
    try:
        return arm.measure(budget_s)
    finally:
        arm.close()


def run_reference_arm(args):
    rank, local_rank, world = rank_info()
    if rank != 0:
        return  # rank 0 alone runs and prints the CPU arm
    from fiducials_b200 import synth

    frames, truths, K, D, dict_id = synth.make_config_stream(WORKLOAD, 8, seed=0)
    per_step = []
    detail = None
    budget = max(4.0, min(20.0, 150.0 / max(1, args.steps + args.warmup)))  # whole run: a few minutes
    arm = CpuArm(frames, dict_id, K, D)
    try:
        for i in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            detail = arm.measure(budget_s=budget)
            if i >= args.warmup:
                per_step.append((detail["value"], time.perf_counter() - t0))
    finally:
        arm.close()
    fps = float(np.median([p[0] for p in per_step]))
    ms = float(np.mean([p[1] for p in per_step]) * 1e3)
    detail["value"] = fps
    out = {
        "impl": "reference",
        "metric": "frames/sec 1920x1080 (detect+pose)",
        "value": fps,
        "unit": "frames/s",
        "n_gpus": args.gpus,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": ms,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "u8/f32/f64",
        "data": "synthetic",
        "config": {"workload": "C2: 1920x1080 BGR8 stream, 16 markers/frame, DICT_6X6_250, detect+pose; each step = bounded sample of the stream on host cores"},
        "cpu_baseline": detail,
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(out))


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_gpu_arm(args):
    import torch

    rank, local_rank, world = rank_info()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the hot path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod

        dist = dist_mod
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"  # "NCCL version ..." goes to stdout, which must hold the one JSON line only
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from fiducials_b200 import _lib, synth
    from fiducials_b200.multigpu import allgather_tables
    from fiducials_b200.node import MAXM, Detector, FiducialSlam, default_params

    lib = _lib.load()
    W, H, n_markers, dict_id = synth.CONFIGS[WORKLOAD]
    nf = FRAMES_PER_STEP
    frames, truths, K, D, _ = synth.make_config_stream(WORKLOAD, nf, seed=rank, realizations=8)
    det = Detector(default_params(dictionary=dict_id), local_rank, W, H, SLOT_FRAMES)
    slam = FiducialSlam(device=local_rank, max_fiducials=512, n_instances=1)
    ident = [0, 0, 0, 0, 0, 0, 1]

    # pinned host copy + device-resident copy of the stream
    hptr = C.c_void_p()
    _lib.check(lib.fid_host_alloc(frames.nbytes, C.byref(hptr)))
    pinned = np.ctypeslib.as_array(C.cast(hptr, C.POINTER(C.c_uint8)), shape=(frames.nbytes,)).reshape(frames.shape)
    pinned[...] = frames
    dptr = C.c_void_p()
    _lib.check(lib.fid_device_alloc(det.h, frames.nbytes, C.byref(dptr)))
    _lib.check(lib.fid_memcpy_h2d(det.h, dptr, hptr, frames.nbytes))

    launches = [0]

    def submit(on_device):
        # one batch (= one step's frames) into the library's queue; returns at once
        if on_device:
            det.submit_batch(dptr.value, K, D, FIDUCIAL_LEN, on_device=True, n_frames=nf, width=W, height=H)
        else:
            det.submit_batch(pinned, K, D, FIDUCIAL_LEN)  # H2D of this batch is queued here, inside the timed region

    outs = [None] * 4

    def finish(k):
        # results of the oldest batch in flight (host arrays) + its map update
        outs[k & 3] = det.collect_batch(outs[k & 3])
        counts, ids, corners, tfs = outs[k & 3]
        launches[0] += det.last_counters()["kernel_launches"]
        # fiducial_slam: the frames of this step are one camera stream -> one message per frame.  The
        # sequential fold is enqueued asynchronously so that it overlaps the detection of the next step
        # (the timed region ends with slam.sync()).  With N > 1 the per-rank map tables (as of the
        # previous step's fold, which finished long ago) are first exchanged with ONE NCCL all-gather and
        # merged identically on every rank.
        if dist is not None:
            tables = allgather_tables(slam.export_table(0), dist, device="cuda")
            slam.merge_tables(tables.reshape(-1), world, instance=0)
            launches[0] += 2
        slam.update_frames(counts, tfs, ident, ident, asynchronous=True)
        launches[0] += 1
        return counts

    def run_steps(on_device, steps):
        # software pipeline over steps: batch k+1 is submitted before batch k is collected, so the
        # latency-bound tail of one batch (grouping, identification, pose, D2H) runs under the threshold /
        # border-walk stages of the next.  Exactly `steps` batches are submitted AND collected in here.
        total = 0
        ahead = min(DEPTH - 1, steps)
        for _ in range(ahead):
            submit(on_device)
        for k in range(steps):
            if k + ahead < steps:
                submit(on_device)
            total += int(finish(k).sum())
        return total

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(on_device, steps):
        barrier()
        launches[0] = 0
        _lib.check(lib.fid_timer_start(det.h))
        t0 = time.perf_counter()
        total = run_steps(on_device, steps)
        slam.sync()
        ms = C.c_float(0)
        _lib.check(lib.fid_timer_stop(det.h, C.byref(ms)))
        wall = time.perf_counter() - t0
        barrier()
        t = torch.tensor([ms.value / 1e3, wall], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0]), float(t[1]), total, launches[0]

    run_steps(True, max(args.warmup, 3))
    run_steps(False, 2)
    slam.sync()

    sampler = ClockSampler(local_rank)
    sampler.start()
    dev_s, dev_wall, n_markers_found, n_launch = timed(True, args.steps)
    stage_ms = det.last_stage_ms()  # stages of the last batch call (nf frames)
    counters = det.last_counters()
    e2e_s, e2e_wall, _, _ = timed(False, args.steps)
    clocks = sampler.stop()

    # extra (not the metric): latency of ONE frame through the synchronous per-frame call the reference node makes
    # (imageCallback + poseEstimateCallback), host frame in, host results out, nothing else on the GPU
    single_ms = None
    if rank == 0:
        try:
            det1 = Detector(default_params(dictionary=dict_id), local_rank, W, H, 1)
            one = np.ascontiguousarray(pinned[:1])
            for _ in range(3):
                det1.detect_pose_batch(one, K, D, FIDUCIAL_LEN)
            t0 = time.perf_counter()
            for i in range(20):
                det1.detect_pose_batch(np.ascontiguousarray(pinned[i % nf : i % nf + 1]), K, D, FIDUCIAL_LEN)
            single_ms = (time.perf_counter() - t0) / 20 * 1e3
            det1.close()
        except Exception as e:  # pragma: no cover
            print("single-frame latency probe failed: %r" % (e,), file=sys.stderr)

    frames_total = nf * args.steps * world
    value = frames_total / dev_s
    e2e_value = frames_total / e2e_s

    if rank == 0:
        n_scales = 13
        peak, peak_src = measured_hbm_peak()
        bytes_per_frame = 3 * W * H + n_scales * W * H / 8.0  # BGR in, 13 bit planes out
        # (1) the threshold stage timed ALONE on one chunk (what a launch costs; compared with the burst copy peak)
        chunk = min(SLOT_FRAMES, nf)
        alone = C.c_float(0)
        _lib.check(lib.fid_debug_time_threshold(det.h, chunk, dptr, W, H, W * 3, W * 3 * H, 5, C.byref(alone)), "fid_debug_time_threshold")
        algo_bytes = bytes_per_frame * chunk
        achieved = algo_bytes / (alone.value / 1e3) / 1e9
        # (2) the same stage inside the pipelined step (CUDA events on its stream, while up to three other
        #     chunks run their own stages on the same SMs): share of the step
        n_launch_thr = (nf + SLOT_FRAMES - 1) // SLOT_FRAMES
        thr_pipe_ms = stage_ms["threshold"] / n_launch_thr
        total_stage = sum(v for k, v in stage_ms.items() if k not in ("h2d", "d2h") and not k.startswith("walk_r"))
        traffic = None
        try:  # DRAM bytes of the stage from the committed ncu --set full capture (per frame, scaled to this launch)
            traffic = float(json.load(open(os.path.join(ROOT, "profiles", "r01_threshold_traffic.json")))["dram_bytes_per_frame"]) * chunk
        except Exception:
            pass
        roofline = {
            "bound": "hbm",
            "kernel": "k_gray + k_threshold (threshold stage; the start-crack queues it also writes are not counted)",
            "achieved": achieved,
            "peak": peak,
            "unit": "GB/s",
            "frac": achieved / peak,
            "traffic": traffic,
            "peak_source": peak_src + "; kernel timed alone -> burst figure",
            "algorithmic_bytes_per_launch": algo_bytes,
            "frames_per_launch": chunk,
            "launch_ms": alone.value,
            "in_pipeline": {
                "launch_ms": thr_pipe_ms,
                "achieved": bytes_per_frame * SLOT_FRAMES / (thr_pipe_ms / 1e3) / 1e9 if thr_pipe_ms > 0 else None,
                "share_of_step": stage_ms["threshold"] / total_stage if total_stage else None,
                "note": "event-bracketed on the launching stream while other chunks' kernels share the SMs",
            },
            "stage_ms_per_batch": stage_ms,
            "work_per_batch": counters,
        }
        if os.environ.get("FID_BENCH_SKIP_CPU"):  # profiling runs (ncu) only
            cpu = {"value": None, "unit": "frames/s", "cores": 0, "kind": "reference", "sample": "skipped (FID_BENCH_SKIP_CPU set)"}
        else:
            cpu = cpu_reference_fps(frames[:8], dict_id, K, D, budget_s=16.0)
        d2h = nf * (4 + MAXM * 4 + MAXM * 32 + MAXM * C.sizeof(_lib.fid_transform))
        out = {
            "metric": "frames/sec 1920x1080 (detect+pose)",
            "value": value,
            "unit": "frames/s",
            "n_gpus": world,
            "steps": args.steps,
            "warmup": max(args.warmup, 3),
            "ms_per_step": dev_s * 1e3 / args.steps,
            "higher_is_better": True,
            "scaling": "weak",
            "vs_baseline": None,
            "dtype": "u8/i32 (threshold, contours), f32/f64 (sub-pixel, pose, map)",
            "data": "synthetic",
            "config": {
                "workload": "C2: 1920x1080 BGR8 stream, 16 markers/frame, DICT_6X6_250, detect+pose+map update, %d distinct frames per step per GPU" % nf,
                "frames_per_step_per_gpu": nf,
                "l2": "inputs larger than L2 (%d MB of distinct frames per step vs 126 MB L2)" % (frames.nbytes // 2**20),
                "parallelism": "one camera stream per GPU, no collective on detect/pose; map tables all-gathered (NCCL) and merged once per step" if world > 1 else "1 GPU",
                "markers_found_per_step": n_markers_found // max(1, args.steps),
            },
            "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": int(frames.nbytes), "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_s * 1e3 / args.steps},
            "gpu_launches": n_launch,
            "clocks": clocks,
            "roofline": roofline,
            "cpu_baseline": cpu,
            "wallclock_s": {"device_resident": dev_wall, "e2e": e2e_wall},
            "single_frame_latency_ms": single_ms,
        }
        print(json.dumps(out))
    lib.fid_device_free(det.h, dptr)
    lib.fid_host_free(hptr)
    det.close()
    slam.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
