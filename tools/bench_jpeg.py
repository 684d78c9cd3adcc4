#!/usr/bin/env python
"""JPEG ingest on workload C2: the frames of bench.py's stream as JPEG (cv2.imencode, quality 90, 4:2:0 -- what
compressed_image_transport publishes by default is quality 80), decoded by fid_jpeg_decode_batch straight into HBM and fed to
the detector with bgr_on_device = 1.  Reports, per GPU: end-to-end frames/s from JPEG bytes in host memory to results on the
host, the bytes that cross PCIe per frame, host entropy-decoding and device times, and cv2.imdecode on the host cores beside it
(1 thread = what the reference node's transport plugin does, and one process per core)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def imdecode_worker(args):
    import cv2

    cv2.setNumThreads(1)
    streams, reps = args
    t0 = time.perf_counter()
    for _ in range(reps):
        for s in streams:
            cv2.imdecode(s, cv2.IMREAD_COLOR)
    return time.perf_counter() - t0


def main():
    import cv2
    import torch

    from fiducials_b200 import synth
    from fiducials_b200.node import Detector, JpegDecoder, default_params

    quality = int(os.environ.get("JPEG_QUALITY", "90"))
    nb, steps, warm = 128, int(os.environ.get("STEPS", "10")), 3
    frames = synth.make_config_stream("C2", nb, seed=0, realizations=8)[0]
    K, D = synth.camera_for(1920, 1080)
    streams = [cv2.imencode(".jpg", f, [cv2.IMWRITE_JPEG_QUALITY, quality])[1] for f in frames]
    jpeg_bytes = float(np.mean([s.size for s in streams]))
    threads = int(os.environ.get("JPEG_THREADS", "0"))
    dec = JpegDecoder(1920, 1080, nb, n_threads=threads)
    det = Detector(default_params(dictionary=10), 0, 1920, 1080, 64)
    bufs = [torch.zeros(frames.shape, dtype=torch.uint8, device="cuda:0") for _ in range(2)]
    # parity of the whole path on this batch: JPEG -> device == imdecode -> upload
    dec.decode(streams, bufs[0].data_ptr(), 1920, 1080)
    ref = np.stack([cv2.imdecode(s, cv2.IMREAD_COLOR) for s in streams[:8]])
    assert np.array_equal(bufs[0][:8].cpu().numpy(), ref), "device JPEG decode differs from cv2.imdecode"
    host_ms, dev_ms, h2d = [], [], []

    def run(n):
        found = 0
        for k in range(n):
            dec.decode(streams, bufs[k & 1].data_ptr(), 1920, 1080, sync=True)
            st = dec.stats()
            host_ms.append(st["host_decode_ms"]); dev_ms.append(st["device_ms"]); h2d.append(st["h2d_bytes"])
            det.submit_batch(bufs[k & 1].data_ptr(), K, D, 0.14, on_device=True, n_frames=nb, width=1920, height=1080)
            if k >= 1:
                found += int(det.collect_batch()[0].sum())
        found += int(det.collect_batch()[0].sum())
        return found

    run(warm)
    del host_ms[:], dev_ms[:], h2d[:]
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    found = run(steps)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    # cv2.imdecode beside it
    t1 = imdecode_worker((streams[:32], 1))
    import multiprocessing as mp

    ncpu = len(os.sched_getaffinity(0))
    with mp.get_context("spawn").Pool(ncpu) as pool:
        pool.map(imdecode_worker, [(streams[:2], 1)] * ncpu)
        t0 = time.perf_counter()
        pool.map(imdecode_worker, [(streams[i % nb : i % nb + 8] if i % nb + 8 <= nb else streams[:8], 2) for i in range(ncpu)])
        tp = time.perf_counter() - t0
    out = {
        "workload": "C2 frames as JPEG (cv2.imencode quality %d, 4:2:0), %d frames per batch, decode -> detect + pose, 1 GPU" % (quality, nb),
        "e2e_frames_per_s": nb * steps / dt,
        "markers_found_per_batch": found / steps,
        "jpeg_bytes_per_frame": jpeg_bytes,
        "pcie_bytes_per_frame": float(np.mean(h2d)) / nb,
        "raw_bgr_bytes_per_frame": 1920 * 1080 * 3,
        "host_entropy_decode_ms_per_batch": float(np.mean(host_ms)),
        "host_entropy_decode_frames_per_s": nb / (float(np.mean(host_ms)) / 1e3),
        "host_threads": threads or min(os.cpu_count(), 64),
        "host_cpus": {"logical": os.cpu_count(), "affinity": ncpu, "cgroup_cpu_max": (open("/sys/fs/cgroup/cpu.max").read().strip() if os.path.exists("/sys/fs/cgroup/cpu.max") else None),
                      "loadavg": os.getloadavg()},
        "device_decode_ms_per_batch (copies + inverse DCT + colour, CUDA events)": float(np.mean(dev_ms)),
        "cv2_imdecode_frames_per_s_1_thread (the reference's transport plugin)": 32 / t1,
        "cv2_imdecode_frames_per_s_%d_processes" % ncpu: ncpu * 16 / tp,
        "parity": "device output == cv2.imdecode on the first 8 frames of the batch (bit-exact); tests/test_gpu_jpeg.py",
    }
    print(json.dumps(out))


if __name__ == "__main__":
    main()
