"""N > 1 host logic on CPU: world_size-2 gloo.  Each rank builds a map table from its half of a
message stream (with the numpy oracle -- there is no GPU here), the tables are exchanged with
fiducials_b200.multigpu.allgather_tables, and every rank must obtain the same merged map, equal to
the single-process merge.  The device merge kernel itself is checked on the GPU
(tests/test_gpu_slam.py::test_merge_matches_oracle_merge)."""
import math
import os
import socket
import struct

import numpy as np
import torch.multiprocessing as mp

RECORD = struct.Struct("<ii3d4dd")  # fid_map_record: id, num_obs, t[3], q[4], variance  (80 bytes)


def _messages(n_fid=10, n_msgs=40, per=5, seed=3):
    from oracle import slam_oracle as so

    rng = np.random.default_rng(seed)
    grid = [(float(i % 4), float(i // 4), 2.5) for i in range(n_fid)]
    msgs = []
    for k in range(n_msgs):
        cam = np.array([1.5 + math.sin(0.1 * k), 1.0 + math.cos(0.07 * k), 0.0])
        vis = sorted(range(n_fid), key=lambda i: (grid[i][0] - cam[0]) ** 2 + (grid[i][1] - cam[1]) ** 2)[:per]
        msgs.append([dict(fiducial_id=100 + i, translation=np.array(grid[i]) - cam + rng.normal(0, 0.004, 3),
                          rotation=np.array(so.q_from_rpy(math.pi, 0.0, math.pi)), image_error=0.1, object_error=float(rng.uniform(1e-4, 1e-2)), fiducial_area=900.0)
                     for i in vis])
    return msgs, grid


def _table(m, cap=32):
    from oracle import slam_oracle as so

    buf = bytearray()
    rows = sorted(m.fiducials.values(), key=lambda f: f.id)
    for f in rows:
        buf += RECORD.pack(f.id, f.numObs, *f.pose.t, *so.m_to_q(f.pose.R), f.pose.var)
    for _ in range(cap - len(rows)):
        buf += RECORD.pack(-1, 0, 0, 0, 0, 0, 0, 0, 1, 0)
    return np.frombuffer(bytes(buf), np.uint8)


def _untable(raw):
    from oracle import slam_oracle as so

    out = []
    for i in range(len(raw) // RECORD.size):
        fid, n, tx, ty, tz, qx, qy, qz, qw, var = RECORD.unpack_from(raw.tobytes(), i * RECORD.size)
        if fid >= 0:
            out.append((fid, so.TWV.from_qt([qx, qy, qz, qw], [tx, ty, tz], var), n))
    return out


def _rank_map(rank, world):
    from fiducials_b200.multigpu import shard_frames
    from oracle import slam_oracle as so

    msgs, grid = _messages()
    lo, hi = shard_frames(len(msgs), rank, world)
    m = so.Map()
    m.load_entry(100, grid[0][0], grid[0][1], grid[0][2], 180, 0, 180, 0, 0)
    ident = so.TWV.identity()
    for msg in msgs[lo:hi]:
        m.update(so.observations_from_transforms(msg), ident, ident)
    return m


def _worker(rank, world, port, q):
    import torch.distributed as dist

    from fiducials_b200.multigpu import allgather_tables
    from oracle import slam_oracle as so

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    tables = allgather_tables(_table(_rank_map(rank, world)), dist)
    merged = so.merge_maps([_untable(t) for t in tables])
    q.put((rank, {fid: (p.t, so.get_rpy(p.R), p.var, n) for fid, (p, n) in merged.items()}))
    dist.barrier()
    dist.destroy_process_group()


def test_allgather_merge_world2():
    from oracle import slam_oracle as so

    world = 2
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-process reference of the same merge
    ref = so.merge_maps([_untable(_table(_rank_map(r, world))) for r in range(world)])
    assert sorted(results[0]) == sorted(results[1]) == sorted(ref)
    for fid, (p, n) in ref.items():
        for r in range(world):
            t, rpy, var, nn = results[r][fid]
            assert np.abs(np.array(t) - np.array(p.t)).max() < 1e-12 and np.abs(np.array(rpy) - np.array(so.get_rpy(p.R))).max() < 1e-12
            assert nn == n and abs(var - p.var) < 1e-15


def test_shard_frames_covers_stream():
    from fiducials_b200.multigpu import shard_frames

    for n in (0, 1, 7, 64, 129):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                lo, hi = shard_frames(n, r, world)
                seen += list(range(lo, hi))
            assert seen == list(range(n))
