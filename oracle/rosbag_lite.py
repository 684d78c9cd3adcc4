"""Minimal ROS bag v2.0 reader (uncompressed chunks only).  TEST INFRASTRUCTURE ONLY.

Used once, by tests/golden/make_golden.py, to lift the reference's two bag fixtures
(fiducial_slam/test/aruco_transforms.bag, aruco_images.bag) into committed .npz vectors.
Wire format as documented in SURVEY.md Appendix D.
"""
from __future__ import annotations

import struct


def _records(buf: bytes, pos: int, end: int):
    while pos < end:
        (hlen,) = struct.unpack_from("<I", buf, pos)
        pos += 4
        hdr = {}
        hend = pos + hlen
        while pos < hend:
            (flen,) = struct.unpack_from("<I", buf, pos)
            pos += 4
            fieldb = buf[pos : pos + flen]
            pos += flen
            k, _, v = fieldb.partition(b"=")
            hdr[k.decode()] = v
        (dlen,) = struct.unpack_from("<I", buf, pos)
        pos += 4
        yield hdr, buf[pos : pos + dlen]
        pos += dlen


def read_bag(path: str):
    """Returns list of (topic, msg_type, raw_message_bytes) in file order."""
    buf = open(path, "rb").read()
    magic = b"#ROSBAG V2.0\n"
    assert buf.startswith(magic), "not a v2.0 bag"
    conns: dict[int, tuple[str, str]] = {}
    msgs = []

    def walk(pos, end):
        for hdr, data in _records(buf_local[0], pos, end):
            op = hdr["op"][0]
            if op == 0x05:  # chunk
                assert hdr["compression"] == b"none", "compressed chunks unsupported"
                saved = buf_local[0]
                buf_local[0] = data
                walk(0, len(data))
                buf_local[0] = saved
            elif op == 0x07:  # connection
                (cid,) = struct.unpack("<I", hdr["conn"])
                chdr = {}
                p = 0
                while p < len(data):
                    (flen,) = struct.unpack_from("<I", data, p)
                    p += 4
                    k, _, v = data[p : p + flen].partition(b"=")
                    chdr[k.decode()] = v
                    p += flen
                conns[cid] = (hdr["topic"].decode(), chdr["type"].decode())
            elif op == 0x02:  # message
                (cid,) = struct.unpack("<I", hdr["conn"])
                msgs.append((cid, data))

    buf_local = [buf]
    walk(len(magic), len(buf))
    return [(conns[c][0], conns[c][1], d) for c, d in msgs]


class _Cur:
    def __init__(self, b):
        self.b = b
        self.p = 0

    def take(self, fmt):
        v = struct.unpack_from("<" + fmt, self.b, self.p)
        self.p += struct.calcsize("<" + fmt)
        return v if len(v) > 1 else v[0]

    def string(self):
        n = self.take("I")
        s = self.b[self.p : self.p + n]
        self.p += n
        return s


def parse_header(c: _Cur):
    seq = c.take("I")
    sec, nsec = c.take("II")
    frame = c.string().decode()
    return dict(seq=seq, stamp=(sec, nsec), frame_id=frame)


def parse_fiducial_transform_array(raw: bytes):
    """fiducial_msgs/FiducialTransformArray (msg/FiducialTransformArray.msg:3-5)."""
    c = _Cur(raw)
    hdr = parse_header(c)
    image_seq = c.take("i")
    n = c.take("I")
    tfs = []
    for _ in range(n):
        fid = c.take("i")
        t = c.take("ddd")
        q = c.take("dddd")
        ie, oe, area = c.take("ddd")
        tfs.append(dict(fiducial_id=fid, translation=list(t), rotation=list(q), image_error=ie, object_error=oe, fiducial_area=area))
    return dict(header=hdr, image_seq=image_seq, transforms=tfs)


def parse_camera_info(raw: bytes):
    c = _Cur(raw)
    hdr = parse_header(c)
    h, w = c.take("II")
    model = c.string().decode()
    nd = c.take("I")
    D = [c.take("d") for _ in range(nd)]
    K = [c.take("d") for _ in range(9)]
    return dict(header=hdr, height=h, width=w, model=model, D=D, K=K)


def parse_compressed_image(raw: bytes):
    c = _Cur(raw)
    hdr = parse_header(c)
    fmt = c.string().decode()
    data = c.string()
    return dict(header=hdr, format=fmt, data=data)
