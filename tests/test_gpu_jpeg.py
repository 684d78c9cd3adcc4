"""JPEG ingest on the device through the C-ABI: fid_jpeg_decode_batch against cv2.imdecode -- the decoder
compressed_image_transport puts in front of the reference's imageCallback (aruco_detect.cpp:332,348; launch default
transport `compressed`, aruco_detect.launch:6,28).  Integer pipeline: bit-exact."""
import ctypes as C

import cv2
import numpy as np
import pytest

from fiducials_b200 import synth

pytestmark = pytest.mark.gpu

SAMPLING = {"444": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444, "422": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_422, "420": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_420}


def scene(h, w, seed):
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    img = np.stack([128 + 100 * np.sin(xx / 17.0 + seed), 128 + 100 * np.cos(yy / 11.0), 64 + (xx + yy) % 160], -1)
    img += rng.normal(0, 12, img.shape)
    img[h // 4 : h // 2, w // 4 : w // 2] = rng.integers(0, 2, (h // 2 - h // 4, w // 2 - w // 4, 1)) * 255
    return np.clip(img, 0, 255).astype(np.uint8)


def device_decode(streams, w, h, n_threads=0):
    import torch
    from fiducials_b200.node import JpegDecoder

    dec = JpegDecoder(w, h, len(streams), n_threads=n_threads)
    out = torch.zeros((len(streams), h, w, 3), dtype=torch.uint8, device="cuda:0")
    status = dec.decode(streams, out.data_ptr(), w, h)
    torch.cuda.synchronize()
    res = out.cpu().numpy()
    st = dec.stats()
    dec.close()
    return status, res, st


@pytest.mark.parametrize("sampling", ["420", "422", "444"])
@pytest.mark.parametrize("shape", [(96, 128), (97, 131), (241, 322), (17, 9)])
def test_batch_matches_imdecode(sampling, shape):
    h, w = shape
    streams, refs = [], []
    for i, q in enumerate([30, 75, 95, 100, 60]):  # a different quantisation table per image of the batch
        ok, buf = cv2.imencode(".jpg", scene(h, w, 10 * i + w), [cv2.IMWRITE_JPEG_QUALITY, q, cv2.IMWRITE_JPEG_SAMPLING_FACTOR, SAMPLING[sampling]])
        assert ok
        streams.append(buf)
        refs.append(cv2.imdecode(buf, cv2.IMREAD_COLOR))
    status, got, _ = device_decode(streams, w, h)
    assert status.tolist() == [0] * 5
    for i in range(5):
        assert np.array_equal(got[i], refs[i]), "image %d: max diff %d" % (i, np.abs(got[i].astype(int) - refs[i]).max())


def test_grey_restart_optimised_tables_single_thread():
    h, w = 120, 200
    img = scene(h, w, 3)
    grey = cv2.cvtColor(img, cv2.COLOR_BGR2GRAY)
    streams = [cv2.imencode(".jpg", grey, [cv2.IMWRITE_JPEG_QUALITY, 90, cv2.IMWRITE_JPEG_RST_INTERVAL, 3])[1]] * 3
    status, got, _ = device_decode(streams, w, h, n_threads=1)
    assert status.tolist() == [0, 0, 0]
    ref = cv2.imdecode(streams[0], cv2.IMREAD_COLOR)
    assert all(np.array_equal(got[i], ref) for i in range(3))
    s2 = [cv2.imencode(".jpg", img, [cv2.IMWRITE_JPEG_QUALITY, 92, cv2.IMWRITE_JPEG_OPTIMIZE, 1, cv2.IMWRITE_JPEG_RST_INTERVAL, 5])[1]]
    status, got, _ = device_decode(s2, w, h)
    assert status.tolist() == [0] and np.array_equal(got[0], cv2.imdecode(s2[0], cv2.IMREAD_COLOR))


def test_bad_images_are_reported_per_image_and_do_not_disturb_the_batch():
    h, w = 64, 96
    good = cv2.imencode(".jpg", scene(h, w, 1), [cv2.IMWRITE_JPEG_QUALITY, 80])[1]
    prog = cv2.imencode(".jpg", scene(h, w, 2), [cv2.IMWRITE_JPEG_PROGRESSIVE, 1])[1]
    other_size = cv2.imencode(".jpg", scene(h + 8, w, 3))[1]
    mixed = cv2.imencode(".jpg", scene(h, w, 4), [cv2.IMWRITE_JPEG_SAMPLING_FACTOR, SAMPLING["444"]])[1]
    status, got, _ = device_decode([good, prog, np.frombuffer(bytes(good)[:150], np.uint8), other_size, mixed, good], w, h)
    assert status[0] == 0 and status[5] == 0
    assert status[1] == -4 and status[2] < 0 and status[3] == -1 and status[4] == -4
    ref = cv2.imdecode(good, cv2.IMREAD_COLOR)
    assert np.array_equal(got[0], ref) and np.array_equal(got[5], ref)
    assert not got[1].any() and not got[3].any()  # skipped frames stay untouched


def test_jpeg_stream_into_the_detector_equals_imdecode_into_the_detector():
    """compressed transport end to end: JPEG -> device decode -> detect + pose, against cv2.imdecode -> the same detector."""
    import torch
    from fiducials_b200.node import Detector, JpegDecoder, default_params

    frames, meta = synth.make_config_stream("C1", 4, seed=2)[:2]
    K, D = synth.camera_for(640, 480)
    streams = [cv2.imencode(".jpg", f, [cv2.IMWRITE_JPEG_QUALITY, 90])[1] for f in frames]
    decoded = np.stack([cv2.imdecode(s, cv2.IMREAD_COLOR) for s in streams])
    det = Detector(default_params(dictionary=6), 0, 640, 480, 4)
    c1, i1, k1, t1 = det.detect_pose_batch(decoded, K, D, 0.14)
    c1, i1, k1 = c1.copy(), i1.copy(), k1.copy()
    assert c1.sum() > 0
    dec = JpegDecoder(640, 480, 4)
    dev = torch.zeros(decoded.shape, dtype=torch.uint8, device="cuda:0")
    assert dec.decode(streams, dev.data_ptr(), 640, 480).tolist() == [0] * 4
    c2, i2, k2, t2 = det.detect_pose_batch(dev.data_ptr(), K, D, 0.14, on_device=True, n_frames=4, width=640, height=480)
    assert np.array_equal(c1, c2) and np.array_equal(i1, i2) and np.array_equal(k1, k2)
    dec.close()
    det.close()


def test_full_hd_frames_and_transfer_size():
    frames = synth.make_config_stream("C2", 3, seed=1)[0]
    streams = [cv2.imencode(".jpg", f, [cv2.IMWRITE_JPEG_QUALITY, 90])[1] for f in frames]
    status, got, st = device_decode(streams, 1920, 1080)
    assert status.tolist() == [0, 0, 0]
    for i in range(3):
        assert np.array_equal(got[i], cv2.imdecode(streams[i], cv2.IMREAD_COLOR))
    assert st["h2d_bytes"] < frames.nbytes / 3  # the sparse coefficients are what crosses PCIe
