"""SURVEY 8f rank 4, last item: the INT8 calibrator's host pre-process.  trtx_calib_letterbox_host (tensorrtx_b200/csrc/calib_host.cu)
against the calls Int8EntropyCalibrator2::getBatch makes (yolov8/src/calibrator.cpp:33-52): preprocess_img (yolov8/include/utils.h:
6-26) restated with the SAME OpenCV functions through cv2 4.13 -- cv2.resize(INTER_LINEAR) into the letterbox rectangle, 128-grey
canvas -- then cv2.dnn.blobFromImages(1 / 255.0, swapRB=True).  The arithmetic lives in OpenCV, a dependency that is not under
/root/reference: the library restates it and this test pins it bit for bit (with OpenCV's IPP accelerator on and off)."""
import ctypes as C

import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")


def _reference_batch(img, net_w, net_h):
    """preprocess_img + blobFromImages, line by line (float / int conversions as in utils.h:7-20)."""
    rows, cols = img.shape[:2]
    r_w = np.float32(net_w / (cols * 1.0))
    r_h = np.float32(net_h / (rows * 1.0))
    if r_h > r_w:
        w, h = net_w, int(np.float32(r_w * np.float32(rows)))
        x, y = 0, (net_h - h) // 2
    else:
        w, h = int(np.float32(r_h * np.float32(cols))), net_h
        x, y = (net_w - w) // 2, 0
    re = cv2.resize(img, (w, h), interpolation=cv2.INTER_LINEAR)
    out = np.full((net_h, net_w, 3), 128, np.uint8)
    out[y:y + h, x:x + w] = re
    blob = cv2.dnn.blobFromImages([out], 1.0 / 255.0, (net_w, net_h), (0, 0, 0), True, False)
    return blob[0], (x, y, w, h)


SIZES = [(1920, 1080), (1080, 1920), (640, 640), (1280, 720), (333, 777), (4000, 3000), (641, 640), (50, 60), (320, 240), (1279, 721),
         (2560, 1440), (7, 5), (640, 480), (639, 1)]


@pytest.mark.parametrize("ipp", [True, False])
def test_calibrator_letterbox_equals_opencv(ipp):
    from tensorrtx_b200 import _lib as L

    lib = L.load()
    rng = np.random.default_rng(11)
    old = cv2.ipp.useIPP()
    cv2.ipp.setUseIPP(ipp)
    try:
        for net_w, net_h in ((640, 640), (608, 352)):
            for (w, h) in SIZES:
                img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
                pad = np.zeros((h, w * 3 + 5), np.uint8)                  # a pitch that is not w * 3
                pad[:, :w * 3] = img.reshape(h, w * 3)
                rect = (C.c_int * 4)()
                assert lib.trtx_calib_letterbox_rect(w, h, net_w, net_h, rect) == 0
                try:
                    ref, rref = _reference_batch(img, net_w, net_h)
                except cv2.error:                                         # an empty letterbox: cv::resize throws, we refuse
                    out = np.zeros((3, net_h, net_w), np.float32)
                    assert lib.trtx_calib_letterbox_host(pad.ctypes.data_as(C.c_void_p), w, h, pad.strides[0], net_w, net_h,
                                                         out.ctypes.data_as(C.c_void_p)) == L.ERR_UNSUPPORTED
                    continue
                assert tuple(rect) == rref
                out = np.full((3, net_h, net_w), -1.0, np.float32)
                assert lib.trtx_calib_letterbox_host(pad.ctypes.data_as(C.c_void_p), w, h, pad.strides[0], net_w, net_h,
                                                     out.ctypes.data_as(C.c_void_p)) == 0
                assert np.array_equal(out, ref), (w, h, net_w, net_h, np.abs(out - ref).max())
    finally:
        cv2.ipp.setUseIPP(old)


def test_calibrator_letterbox_rejects_bad_arguments():
    from tensorrtx_b200 import _lib as L

    lib = L.load()
    img = np.zeros((4, 4, 3), np.uint8)
    out = np.zeros((3, 8, 8), np.float32)
    assert lib.trtx_calib_letterbox_host(None, 4, 4, 12, 8, 8, out.ctypes.data_as(C.c_void_p)) == L.ERR_INVALID
    assert lib.trtx_calib_letterbox_host(img.ctypes.data_as(C.c_void_p), 4, 4, 11, 8, 8, out.ctypes.data_as(C.c_void_p)) == L.ERR_INVALID
    assert lib.trtx_calib_letterbox_host(img.ctypes.data_as(C.c_void_p), 0, 4, 12, 8, 8, out.ctypes.data_as(C.c_void_p)) == L.ERR_INVALID


def test_calibrator_batcher_mirror():
    from tensorrtx_b200 import plugins as P

    rng = np.random.default_rng(12)
    imgs = [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for (w, h) in ((800, 600), (333, 500), (640, 640))]
    cal = P.Int8CalibratorBatcher(3, 640, 640)
    batch = cal.get_batch(imgs)
    assert batch.shape == (3, 3, 640, 640) and batch.size == cal.input_count
    for i, img in enumerate(imgs):
        assert np.array_equal(batch[i], _reference_batch(img, 640, 640)[0])
