"""Golden vectors of the reference's own compiled host code (tests/golden/ref_host.npz, written by tools/make_golden.py where
/root/reference is mounted): the oracle's nms() variants, nms_obb and get_rect -- and the library's host get_rect -- reproduce
them bit for bit on any machine, also where neither /root/reference nor oracle/_ref exists."""
from pathlib import Path

import numpy as np
import pytest

GOLD = Path(__file__).resolve().parent / "golden" / "ref_host.npz"


@pytest.fixture(scope="module")
def gold():
    assert GOLD.exists(), "tests/golden/ref_host.npz is committed with the repository"
    return np.load(GOLD)


def _plugin_buf(rows, F, cols, max_rows):
    buf = np.zeros(1 + max_rows * F, np.float32)
    n = rows.shape[0]
    buf[0] = n
    buf[1:1 + n * F].reshape(n, F)[:, cols] = rows
    return buf


COLS6 = [0, 1, 2, 3, 4, 5]


@pytest.mark.parametrize("name", ["v8_0", "v8_1", "v8_2", "v8_ties"])
def test_v8_nms_equals_golden(oracle, gold, name):
    buf = _plugin_buf(gold[name + "_in"], 90, COLS6, 1000)
    mine, _ = oracle.nms(0, buf, 1000, 90, 0.5, 0.45)
    assert len(mine) > 10 and np.array_equal(mine[:, COLS6], gold[name + "_out"])   # same rows, same order, same bits
    assert not mine[:, 6:].any()


@pytest.mark.parametrize("thr", [0.5, 0.2])
def test_v8_nms_obb_equals_golden(oracle, gold, thr):
    cols = COLS6 + [89]
    buf = _plugin_buf(gold["obb_in"], 90, cols, 1000)
    mine, _ = oracle.nms(3, buf, 1000, 90, 0.3, thr)
    assert np.array_equal(mine[:, cols], gold[f"obb_out_{int(thr * 10)}"])


def test_v5_nms_equals_golden(oracle, gold):
    buf = _plugin_buf(gold["v5_in"], 38, COLS6, 1000)
    mine, _ = oracle.nms(1, buf, 1000, 38, 0.5, 0.45)
    assert np.array_equal(mine[:, COLS6], gold["v5_out"])


def test_retina_nms_equals_golden(oracle, gold):
    tp = oracle.retina_total_priors(480, 640)
    buf = _plugin_buf(gold["retina_in"], 15, list(range(15)), tp)
    mine, _ = oracle.nms(2, buf, tp, 15, 0.1, 0.4)
    assert np.array_equal(mine, gold["retina_out"])


@pytest.mark.parametrize("variant", [0, 1])
def test_get_rect_equals_golden(oracle, gold, variant):
    from tensorrtx_b200 import plugins as P
    boxes, rects, sizes = gold[f"rect{variant}_boxes"], gold[f"rect{variant}_rects"], gold[f"rect{variant}_sizes"]
    for bb, r, (w, h) in zip(boxes, rects, sizes):
        assert np.array_equal(oracle.get_rect(variant, int(w), int(h), bb), r)
        assert P.get_rect(int(w), int(h), bb, variant=variant) == tuple(int(v) for v in r)     # the library's host function


def test_landmark_and_decode_ptr_host_functions_equal_golden(gold):
    """get_rect_adapt_landmark (yolov8 and RetinaFace flavours) and process_decode_ptr_host(_obb): the library's host functions."""
    import ctypes as C

    from tensorrtx_b200 import _lib as L
    from tensorrtx_b200 import plugins as P
    lib = L.load()
    for bb, li, lo, r, (w, h, in_w, in_h) in zip(gold["retina_lmk_boxes"], gold["retina_lmk_in"], gold["retina_lmk_out"], gold["retina_lmk_rects"],
                                                 gold["retina_lmk_meta"]):
        lm, rect = li.copy(), (C.c_int * 4)()
        assert lib.trtx_retina_get_rect_adapt_landmark(int(in_w), int(in_h), int(w), int(h), np.ascontiguousarray(bb).ctypes.data_as(C.POINTER(C.c_float)),
                                                       lm.ctypes.data_as(C.POINTER(C.c_float)), rect) == 0
        assert list(rect) == r.tolist() and np.array_equal(lm, lo)
    for bb, li, lo, r, (w, h) in zip(gold["v8_lmk_boxes"], gold["v8_lmk_in"], gold["v8_lmk_out"], gold["v8_lmk_rects"], gold["v8_lmk_meta"]):
        rect, mapped = P.get_rect_adapt_landmark(int(w), int(h), bb, li)
        assert rect == tuple(int(v) for v in r) and np.array_equal(np.asarray(mapped, np.float32), lo)
    for key, elem, F, fn in (("pdh", 7, 6, lib.trtx_process_decode_ptr_host), ("pdh_obb", 8, 7, lib.trtx_process_decode_ptr_host_obb)):
        buf, ref = np.ascontiguousarray(gold[key + "_in"]), gold[key + "_out"]
        K = int(buf[0])
        out = np.zeros((K, F), np.float32)
        n = fn(buf.ctypes.data_as(C.POINTER(C.c_float)), elem, K, out.ctypes.data_as(C.POINTER(C.c_float)))
        assert n == len(ref) and n > 10 and np.array_equal(out[:n], ref)
