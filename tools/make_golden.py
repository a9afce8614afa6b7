"""Golden vectors from the REFERENCE'S OWN compiled host code (oracle/_ref/libref_*_host.so = /root/reference/yolov8/src/
postprocess.cpp, yolov5/src/postprocess.cpp, retinaface/common.hpp compiled where they lie, oracle/Makefile): inputs and the
reference's outputs for nms() / nms_obb() / get_rect(), committed as tests/golden/ref_host.npz so that the oracle (and the
library's host functions) stay pinned on a machine where neither /root/reference nor oracle/_ref exists.
Run where /root/reference is mounted:  make -C oracle && python tools/make_golden.py
tests/test_golden_cpu.py consumes the file.  Inputs are stored in compact form (the meaningful floats of each plugin row)."""
import ctypes as C
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from oracle import oracle as O  # noqa: E402  (only to BUILD inputs: decoded plugin rows of the synthetic heads)
from tensorrtx_b200 import synth  # noqa: E402

REF = ROOT / "oracle" / "_ref"


def ref_nms(fn, buf, F, *thr):
    out = np.zeros((buf.shape[0] // F + 1) * F, np.float32)
    b = np.ascontiguousarray(buf, np.float32)
    n = fn(b.ctypes.data_as(C.c_void_p), *[C.c_float(t) for t in thr], out.ctypes.data_as(C.c_void_p))
    return out[:n * F].reshape(n, F).copy()


def compact(buf, F, cols):
    n = int(buf[0])
    return buf[1:1 + n * F].reshape(n, F)[:, cols].copy()


def main():
    v8 = C.CDLL(str(REF / "libref_yolov8_host.so"))
    v5 = C.CDLL(str(REF / "libref_yolov5_host.so"))
    rt = C.CDLL(str(REF / "libref_retina_host.so"))
    g = {}
    # ---- yolov8 nms(): 90-float rows, ltrb IoU, tie-break on bbox[0] (yolov8/src/postprocess.cpp:71-121) ----
    cols6 = [0, 1, 2, 3, 4, 5]
    for i, seed in enumerate((100, 101, 102)):
        out, _ = O.yolov8_decode(synth.yolov8_heads(1, seed=seed, n_obj=14))
        g[f"v8_{i}_in"] = compact(out[0], 90, cols6)
        g[f"v8_{i}_out"] = ref_nms(v8.ref_v8_nms, out[0], 90, 0.5, 0.45)[:, cols6]
    rng = np.random.default_rng(1)           # many equal confidences: the bbox[0] tie-break decides the order
    n = 300
    buf = np.zeros(1 + 1000 * 90, np.float32)
    rows = buf[1:1 + n * 90].reshape(n, 90)
    xy, wh = rng.uniform(0, 600, (n, 2)), rng.uniform(10, 60, (n, 2))
    rows[:, :2], rows[:, 2:4] = xy, xy + wh
    rows[:, 4] = np.round(rng.uniform(0.5, 1.0, n), 2)
    rows[:, 5] = rng.integers(0, 3, n)
    buf[0] = n
    g["v8_ties_in"] = compact(buf, 90, cols6)
    g["v8_ties_out"] = ref_nms(v8.ref_v8_nms, buf, 90, 0.5, 0.45)[:, cols6]
    # ---- nms_obb() + probiou (postprocess.cpp:303-393): cx,cy,w,h,conf,cls + angle (row float 89) ----
    cols_obb = [0, 1, 2, 3, 4, 5, 89]
    out = O.yolov8_decode(synth.yolov8_heads(1, seed=130, nc=15, extra=1, n_obj=24), nc=15, is_obb=True)[0]
    g["obb_in"] = compact(out[0], 90, cols_obb)
    for thr in (0.5, 0.2):
        g[f"obb_out_{int(thr * 10)}"] = ref_nms(v8.ref_v8_nms_obb, out[0], 90, 0.3, thr)[:, cols_obb]
    # ---- yolov5 nms(): 38-float rows, cxcywh IoU (yolov5/src/postprocess.cpp:30-80) ----
    out, _ = O.yolov5_decode(synth.yolov5_heads(1, seed=200, n_obj=40), synth.V5_ANCHORS)
    g["v5_in"] = compact(out[0], 38, cols6)
    g["v5_out"] = ref_nms(v5.ref_v5_nms, out[0], 38, 0.5, 0.45)[:, cols6]
    # ---- retinaface nms(): 15-float rows, conf > 0.1 (double literal), +1e-6 IoU (retinaface/common.hpp:91-130) ----
    out, _ = O.retina_decode(synth.retina_heads(1, seed=300, in_h=480, in_w=640, n_obj=30), in_h=480, in_w=640)
    g["retina_in"] = compact(out[0], 15, list(range(15)))
    g["retina_out"] = ref_nms(rt.ref_retina_nms, out[0], 15, 0.4)
    # ---- get_rect (postprocess.cpp:6-36 / yolov5 :4-29) ----
    rng = np.random.default_rng(50)
    sizes = ((1920, 1080), (1080, 1920), (640, 640), (333, 777), (641, 640), (50, 60))
    for variant, lib, fn in ((0, v8, "ref_v8_get_rect"), (1, v5, "ref_v5_get_rect")):
        boxes, rects, wh_ = [], [], []
        for (w, h) in sizes:
            for _ in range(60):
                if variant == 0:
                    x1, y1 = rng.uniform(-30, 650, 2)
                    bb = np.array([x1, y1, x1 + rng.uniform(-5, 400), y1 + rng.uniform(-5, 400)], np.float32)
                else:
                    bb = np.array([rng.uniform(-30, 670), rng.uniform(-30, 670), rng.uniform(0, 500), rng.uniform(0, 500)], np.float32)
                r = np.zeros(4, np.int32)
                getattr(lib, fn)(w, h, bb.ctypes.data_as(C.c_void_p), r.ctypes.data_as(C.c_void_p))
                boxes.append(bb), rects.append(r), wh_.append((w, h))
        g[f"rect{variant}_boxes"], g[f"rect{variant}_rects"], g[f"rect{variant}_sizes"] = np.array(boxes), np.array(rects), np.array(wh_, np.int32)
    # ---- RetinaFace get_rect_adapt_landmark (retinaface/common.hpp:65-89) and yolov8 get_rect_adapt_landmark (postprocess.cpp:38-69) ----
    rng = np.random.default_rng(73)
    boxes, lm_in, lm_out, rects, meta = [], [], [], [], []
    for (in_w, in_h) in ((640, 640), (640, 480)):
        for (w, h) in sizes:
            for _ in range(25):
                x1, y1 = rng.uniform(-30, in_w + 10), rng.uniform(-30, in_h + 10)
                bb = np.array([x1, y1, x1 + rng.uniform(-5, 400), y1 + rng.uniform(-5, 400)], np.float32)
                lmk = rng.uniform(-20, 660, 10).astype(np.float32)
                lo, r = lmk.copy(), np.zeros(4, np.int32)
                rt.ref_retina_get_rect_adapt_landmark(w, h, in_w, in_h, bb.copy().ctypes.data_as(C.c_void_p), lo.ctypes.data_as(C.c_void_p),
                                                      r.ctypes.data_as(C.c_void_p))
                boxes.append(bb), lm_in.append(lmk), lm_out.append(lo), rects.append(r), meta.append((w, h, in_w, in_h))
    g["retina_lmk_boxes"], g["retina_lmk_in"], g["retina_lmk_out"] = np.array(boxes), np.array(lm_in), np.array(lm_out)
    g["retina_lmk_rects"], g["retina_lmk_meta"] = np.array(rects), np.array(meta, np.int32)
    boxes, lm_in, lm_out, rects, meta = [], [], [], [], []
    for (w, h) in sizes:
        for _ in range(40):
            x1, y1 = rng.uniform(-30, 650, 2)
            bb = np.array([x1, y1, x1 + rng.uniform(-5, 400), y1 + rng.uniform(-5, 400)], np.float32)
            lmk = rng.uniform(-20, 660, 51).astype(np.float32)
            lo, r = lmk.copy(), np.zeros(4, np.int32)
            v8.ref_v8_get_rect_adapt_landmark(w, h, bb.copy().ctypes.data_as(C.c_void_p), lo.ctypes.data_as(C.c_void_p), r.ctypes.data_as(C.c_void_p))
            boxes.append(bb), lm_in.append(lmk), lm_out.append(lo), rects.append(r), meta.append((w, h))
    g["v8_lmk_boxes"], g["v8_lmk_in"], g["v8_lmk_out"] = np.array(boxes), np.array(lm_in), np.array(lm_out)
    g["v8_lmk_rects"], g["v8_lmk_meta"] = np.array(rects), np.array(meta, np.int32)
    # ---- process_decode_ptr_host / _obb (postprocess.cpp:131-147, 273-290) ----
    K = 120
    for elem, fn, F, key in ((7, v8.ref_v8_process_decode_ptr_host, 6, "pdh"), (8, v8.ref_v8_process_decode_ptr_host_obb, 7, "pdh_obb")):
        buf = np.zeros(1 + K * elem, np.float32)
        buf[0] = K
        rows = buf[1:].reshape(K, elem)
        rows[:, :] = rng.uniform(-3, 640, (K, elem)).astype(np.float32)
        rows[:, 6] = rng.integers(0, 3, K)
        out = np.zeros((K, F), np.float32)
        n = fn(buf.ctypes.data_as(C.c_void_p), elem, K, out.ctypes.data_as(C.c_void_p))
        g[key + "_in"], g[key + "_out"] = buf, out[:n].copy()
    dst = ROOT / "tests" / "golden" / "ref_host.npz"
    np.savez_compressed(dst, **g)
    print(dst, dst.stat().st_size, "bytes;", {k: v.shape for k, v in g.items()})


if __name__ == "__main__":
    main()
