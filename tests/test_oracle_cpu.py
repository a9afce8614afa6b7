"""CPU-only checks of the oracle (test infrastructure) against independent implementations:
torchvision.ops.nms / batched_nms (greedy, plain IoU, `>` threshold), cv2.invertAffineTransform and
cv2.warpAffine-free closed forms.  These pin the oracle where the reference has no fixtures
(SURVEY.md 8c)."""
import numpy as np
import pytest
import torch

from tensorrtx_b200 import synth


def test_decode_count_and_box_formula(oracle):
    heads = synth.yolov8_heads(2, seed=0)
    out, idx = oracle.yolov8_decode(heads)
    # independent numpy restatement of yololayer.cu:193-220
    lv_off = 0
    exp = [[] for _ in range(2)]
    for h, s in zip(heads, (8, 16, 32)):
        gw = 640 // s
        x = h[:, 4:84].astype(np.float32)
        p = (1.0 / (1.0 + np.exp(-x.astype(np.float64)))).astype(np.float32)
        best = p.max(1)
        cls = p.argmax(1)  # first max
        for b in range(2):
            for e in np.where(best[b] >= np.float32(0.1))[0]:
                row, col = divmod(int(e), gw)
                d = h[b, :4, e]
                exp[b].append((lv_off + e, (col + 0.5 - d[0]) * s, (row + 0.5 - d[1]) * s, (col + 0.5 + d[2]) * s,
                               (row + 0.5 + d[3]) * s, best[b, e], cls[b, e]))
        lv_off += h.shape[2]
    for b in range(2):
        n = int(out[b, 0])
        assert n == len(exp[b])
        rows = out[b, 1:1 + n * 90].reshape(n, 90)
        e = np.asarray(exp[b], dtype=np.float64)
        assert np.array_equal(idx[b, :n], e[:, 0].astype(np.int32))
        np.testing.assert_allclose(rows[:, :4], e[:, 1:5], atol=1e-4)
        np.testing.assert_allclose(rows[:, 4], e[:, 5], atol=1e-6)
        assert np.array_equal(rows[:, 5], e[:, 6])


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_nms_v8_matches_torchvision_batched_nms(oracle, seed):
    import torchvision

    heads = synth.yolov8_heads(1, seed=seed)
    out, _ = oracle.yolov8_decode(heads)
    n = int(out[0, 0])
    rows = out[0, 1:1 + n * 90].reshape(n, 90)
    res, src = oracle.nms(0, out[0], 1000, 90, 0.5, 0.45)
    valid = np.where(rows[:, 4] > 0.5)[0]
    keep = torchvision.ops.batched_nms(torch.from_numpy(rows[valid, :4].copy()), torch.from_numpy(rows[valid, 4].copy()),
                                       torch.from_numpy(rows[valid, 5].astype(np.int64)), 0.45)
    assert sorted(valid[keep.numpy()].tolist()) == sorted(src.tolist())
    # order of `res`: class ascending, conf descending (std::map + std::sort)
    key = list(zip(res[:, 5], -res[:, 4]))
    assert key == sorted(key)


def test_nms_v5_cxcywh_matches_torchvision(oracle):
    import torchvision

    heads = synth.yolov5_heads(1, seed=3)
    out, _ = oracle.yolov5_decode(heads, synth.V5_ANCHORS)
    n = int(out[0, 0])
    rows = out[0, 1:1 + n * 38].reshape(n, 38)
    res, src = oracle.nms(1, out[0], 1000, 38, 0.5, 0.45)
    valid = np.where(rows[:, 4] > 0.5)[0]
    b = rows[valid, :4]
    xyxy = np.stack([b[:, 0] - b[:, 2] / 2, b[:, 1] - b[:, 3] / 2, b[:, 0] + b[:, 2] / 2, b[:, 1] + b[:, 3] / 2], 1)
    keep = torchvision.ops.batched_nms(torch.from_numpy(xyxy), torch.from_numpy(rows[valid, 4].copy()),
                                       torch.from_numpy(rows[valid, 5].astype(np.int64)), 0.45)
    assert sorted(valid[keep.numpy()].tolist()) == sorted(src.tolist())


def test_nms_properties_permutation_invariance(oracle):
    heads = synth.yolov8_heads(1, seed=4)
    out, _ = oracle.yolov8_decode(heads)
    n = int(out[0, 0])
    rows = out[0, 1:1 + n * 90].reshape(n, 90).copy()
    res, _ = oracle.nms(0, out[0], 1000, 90, 0.5, 0.45)
    perm = np.random.default_rng(0).permutation(n)
    out2 = out[0].copy()
    out2[1:1 + n * 90] = rows[perm].reshape(-1)
    res2, _ = oracle.nms(0, out2, 1000, 90, 0.5, 0.45)
    assert np.array_equal(res, res2)  # tie-free input -> identical kept rows in identical order


def _s2d(sw, sh, dw=640, dh=640):
    f = np.float32
    scale = f(min(f(dh) / f(sh), f(dw) / f(sw)))  # preprocess.cu:98
    m2 = f(float(f(-scale * f(sw))) * 0.5 + dw * 0.5)
    m5 = f(float(f(-scale * f(sh))) * 0.5 + dh * 0.5)
    return np.array([[scale, 0, m2], [0, scale, m5]], np.float32)


def test_letterbox_matrix_matches_opencv(oracle):
    """cv::invertAffineTransform is third-party arithmetic (OpenCV); pin the restatement to cv2 4.13."""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(0)
    sizes = [(640, 640), (1920, 1080), (1080, 1920), (333, 517), (3000, 2000)]
    sizes += [tuple(int(v) for v in rng.integers(50, 4000, 2)) for _ in range(500)]
    for (sw, sh) in sizes:
        d2s = cv2.invertAffineTransform(_s2d(sw, sh))
        m = oracle.letterbox_matrix(sw, sh, 640, 640)
        assert np.array_equal(m, d2s.reshape(-1).astype(np.float32)), (sw, sh, m, d2s)


def test_warpaffine_identity_scale_is_2x2_mean(oracle):
    # at scale 1 the reference's +0.5 offset makes every pixel the mean of a 2x2 neighbourhood
    img = synth.frames(1, seed=1, h=64, w=64)[0]
    dst = oracle.warpaffine(img, 64, 64)
    f = img.astype(np.float32)
    exp = (f[:-1, :-1] + f[:-1, 1:] + f[1:, :-1] + f[1:, 1:]) * 0.25 / 255.0
    np.testing.assert_allclose(dst[2, :-1, :-1], exp[:, :, 0], atol=1e-6)  # dst plane 2 = B (bgr->rgb swap)
    np.testing.assert_allclose(dst[0, :-1, :-1], exp[:, :, 2], atol=1e-6)
    assert dst.shape == (3, 64, 64)


def test_warpaffine_letterbox_padding_value(oracle):
    img = synth.frames(1, seed=2, h=90, w=160)[0]
    dst = oracle.warpaffine(img, 64, 64)
    # 160x90 -> scale 0.4 -> 64x36 band, rows < 13 and > 51 are pure border 128/255
    assert np.allclose(dst[:, :12, :], 128 / 255.0)
    assert np.allclose(dst[:, 52:, :], 128 / 255.0)
    assert not np.allclose(dst[:, 20:40, :], 128 / 255.0)


def test_rcnn_oracles_against_torchvision(oracle):
    import torchvision

    # rpnNms: greedy, class-agnostic, thr 0.7, then first `post` survivors in score order
    rng = np.random.default_rng(0)
    n = 600
    xy = rng.uniform(0, 500, (n, 2)).astype(np.float32)
    wh = rng.uniform(20, 200, (n, 2)).astype(np.float32)
    boxes = np.concatenate([xy, xy + wh], 1)[None]
    scores = rng.standard_normal((1, n)).astype(np.float32)
    out = oracle.rpn_nms(scores, boxes, 100, 0.7)
    keep = torchvision.ops.nms(torch.from_numpy(boxes[0]), torch.from_numpy(scores[0]), 0.7).numpy()
    assert len(keep) >= 100
    assert np.array_equal(out[0], boxes[0][keep[:100]])
    # batchedNms hard mode vs torchvision.batched_nms
    cls = rng.integers(0, 5, (1, n)).astype(np.float32)
    sc = rng.uniform(0.05, 1, (1, n)).astype(np.float32)
    os_, ob, oc = oracle.batched_nms(0, sc, boxes, cls, 100, 0.5)
    keep = torchvision.ops.batched_nms(torch.from_numpy(boxes[0]), torch.from_numpy(sc[0]),
                                       torch.from_numpy(cls[0].astype(np.int64)), 0.5).numpy()
    k = min(100, len(keep))
    assert np.array_equal(os_[0, :k], sc[0][keep[:k]])
    assert np.array_equal(ob[0, :k], boxes[0][keep[:k]])
    assert np.array_equal(oc[0, :k], cls[0][keep[:k]])


def test_rpn_decode_topn_and_clip(oracle):
    scores, deltas = synth.rpn_inputs(2, seed=1, A=15, H=10, W=12)
    anchors = synth.rcnn_anchors()
    os_, ob = oracle.rpn_decode(scores, deltas, 160, 192, 16.0, anchors, 300)
    for b in range(2):
        flat = scores[b].reshape(-1)
        order = np.argsort(-flat, kind="stable")[:300]
        valid = os_[b] > -1e38
        assert np.array_equal(os_[b][valid], flat[order][valid])
        assert ob[b][:, 0].min() >= 0 and ob[b][:, 2].max() <= 192 and ob[b][:, 3].max() <= 160
        assert np.all(np.diff(flat[order]) <= 0)


def test_resize_bilinear_equals_cv2(oracle):
    """cv::resize(CV_32FC1, INTER_LINEAR) as process_mask uses it (160x160 -> 640x640, yolov8_seg.cpp:55): OpenCV's SIMD
    build may fuse the multiply-adds, so <= 1 ulp of a value in [0, 1]."""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(5)
    for (sh, sw, dh, dw) in ((160, 160, 640, 640), (96, 160, 384, 640), (40, 24, 160, 96)):
        src = rng.random((sh, sw), dtype=np.float32)
        src[rng.random((sh, sw)) < 0.5] = 0.0          # masks are zero outside the box
        ref = cv2.resize(src, (dw, dh), interpolation=cv2.INTER_LINEAR)
        got = oracle.resize_bilinear(src, dh, dw)
        assert np.abs(got - ref).max() <= 1.2e-7


@pytest.mark.parametrize("variant", [0, 1])
def test_process_mask_against_numpy_and_cv2(oracle, variant):
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(6 + variant)
    proto = rng.standard_normal((32, 160, 160)).astype(np.float32)
    for bbox in ([100.0, 60.0, 200.0, 120.0], [-30.0, 500.0, 180.0, 300.0], [620.0, 10.0, 5.0, 7.0], [300.0, 300.0, 0.5, 0.5]):
        coeffs = rng.standard_normal(32).astype(np.float32) * 0.5
        got = oracle.process_mask(variant, proto, bbox, coeffs)
        b = np.asarray(bbox, np.float32)
        if variant == 0:   # yolov8_seg.cpp:17-34: x, y, w, h; clamp; /4; int()
            l, t, r, bt = b[0], b[1], b[0] + b[2], b[1] + b[3]
            l, t, r, bt = max(l, np.float32(0)), max(t, np.float32(0)), min(r, np.float32(640)), min(bt, np.float32(640))
            l, t, r, bt = (np.float32(v) / np.float32(4) for v in (l, t, r, bt))
            rx, ry, rw, rh = int(l), int(t), int(r - l), int(bt - t)
        else:              # yolov5 postprocess.cpp:94-104: cx, cy, w, h; /4; round()
            l, t, r, bt = b[0] - b[2] / 2, b[1] - b[3] / 2, b[0] + b[2] / 2, b[1] + b[3] / 2
            l, t, r, bt = (np.float32(v) / np.float32(4) for v in (l, t, r, bt))
            rnd = lambda v: int(np.floor(abs(v) + 0.5) * np.sign(v))  # C round(): half away from zero
            rx, ry, rw, rh = rnd(l), rnd(t), rnd(r - l), rnd(bt - t)
        m = np.zeros((160, 160), np.float32)
        xs, ys = range(max(rx, 0), min(rx + rw, 160)), range(max(ry, 0), min(ry + rh, 160))
        for y in ys:
            for x in xs:
                e = np.float32(0)
                for j in range(32):
                    e = np.float32(e + np.float32(coeffs[j] * proto[j, y, x]))
                m[y, x] = np.float32(1) / (np.float32(1) + np.exp(-e, dtype=np.float32))
        ref = cv2.resize(m, (640, 640), interpolation=cv2.INTER_LINEAR)
        assert np.abs(got - ref).max() <= 3e-7
        assert (got > 0).any() == (len(xs) > 0 and len(ys) > 0)


@pytest.mark.parametrize("sampling", [0, 2])
def test_roi_align_against_torchvision(oracle, sampling):
    """RoIAlignForward (rcnn/RoiAlign.cu) is detectron2's aligned ROIAlign: torchvision.ops.roi_align(aligned=True)."""
    tv = pytest.importorskip("torchvision")
    import torch
    rng = np.random.default_rng(9)
    Cc, H, W, Pp, N = 6, 50, 67, 14, 40
    feat = rng.standard_normal((Cc, H, W)).astype(np.float32)
    x1 = rng.uniform(0, 900, N); y1 = rng.uniform(0, 700, N)
    w = np.exp(rng.uniform(np.log(16), np.log(800), N)); h = np.exp(rng.uniform(np.log(16), np.log(600), N))
    rois = np.stack([x1, y1, x1 + w, y1 + h], -1).astype(np.float32)
    got = oracle.roi_align(rois, feat, Pp, 1 / 16, sampling)
    boxes = torch.cat([torch.zeros(N, 1), torch.from_numpy(rois)], 1)
    ref = tv.ops.roi_align(torch.from_numpy(feat)[None], boxes, (Pp, Pp), 1 / 16, sampling, aligned=True).numpy()
    np.testing.assert_allclose(got, ref, rtol=0, atol=3e-5)



# ------------------------------------------------------------------ SURVEY 8f rank 4 variants: numpy re-derivations
def _sig(x):
    return (1.0 / (1.0 + np.exp(-x.astype(np.float64))))


def test_yolov3_decode_oracle_vs_numpy(oracle):
    """yolov3-spp/yololayer.cu:148-191 re-derived with numpy (fp64 sigmoid/exp, so values to 1e-4; counts, classes, anchor ids exact)."""
    from tensorrtx_b200 import plugins  # noqa: F401  (YOLOV3_KERNELS needs the library; skip cleanly if it is not built)
    B = 2
    heads = synth.yolov3_heads(B, seed=5, net_w=416, net_h=320)
    anchors = [(116, 90, 156, 198, 373, 326), (30, 61, 62, 45, 59, 119), (10, 13, 16, 30, 33, 23)]
    out, idx = oracle.yolov3_decode(heads, anchors)
    exp = [[] for _ in range(B)]
    off = 0
    for h, s, anc in zip(heads, (32, 16, 8), anchors):
        gh, gw = h.shape[2], h.shape[3]
        x = h.reshape(B, 3, 85, gh * gw)
        obj = _sig(x[:, :, 4])
        cls_p = _sig(x[:, :, 5:])
        best, arg = cls_p.max(2), cls_p.argmax(2)
        for b in range(B):
            for e in range(gh * gw):
                for k in range(3):
                    if np.float32(best[b, k, e]) < np.float32(0.1) or np.float32(obj[b, k, e]) < np.float32(0.1):
                        continue
                    row, col = divmod(e, gw)
                    exp[b].append((off + e * 3 + k, (col + _sig(x[b, k, 0, e])) * s, (row + _sig(x[b, k, 1, e])) * s,
                                   np.exp(np.float64(x[b, k, 2, e])) * anc[2 * k], np.exp(np.float64(x[b, k, 3, e])) * anc[2 * k + 1],
                                   obj[b, k, e], arg[b, k, e], best[b, k, e]))
        off += gh * gw * 3
    for b in range(B):
        n = int(out[b, 0])
        assert n == len(exp[b]) and n > 20
        rows = out[b, 1:1 + n * 7].reshape(n, 7)
        e = np.asarray(sorted(exp[b]), dtype=np.float64)        # the oracle walks level by level, cell by cell: same order
        assert np.array_equal(idx[b, :n], e[:, 0].astype(np.int32))
        np.testing.assert_allclose(rows[:, :4], e[:, 1:5], rtol=1e-5, atol=1e-4)
        np.testing.assert_allclose(rows[:, [4, 6]], e[:, [5, 7]], atol=1e-6)
        assert np.array_equal(rows[:, 5], e[:, 6])


def test_yolo26_gather_oracle_vs_numpy(oracle):
    rows = synth.yolo26_rows(2, seed=6, nc=15, anchors=2000, obb=True)
    out, idx = oracle.yolo26_gather(rows, nc=15, obb=True, max_out=300, conf_thresh=0.3)
    for b in range(2):
        sc = rows[b, :, 4:19]
        best, arg = sc.max(1), sc.argmax(1)
        keep = np.where(~(best < np.float32(0.3)))[0]
        n = int(out[b, 0])
        assert n == len(keep) and n > 20
        r = out[b, 1:1 + n * 90].reshape(n, 90)
        assert np.array_equal(idx[b, :n], keep)
        assert np.array_equal(r[:, :4], rows[b, keep, :4]) and np.array_equal(r[:, 4], best[keep])
        assert np.array_equal(r[:, 5], arg[keep].astype(np.float32)) and np.array_equal(r[:, 89], rows[b, keep, 19])
        assert np.all(r[:, 6:89] == 0) and np.all(out[b, 1 + n * 90:] == 0)


def test_anticov_decode_oracle_vs_numpy(oracle):
    heads = synth.anticov_heads(2, seed=7)
    out, idx = oracle.anticov_decode(heads)
    for b in range(2):
        exp = []
        off = 0
        for h, step, anchor in zip(heads, (8, 16, 32), (16, 64, 256)):
            w = 640 // step
            g = h.shape[2]
            for e in range(g):
                for k in range(2):
                    conf = h[b, 2 + k, e]
                    if conf < 0.5:
                        continue
                    y, x = divmod(e, w)
                    p0, p1, p2 = 7.5 + x * step, 7.5 + y * step, float(anchor * 2 // (k + 1))
                    bx = h[b, 4 + 4 * k:8 + 4 * k, e].astype(np.float64)
                    bw, bh = p2 * np.exp(bx[2]), p2 * np.exp(bx[3])
                    x1, y1 = p0 + bx[0] * p2 - (bw - 1) / 2, p1 + bx[1] * p2 - (bh - 1) / 2
                    lm = h[b, 12 + 10 * k:22 + 10 * k, e].astype(np.float64)
                    lmk = [(p0 if i % 2 == 0 else p1) + lm[i] * 0.2 * p2 for i in range(10)]
                    exp.append([off + e * 2 + k, x1, y1, x1 + bw, y1 + bh, conf] + lmk + [h[b, 36 + k, e]])
            off += g * 2
        n = int(out[b, 0])
        assert n == len(exp) and n > 50
        e = np.asarray(exp, np.float64)
        assert np.array_equal(idx[b, :n], e[:, 0].astype(np.int32))
        np.testing.assert_allclose(out[b, 1:1 + n * 16].reshape(n, 16), e[:, 1:], rtol=1e-5, atol=1e-3)


def test_resize_of_a_crop_matches_cv2(oracle):
    """scale_mask = cv::resize(mask(r), img.size()): the oracle's bilinear restatement on a cropped, strided view against
    cv2.resize (INTER_LINEAR) for up- and down-scaling targets."""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(8)
    mask = rng.uniform(0, 1, (640, 640)).astype(np.float32)
    for (x, y, w, h, W, H) in ((0, 140, 640, 360, 1920, 1080), (140, 0, 360, 640, 1080, 1920), (0, 0, 640, 640, 333, 333), (0, 80, 640, 480, 500, 375)):
        crop = np.ascontiguousarray(mask[y:y + h, x:x + w])
        got = oracle.resize_bilinear(crop, H, W)
        # OpenCV's own C++ kernel (what the restatement follows): within 1 ulp (its SIMD path contracts to FMA).  With the
        # optional IPP accelerator (on in this cv2 build) the coefficients are computed differently: <= 3e-5 at these sizes,
        # inside the 1e-4 tolerance of the north star -- the reference's own result depends on its OpenCV build there.
        ipp = cv2.ipp.useIPP()
        try:
            cv2.ipp.setUseIPP(False)
            np.testing.assert_allclose(got, cv2.resize(crop, (W, H), interpolation=cv2.INTER_LINEAR), atol=2e-7, rtol=0)
        finally:
            cv2.ipp.setUseIPP(ipp)
        np.testing.assert_allclose(got, cv2.resize(crop, (W, H), interpolation=cv2.INTER_LINEAR), atol=1e-4, rtol=0)
