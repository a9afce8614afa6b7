"""roi_align_window_kernel<14> (tensorrtx_b200/csrc/roi_align.cu) replayed on the CPU, formula by formula: sample tables, hull of
the valid taps, window-relative byte offsets, channels per pass and row pitch, the cp.async fill items ((cell, channel % 4), a
warp = 4 channels x 8 cells, the row of a cell by multiplication), the tap rounds of 8 channels through 16-byte shared loads, the
exact reciprocal for power-of-two sample counts.  The replay works on a flat float32 array of exactly the kernel's 16384 shared
floats and must reproduce the oracle's RoIAlign (= the reference's, tests/test_vs_reference_gpu.py) BIT FOR BIT on maps and
proposals the GPU test does not cover (other H x W, windows at every border, 1-cell windows, whole-map windows that need many
passes of 4 channels).  It pins the index algebra of the kernel; the CUDA code itself is pinned on hardware."""
import numpy as np
import pytest

F = np.float32
WINDOW_FLOATS, THREADS, CU, CPC, MAX_TABLE = 16384, 224, 8, 64, 256   # kRoiWindowFloats, kRoiWindowThreads, kRoiCU, channels per CTA


def _axis_tap(y, size):
    valid = not (float(y) < -1.0 or float(y) > size)
    if y <= 0:
        y = F(0)
    low = int(y)
    if low >= size - 1:
        high = low = size - 1
        y = F(low)
    else:
        high = low + 1
    l = F(y - F(low))
    h = F(1.0 - float(l))                      # `1. - l`: a double subtraction rounded to float
    return low, high, l, h, valid


def emulate(rois, feat, P, scale, sampling):
    C, H, W = feat.shape
    N = rois.shape[0]
    out = np.full((N, C, P, P), np.nan, F)
    scale = F(scale)
    used = {"passes": 0, "magic": 0, "division": 0, "pow2": 0, "div": 0, "narrow": 0}
    for n in range(N):
        r = rois[n].astype(F)
        start_w, start_h = F(r[0] * scale - F(0.5)), F(r[1] * scale - F(0.5))
        end_w, end_h = F(r[2] * scale - F(0.5)), F(r[3] * scale - F(0.5))
        roi_w, roi_h = F(end_w - start_w), F(end_h - start_h)
        bin_h, bin_w = F(roi_h / F(P)), F(roi_w / F(P))
        gh = sampling if sampling > 0 else int(np.ceil(F(roi_h / F(P))))
        gw = sampling if sampling > 0 else int(np.ceil(F(roi_w / F(P))))
        count = F(gh * gw)
        assert gh > 0 and gw > 0 and P * gh <= MAX_TABLE and P * gw <= MAX_TABLE, "the test only feeds tabulated proposals"
        s_y = [_axis_tap(F(F(start_h + F(ph * bin_h)) + F(F(F(iy + 0.5) * bin_h) / F(gh))), H) for ph in range(P) for iy in range(gh)]
        s_x = [_axis_tap(F(F(start_w + F(pw * bin_w)) + F(F(F(ix + 0.5) * bin_w) / F(gw))), W) for pw in range(P) for ix in range(gw)]
        vy, vx = [t for t in s_y if t[4]], [t for t in s_x if t[4]]
        if not vy or not vx:                    # every sample outside the map
            out[n] = F(0) / count
            continue
        y0, y1 = min(t[0] for t in vy), max(t[1] for t in vy)
        x0, x1 = min(t[0] for t in vx), max(t[1] for t in vx)
        wh, ww = y1 - y0 + 1, x1 - x0 + 1
        cells = wh * ww
        cpass = min(CPC, (WINDOW_FLOATS // cells - 4) & ~7)
        pitch = cpass + 4
        if cpass < 8:
            cpass, pitch = 4, 4
            used["narrow"] += 1
        assert cells * pitch <= WINDOW_FLOATS
        w_y = [(((t[0] - y0) * ww * pitch * 4) if t[4] else -1, (t[1] - y0) * ww * pitch * 4, t[2], t[3]) for t in s_y]
        w_x = [(((t[0] - x0) * pitch * 4) if t[4] else -1, (t[1] - x0) * pitch * 4, t[2], t[3]) for t in s_x]
        ngroups, m_ww = (cells + 7) >> 3, (65536 + ww - 1) // ww
        pow2 = (gh * gw) & (gh * gw - 1) == 0
        used["pow2" if pow2 else "div"] += 1
        for c_begin in range(0, C, CPC):
            c_end = min(C, c_begin + CPC)
            for cb in range(c_begin, c_end, cpass):
                nc = min(cpass, c_end - cb)
                used["passes"] += 1
                smem = np.full(WINDOW_FLOATS, np.nan, F)           # stale shared memory
                nq, ntail = nc >> 2, nc & 3
                assert ntail == 0, "the dispatcher only sends channel counts that are a multiple of 4 to this kernel"
                e = np.arange(ngroups * 32, dtype=np.int64)         # every thread's fill items
                fx, fc = e & 7, (e >> 3) & 3
                cell = (e >> 5) * 8 + fx
                ok = cell < cells
                e, fx, fc, cell = e[ok], fx[ok], fc[ok], cell[ok]
                if cells * ww <= 65536:
                    y = (cell * m_ww) >> 16
                    used["magic"] += 1
                else:
                    y = cell // ww
                    used["division"] += 1
                x = cell - y * ww
                assert np.array_equal(y, cell // ww)
                for cq in range(nq):
                    smem[cell * pitch + fc + 4 * cq] = feat[cb + fc + 4 * cq, y0 + y, x0 + x]
                for k0 in range(0, nc, CU):
                    nch = min(CU, nc - k0)
                    for ph in range(P):
                        for pw in range(P):
                            acc = np.zeros(CU, F)
                            for iy in range(gh):
                                ty = w_y[ph * gh + iy]
                                if ty[0] < 0:
                                    continue
                                for ix in range(gw):
                                    tx = w_x[pw * gw + ix]
                                    if tx[0] < 0:
                                        continue
                                    ly, hy, lx, hx = ty[2], ty[3], tx[2], tx[3]
                                    w1, w2, w3, w4 = F(hy * hx), F(hy * lx), F(ly * hx), F(ly * lx)
                                    groups = range(CU // 4) if nch >= CU else [g for g in range(CU // 4) if 4 * g < nch]
                                    for g in groups:   # one LDS.128 per tap: byte offset / 4 + k0 + 4g .. + 3
                                        i1, i2 = (ty[0] + tx[0]) // 4 + k0 + 4 * g, (ty[0] + tx[1]) // 4 + k0 + 4 * g
                                        i3, i4 = (ty[1] + tx[0]) // 4 + k0 + 4 * g, (ty[1] + tx[1]) // 4 + k0 + 4 * g
                                        assert max(i1, i2, i3, i4) + 3 < WINDOW_FLOATS and (ty[0] + tx[0]) % 16 == 0
                                        v1, v2, v3, v4 = smem[i1:i1 + 4], smem[i2:i2 + 4], smem[i3:i3 + 4], smem[i4:i4 + 4]
                                        val = (w1 * v1 + w2 * v2) + w3 * v3
                                        val = val + w4 * v4
                                        acc[4 * g:4 * g + 4] += val
                            res = acc * F(F(1.0) / count) if (pow2 and nch >= CU) else acc / count
                            out[n, cb + k0:cb + k0 + nch, ph, pw] = res[:nch]
    return out, used


@pytest.mark.parametrize("shape,sampling", [((8, 50, 67), 0), ((12, 20, 23), 0), ((8, 64, 64), 0), ((4, 7, 90), 2), ((8, 33, 3), 3)])
def test_window_kernel_algebra_reproduces_the_oracle(oracle, shape, sampling):
    C, H, W = shape
    rng = np.random.default_rng(C * 1000 + H + sampling)
    feat = rng.standard_normal((C, H, W)).astype(F)
    sw, sh = W * 16.0, H * 16.0
    x1, y1 = rng.uniform(-40, sw, 14), rng.uniform(-40, sh, 14)
    w_, h_ = np.exp(rng.uniform(np.log(4), np.log(sw * 1.2), 14)), np.exp(rng.uniform(np.log(4), np.log(sh * 1.2), 14))
    rois = np.stack([x1, y1, x1 + w_, y1 + h_], -1).astype(F)
    extra = np.array([[0, 0, sw, sh],                    # the whole map: passes of 4 channels, row pitch 4, division for the cell's row
                      [5, 5, 9, 9],                      # a fraction of one cell
                      [sw - 8, sh - 8, sw + 90, sh + 70],  # hangs over the bottom-right corner
                      [-300, -300, -200, -100],          # entirely outside
                      [16 * 3, 16 * 2, 16 * 3 + 1, sh],  # one column wide, full height
                      [0, 16 * 5, sw, 16 * 5 + 2]], F)   # one row high, full width
    rois = np.concatenate([rois, extra])
    if sampling == 0:                                    # keep the adaptive grids inside the 256-entry tables (as the kernel's dispatcher does)
        rois = rois[np.ceil((rois[:, 2] - rois[:, 0]) / 16 / 14) * 14 <= MAX_TABLE]
    got, used = emulate(rois, feat, 14, 1 / 16, sampling)
    ref = oracle.roi_align(rois, feat, 14, 1 / 16, sampling)
    assert np.array_equal(np.isnan(got), np.isnan(ref))
    assert np.array_equal(got[~np.isnan(got)], ref[~np.isnan(ref)]), np.abs(got - ref).max()
    assert used["passes"] >= len(rois) - 4 and used["magic"] > 0
    if shape == (8, 50, 67):
        assert used["narrow"] > 0 and used["division"] > 0 and used["pow2"] > 0
