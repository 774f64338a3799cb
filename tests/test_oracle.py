"""Known-answer and self-consistency tests of the CPU oracle (SURVEY §4 tier 1): Snell round trip,
quartic residual / numpy.roots cross-check, weight KATs, depth sampling, label-vs-curve
consistency.  CPU only."""
import ctypes as C

import numpy as np
import pytest

from oracle import oracle_api as O
from stereoreconstruction_b200 import scenes, types as T
from scene_util import refractive_arc_scene, rectified_scene


def test_project_unproject_round_trip_refractive_no_distortion():
    w, h = 64, 48
    cams = scenes.arc_cameras(3, w, h, arc_deg=20.0, distortion=None)
    sc = O.Scene(cams, [np.zeros((h, w, 4), np.uint8)] * 3)
    gx, gy = np.meshgrid(np.arange(w) + 0.5, np.arange(h) + 0.5)
    want = np.stack([gx, gy], -1).reshape(-1, 2)
    for v in range(3):
        rays = sc.unproject_grid(v)
        for t in (300.0, 520.0):
            pts = (rays[..., :3] + t * rays[..., 3:]).reshape(-1, 3)
            for mode in (0, 1):
                xy, ok = sc.project_points(v, pts, root_mode=mode)
                assert ok.all()
                assert np.abs(xy - want).max() < 1e-8


def test_quartic_and_monotone_roots_agree_and_stats_are_clean():
    cams, imgs, ms, surf = refractive_arc_scene(V=3, w=48, h=32)
    sc = O.Scene(cams, imgs)
    rays = sc.unproject_grid(0)
    rng = np.random.RandomState(0)
    pts = (rays[..., :3] + rng.uniform(350, 650, rays.shape[:2] + (1,)) * rays[..., 3:]).reshape(-1, 3)
    O.stats_reset()
    for v in (1, 2):
        sc.project_points(v, pts, root_mode=2)
    st = O.stats()
    assert st["project_calls"] == 2 * pts.shape[0]
    assert st["quartic_fail"] == 0 and st["no_root"] == 0 and st["root_mismatch"] == 0
    assert st["max_root_diff"] < 1e-9


def test_snell_root_against_numpy_roots():
    L = O.lib()
    rng = np.random.RandomState(1)
    for _ in range(300):
        r, d, h, n = rng.uniform(0.1, 400), rng.uniform(1, 60), rng.uniform(50, 900), rng.uniform(1.05, 1.6)
        x = L.orc_snell_root(r, d, h, n)
        nn, dd, rr = n * n, d * d, r * r
        coeffs = [nn - 1, -2 * r * (nn - 1), rr * (nn - 1) + dd * nn - h * h, -2 * dd * nn * r, dd * nn * rr]
        roots = np.roots(coeffs)
        real = roots[np.abs(roots.imag) < 1e-9].real
        inside = real[(real >= -1e-9) & (real <= r + 1e-9)]
        assert inside.size == 1, (r, d, h, n, roots)
        assert abs(inside[0] - x) < 1e-8 * max(1.0, r)
        # residual of the un-squared Snell equation
        g = x / np.hypot(x, d) - n * (r - x) / np.hypot(r - x, h)
        assert abs(g) < 1e-14
        re, im = np.empty(4), np.empty(4)
        c5 = np.array(coeffs[::-1], dtype=np.float64)
        assert L.orc_poly_roots4(c5.ctypes.data_as(C.POINTER(C.c_double)), re.ctypes.data_as(C.POINTER(C.c_double)),
                                 im.ctypes.data_as(C.POINTER(C.c_double)))
        mine = np.sort_complex(re + 1j * im)
        assert np.allclose(mine, np.sort_complex(roots), rtol=1e-7, atol=1e-7)


def test_weight_kats():
    h, w = 21, 23
    const = np.full((h, w, 4), 90, np.uint8)
    const[..., 3] = 255
    cam = T.make_camera(np.eye(3), np.eye(3), np.zeros(3))
    sc = O.Scene([cam], [const])
    for r in (1, 2, 5):
        g = sc.weights(0, T.SR_WEIGHT_GEODESIC, r, [10], [10])[0]
        assert (g == 1.0).all()  # constant image: every geodesic distance is 0
        a = sc.weights(0, T.SR_WEIGHT_ADAPTIVE, r, [10], [10])[0]
        assert a[r, r] == 1.0
        i = np.arange(-r, r + 1)
        want = np.exp(-np.abs(i)[:, None] / r) * np.exp(-np.abs(i)[None, :] / r)
        assert np.allclose(a, want, rtol=1e-15)
        # window hanging over the image corner: out-of-image taps weigh 0
        g0 = sc.weights(0, T.SR_WEIGHT_GEODESIC, r, [0], [0])[0]
        assert (g0[:r, :] == 0).all() and (g0[:, :r] == 0).all() and (g0[r:, r:] == 1).all()
        a0 = sc.weights(0, T.SR_WEIGHT_ADAPTIVE, r, [0], [0])[0]
        assert (a0[:r, :] == 0).all() and (a0[:, :r] == 0).all()
    # a colour step: geodesic weight across the step is exp(-|step|/50)
    img = const.copy()
    img[:, 12:, :3] = 140
    sc2 = O.Scene([cam], [img])
    g = sc2.weights(0, T.SR_WEIGHT_GEODESIC, 2, [11], [10])[0]
    step = np.sqrt(3 * 50.0 ** 2)
    assert np.allclose(g[:, :3], 1.0) and np.allclose(g[:, 3:], np.exp(-step / 50.0))


def test_depth_from_label():
    L = O.lib()
    P = T.default_params(False, 100.0, 500.0, 256)
    d = np.array([L.orc_depth_from_label(C.byref(O.as_params(P)), k) for k in range(256)])
    assert d[0] == 100.0 and d[-1] == 500.0
    inv = 1.0 / d  # max = 5*min  =>  uniform in inverse depth
    assert np.allclose(np.diff(inv), np.diff(inv)[0], rtol=1e-9)
    P2 = T.default_params(True, 300.0, 800.0, 100)
    d2 = np.array([L.orc_depth_from_label(C.byref(O.as_params(P2)), k) for k in range(100)])
    assert np.allclose(np.diff(d2), 500.0 / 99)


def test_cost_identities():
    cams, imgs, ms, surf = rectified_scene(w=96, h=40)
    sc = O.Scene([cams[0], cams[0]], [imgs[0], imgs[0]])
    for kind, r in ((T.SR_WEIGHT_ADAPTIVE, 3), (T.SR_WEIGHT_GEODESIC, 2)):
        P = T.default_params(False, 100.0, 500.0, 8, radius=r, weight_kind=kind)
        # identical windows: NCC cost 0 (up to rounding), SAD cost 0
        assert abs(sc.cost(P, 0, 1, 40, 20, 40, 20)) < 1e-9
        P.cost_kind = T.SR_COST_SAD_TWOVIEW
        assert sc.cost(P, 0, 1, 40, 20, 40, 20) == 0.0
        P.cost_kind = T.SR_COST_NCC_MVS
        assert abs(sc.cost(P, 0, 1, 40, 20, 40, 20) - 1.0) < 1e-12
        # a window entirely outside the neighbour image
        P.cost_kind = T.SR_COST_NCC_TWOVIEW
        assert sc.cost(P, 0, 1, 40, 20, -500, 20) == 1000.0  # BAD_RET
        P.cost_kind = T.SR_COST_NCC_MVS
        assert sc.cost(P, 0, 1, 40, 20, -500, 20) == 0.0


def test_label_and_curve_modes_reconstruct_the_same_surface():
    cams, imgs, ms, surf = refractive_arc_scene(V=3, w=80, h=48, arc_deg=16.0)
    sc = O.Scene(cams, imgs)
    P = T.default_params(True, 430.0, 570.0, 48)
    rays = sc.unproject_grid(1)
    gt = (surf.hit(rays) - np.array(cams[1].C[:])) @ np.array(cams[1].prin_dir[:])
    dl, il, _, _, _ = sc.mvs_view(P, 1, [0, 2])
    dc, _, _, _, _ = sc.mvs_view(P, 1, [0, 2], curve_mode=True)
    okl, okc = il >= 0, dc > 0
    assert okl.mean() > 0.8 and okc.mean() > 0.8
    # depth resolution of this miniature rig: one pixel of disparity = Z^2 / (f * baseline) units
    f = cams[1].K[0]
    base = np.linalg.norm(np.array(cams[1].C[:]) - np.array(cams[0].C[:]))
    px = 500.0 ** 2 / (f * base)
    assert np.median(np.abs(dl - gt)[okl]) < 0.75 * px
    assert np.median(np.abs(dc - gt)[okc]) < 0.75 * px


def test_mvs_curve_peaks_are_sorted_topk():
    cams, imgs, ms, surf = refractive_arc_scene(V=3, w=48, h=32, arc_deg=16.0)
    sc = O.Scene(cams, imgs)
    P = T.default_params(True, 430.0, 570.0, 32)
    d, _, b, _, peaks = sc.mvs_view(P, 1, [0, 2], curve_mode=True, want_peaks=True)
    assert (np.diff(peaks[..., 0], axis=-1) >= 0).all()  # ascending ncc, best last
    assert (peaks[..., -1, 1] == d).all() and (peaks[..., -1, 0] == b).all()
    none = d == -1
    assert (peaks[none][:, :, 0] == 0).all()


def test_neighbour_selection_matches_numpy_restatement():
    cams = scenes.arc_cameras(8, 64, 48)
    sc = O.Scene(cams, [np.zeros((48, 64, 4), np.uint8)] * 8)
    assert [[int(v) for v in r] for r in sc.select_neighbours(3)] == scenes.nearest_neighbours(cams, 3)


def test_row_band_equals_full_run():
    cams, imgs, ms, surf = refractive_arc_scene(V=3, w=48, h=32)
    sc = O.Scene(cams, imgs)
    P = T.default_params(False, 420.0, 580.0, 12, radius=2)
    full_d, full_i, _, _ = sc.twoview_label(P, 0, 1)
    P.row_begin, P.row_end = 10, 20
    d, i, _, vol = sc.twoview_label(P, 0, 1, want_volume=True)
    assert (i[10:20] == full_i[10:20]).all() and vol.shape == (10, 48, 12)


def _numpy_costs(img_l, img_r, mask_l, mask_r, W, x1, y1, x2, y2, R):
    """An independent restatement (written from the reference text, numpy, no shared code with
    oracle.cpp) of cost_ncc of MultiViewStereo (stereo/multiviewstereo.cpp:113-189), cost_ncc and
    cost_sad of TwoViewStereo (stereo/twoviewstereo.cpp:909-977, 864-905).  W[row+R][col+R] are the
    support weights of the reference pixel (pinned against the reference's own sources elsewhere)."""
    h, w = img_l.shape[:2]

    def gray(img, x, y):  # RGBA::toGray, util/vectorimage.hpp:60-62
        r, g, b = (float(v) for v in img[y, x, :3])
        return 0.11 * r + 0.59 * g + 0.3 * b

    def pixel_ok(x, y):  # VectorImage::pixel
        return 0 <= x < w and 0 <= y < h

    def sample_ok(x, y):  # VectorImage::sample at integer coordinates
        return x >= 0 and y >= 0 and x + 1 < w and y + 1 < h

    def white(m, x, y):
        return pixel_ok(x, y) and m[y, x] == 255

    def taps(kind):
        out = []
        for row in range(-R, R + 1):
            for col in range(-R, R + 1):
                xl, yl, xr, yr = x1 + col, y1 + row, x2 + col, y2 + row
                if kind == "mvs":
                    if not (pixel_ok(xl, yl) and pixel_ok(xr, yr)):
                        continue
                else:
                    if not (white(mask_l, xl, yl) and white(mask_r, xr, yr)):
                        continue
                    if not sample_ok(xl, yl):
                        continue
                    if not (sample_ok(xr, yr) if kind == "ncc2" else pixel_ok(xr, yr)):
                        continue
                wt = W[row + R, col + R]
                if wt > 1e-10:
                    out.append((wt, gray(img_l, xl, yl), gray(img_r, xr, yr)))
        return out

    def ncc(kind):
        t = taps(kind)
        tot = sum(wt for wt, _, _ in t)
        if tot < 1e-10:
            return None
        mL = sum(wt * gl for wt, gl, _ in t) / tot
        mR = sum(wt * gr for wt, _, gr in t) / tot
        s1 = sum((wt * gl - mL) * (wt * gr - mR) for wt, gl, gr in t)
        s2 = sum((wt * gl - mL) ** 2 for wt, gl, _ in t)
        s3 = sum((wt * gr - mR) ** 2 for wt, _, gr in t)
        return s1, s2, s3

    r = ncc("mvs")
    c_mvs = 0.0 if r is None or r[1] * r[2] < 1e-10 else r[0] / np.sqrt(r[1] * r[2])
    r = ncc("ncc2")
    if r is None:
        c_two = 1000.0
    else:
        with np.errstate(invalid="ignore", divide="ignore"):
            v = 255 * (1.0 - abs(r[0]) / np.sqrt(np.float64(r[1] * r[2])))
        c_two = 120.0 if not (v < 120.0) else float(v)  # std::min(120.0, NaN) is 120
    t = taps("sad")
    tot = sum(wt for wt, _, _ in t)
    c_sad = 1000.0 if (len(t) <= 4 or tot <= 1e-10) else sum(wt * min(120.0, abs(gl - gr)) for wt, gl, gr in t) / tot
    return c_mvs, c_two, c_sad


def test_costs_against_independent_numpy_restatement():
    rng = np.random.RandomState(17)
    h, w = 40, 56
    base = rng.randint(0, 256, size=(h, w, 3))
    imgs = []
    for k in range(2):
        im = np.empty((h, w, 4), np.uint8)
        im[..., :3] = np.clip(base + rng.randint(-12, 13, size=base.shape) + 3 * k, 0, 255)
        im[..., 3] = 255
        imgs.append(im)
    masks = [np.where(rng.rand(h, w) > 0.1, 255, 0).astype(np.uint8) for _ in range(2)]
    cams = [T.make_camera(np.eye(3), np.eye(3), np.zeros(3)), T.make_camera(np.eye(3), np.eye(3), np.array([-1.0, 0, 0]))]
    sc = O.Scene(cams, imgs, masks)
    for R, kind in ((2, T.SR_WEIGHT_GEODESIC), (2, T.SR_WEIGHT_ADAPTIVE), (5, T.SR_WEIGHT_GEODESIC)):
        pts = [(rng.randint(-1, w + 1), rng.randint(-1, h + 1), rng.randint(-3, w + 3), rng.randint(-3, h + 3)) for _ in range(120)]
        pts += [(0, 0, 0, 0), (w - 1, h - 1, w - 1, h - 1), (w - 2, h - 2, 1, 1), (R, R, w - 1 - R, h - 1 - R)]
        for (x1, y1, x2, y2) in pts:
            if not (0 <= x1 < w and 0 <= y1 < h):
                continue  # the reference never centres a window outside the reference image
            W = sc.weights(0, kind, R, [x1], [y1])[0]
            want = _numpy_costs(imgs[0], imgs[1], masks[0], masks[1], W, x1, y1, x2, y2, R)
            for cost_kind, ref in ((T.SR_COST_NCC_MVS, want[0]), (T.SR_COST_NCC_TWOVIEW, want[1]), (T.SR_COST_SAD_TWOVIEW, want[2])):
                P = T.default_params(cost_kind == T.SR_COST_NCC_MVS, 1.0, 2.0, 4, radius=R, weight_kind=kind, cost_kind=cost_kind)
                got = sc.cost(P, 0, 1, x1, y1, x2, y2)
                assert abs(got - ref) <= 1e-9 * max(1.0, abs(ref)), (R, kind, cost_kind, x1, y1, x2, y2, got, ref)


def test_calibration_lm_recovers_the_interface():
    """util/lm.cpp + RefractiveCalibrationFunction restated (oracle.cpp, orc_calibration_lm): from a wrong
    first guess the loop drives chi^2 down to the noise floor (the value at the true model) and recovers
    the well-determined parameters; with the reference's text as written (gradient attribution by
    paramIndex / 3, or the inverted solve test) no step is ever accepted."""
    import sys
    import os
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from calib_util import calibration_problem
    cams, pairs, pix, truth, start = calibration_problem(O.Scene)
    fixed = np.zeros(truth.size, np.uint8)
    fixed[0] = 1
    _, _, at_truth, _ = O.calibration_lm(cams, pairs, pix, truth, np.ones(truth.size, np.uint8))
    m, it, c0, c1 = O.calibration_lm(cams, pairs, pix, start, fixed)
    assert c0 > 100 * at_truth and c1 <= 1.05 * at_truth and 5 <= it < 100
    assert m[0] == start[0]                                  # fixed parameters do not move
    assert (np.abs(m - truth)[2::3] < 2.0).all()             # pixel y of every normal
    # chi^2 reported == sum of squared residuals of the cameras configured with the model
    for flags in (dict(exact_attribution=False), dict(literal_check=True)):
        m2, it2, c0b, c1b = O.calibration_lm(cams, pairs, pix, start, fixed, **flags)
        assert (m2 == start).all() and c1b == c0b == c0
