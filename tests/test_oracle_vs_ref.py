"""Pins oracle/oracle.cpp against the REFERENCE'S OWN leaf sources compiled from /root/reference
(oracle/_ref/libref.so: util/lineiter.cpp, util/ray.cpp, util/vectorimage.cpp,
stereo/adaptiveweight.cpp, stereo/geodesicweight.cpp, project/camera.cpp — see oracle/Makefile).
Bit-exact for the leaves; to rounding (1e-12) for the camera model, whose matrix algebra goes
through a stand-in for Eigen on the reference side."""
import ctypes as C

import numpy as np
import pytest

from oracle import oracle_api as O

REF = O.ref_lib()
pytestmark = pytest.mark.skipif(REF is None, reason="oracle/_ref/libref.so not built (no /root/reference here)")


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _lines(fn, args, maxn=4096):
    out = np.empty((maxn, 2), dtype=np.int32)
    n = fn(*args, _ip(out), maxn)
    return n, out[:min(n, maxn)].copy()


def test_line_iterator_and_clip_match_reference():
    L = O.lib()
    rng = np.random.RandomState(11)
    w, h = 64, 48
    cases = [(0, 0, 10, 3), (10, 3, 0, 0), (5, 5, 5, 5), (0, 0, 0, 9), (7, 2, 7, -6), (-5, -5, 70, 60),
             (-20, 10, 90, 12), (30, -40, 31, 90), (63, 47, 0, 0), (64, 48, -1, -1), (100, 100, 200, 150)]
    for _ in range(400):
        cases.append(tuple(int(v) for v in rng.randint(-40, 110, 4)))
    for (x0, y0, x1, y1) in cases:
        for clip in (0, 1):
            n_o, p_o = _lines(L.orc_line, (x0, y0, x1, y1, clip, w, h))
            n_r, p_r = _lines(REF.ref_line, (x0, y0, x1, y1, clip, w, h))
            assert n_o == n_r, (x0, y0, x1, y1, clip)
            assert (p_o == p_r).all(), (x0, y0, x1, y1, clip)
        a = np.array([x0, y0, x1, y1], dtype=np.int32)
        b = a.copy()
        ok_o = L.orc_clip_line(_ip(a), w, h)
        ok_r = REF.ref_clip_line(_ip(b), w, h)
        assert ok_o == ok_r
        if ok_o:
            assert (a == b).all()


def test_ray_ops_match_reference():
    L = O.lib()
    rng = np.random.RandomState(5)
    for _ in range(500):
        s1, d1, s2, d2, pn = (rng.normal(size=3) for _ in range(5))
        pd = float(rng.normal() * 3)
        n = float(rng.uniform(0.6, 1.7))
        a, b = np.empty(3), np.empty(3)
        ra = L.orc_intersect(_dp(s1), _dp(d1), _dp(pn), C.c_double(pd), _dp(a))
        rb = REF.ref_intersect(_dp(s1), _dp(d1), _dp(pn), C.c_double(pd), _dp(b))
        assert ra == rb
        if ra:
            assert (a == b).all()
        a6, b6 = np.empty(6), np.empty(6)
        ra = L.orc_refract(_dp(s1), _dp(d1), _dp(pn), C.c_double(pd), C.c_double(n), _dp(a6))
        rb = REF.ref_refract(_dp(s1), _dp(d1), _dp(pn), C.c_double(pd), C.c_double(n), _dp(b6))
        assert ra == rb and (a6 == b6).all()
        L.orc_closest_points(_dp(s1), _dp(d1), _dp(s2), _dp(d2), _dp(a6))
        REF.ref_closest_points(_dp(s1), _dp(d1), _dp(s2), _dp(d2), _dp(b6))
        assert (a6 == b6).all()


@pytest.fixture(scope="module")
def image():
    rng = np.random.RandomState(2)
    h, w = 37, 53
    img = rng.randint(0, 256, size=(h, w, 4)).astype(np.uint8)
    img[..., 3] = 255
    img[5:9, 7:12] = 255  # a WHITE patch
    return img


def test_pixel_sample_gray_match_reference(image):
    h, w = image.shape[:2]
    from stereoreconstruction_b200 import types as T
    cam = T.make_camera(np.eye(3), np.eye(3), np.zeros(3))
    sc = O.Scene([cam], [image])
    REF.ref_image_create.restype = C.c_void_p
    REF.ref_to_gray.restype = C.c_double
    r = C.c_void_p(REF.ref_image_create(image.ctypes.data_as(C.c_void_p), w, h))
    L = O.lib()
    rng = np.random.RandomState(9)
    a, b = np.empty(4), np.empty(4)
    for _ in range(2000):
        x, y = float(rng.uniform(-2, w + 1)), float(rng.uniform(-2, h + 1))
        L.orc_sample(sc.ptr, 0, C.c_double(x), C.c_double(y), _dp(a))
        REF.ref_sample(r, C.c_double(x), C.c_double(y), _dp(b))
        assert ((a == b) | (np.isnan(a) & np.isnan(b))).all()
    for x in range(-2, w + 2):
        for y in range(-2, h + 2):
            # sampleInt(x,y) == sample((double)x,(double)y) of the reference
            L.orc_sample_int(sc.ptr, 0, x, y, _dp(a))
            REF.ref_sample(r, C.c_double(x), C.c_double(y), _dp(b))
            assert ((a == b) | (np.isnan(a) & np.isnan(b))).all(), (x, y)
    REF.ref_image_destroy(r)


@pytest.mark.parametrize("kind", [0, 1])
@pytest.mark.parametrize("radius", [1, 2, 5, 9])
def test_weights_match_reference(image, kind, radius):
    h, w = image.shape[:2]
    from stereoreconstruction_b200 import types as T
    cam = T.make_camera(np.eye(3), np.eye(3), np.zeros(3))
    sc = O.Scene([cam], [image])
    REF.ref_image_create.restype = C.c_void_p
    r = C.c_void_p(REF.ref_image_create(image.ctypes.data_as(C.c_void_p), w, h))
    rng = np.random.RandomState(4)
    cx = np.concatenate([rng.randint(0, w, 30), [0, w - 1, 0, w - 1]]).astype(np.int32)
    cy = np.concatenate([rng.randint(0, h, 30), [0, 0, h - 1, h - 1]]).astype(np.int32)
    mine = sc.weights(0, kind, radius, cx, cy)
    ref = np.empty_like(mine)
    REF.ref_weights(r, kind, radius, cx.size, _ip(cx), _ip(cy), _dp(ref))
    assert (mine == ref).all()
    REF.ref_image_destroy(r)


# ---- project/camera.cpp: the reference's own Camera class ------------------------------------------
def _cams():
    """Cameras covering every branch: refractive + distorted (cfg4's), refractive only, distorted
    only, plain pinhole."""
    from stereoreconstruction_b200 import scenes
    out = []
    for interface, dist in ((True, (-0.1, 0.05, 0.001, 0.001, 0.0)), (True, (0.0,) * 5),
                            (False, (-0.1, 0.05, 0.001, 0.001, 0.0)), (False, (0.0,) * 5)):
        out += scenes.arc_cameras(4, 640, 480, arc_deg=40.0, interface=interface, distortion=dist)[1:3]
    return out


def _pod(c):
    arr = O.as_cam_array([c])
    return arr


def test_camera_derived_state_matches_reference():
    """Camera::set / setLensDistortion / setRefractiveIndex / setPlane (camera.cpp:205-222,302-342):
    Kinv, Rinv, C, the principal ray and the two flags, as the host code derives them."""
    for c in _cams():
        got = O.OrcCamera()
        REF.ref_camera_derived(_pod(c), C.byref(got))
        for f in ("K", "R", "t", "dist", "plane_n"):
            assert np.allclose(np.array(getattr(got, f)[:]), np.array(getattr(c, f)[:]), rtol=0, atol=1e-15), f
        for f in ("Kinv", "Rinv", "C", "prin_dir"):
            assert np.allclose(np.array(getattr(got, f)[:]), np.array(getattr(c, f)[:]), rtol=1e-13, atol=1e-13), f
        assert got.plane_d == c.plane_d and got.n == c.n
        assert got.is_refractive == c.is_refractive and got.is_distorted == c.is_distorted


def test_camera_unproject_matches_reference():
    """Camera::unproject (camera.cpp:423-459): distortion removal (5 fixed-point rounds), K^-1, refract
    at the interface, local -> global."""
    rng = np.random.RandomState(3)
    xy = np.concatenate([rng.uniform(-20, 660, (400, 1)), rng.uniform(-20, 500, (400, 1))], axis=1)
    xy = np.concatenate([xy, [[319.5, 239.5], [0.5, 0.5], [639.5, 479.5]]])
    for c in _cams():
        ref = np.empty((len(xy), 6))
        REF.ref_camera_unproject(_pod(c), len(xy), _dp(np.ascontiguousarray(xy)), _dp(ref))
        got = np.empty((len(xy), 6))
        L = O.lib()
        for i, (x, y) in enumerate(xy):
            L.orc_unproject(_pod(c), C.c_double(x), C.c_double(y), _dp(got[i]))
        assert np.isfinite(ref).all()
        assert np.abs(got - ref).max() <= 1e-12 * max(1.0, np.abs(ref).max())


@pytest.mark.parametrize("root_mode", [0, 1])
def test_camera_project_matches_reference(root_mode):
    """Camera::project + projectRefraction (camera.cpp:95-138,380-419) — the quartic's coefficients, the
    root acceptance test, the point on the interface, K, distortion with the reference's y-from-distorted-x
    — against the oracle in both root modes (0: the same quartic + acceptance order; 1: the monotone
    1-D solve the GPU uses).  The GSL call inside the reference is answered by the restated solver."""
    rng = np.random.RandomState(4)
    for c in _cams():
        sc = O.Scene([c], [np.zeros((480, 640, 4), np.uint8)])
        rays = sc.unproject_grid(0).reshape(-1, 6)
        rays = rays[rng.randint(0, rays.shape[0], 600)]
        pts = rays[:, :3] + rng.uniform(200, 900, (600, 1)) * rays[:, 3:]       # in front, in view
        pts = np.concatenate([pts, rng.uniform(-400, 400, (200, 3)) + [0, 0, 100]])  # anywhere
        ref_xy = np.empty((len(pts), 2))
        ref_ok = np.empty(len(pts), np.int32)
        REF.ref_camera_project(_pod(c), len(pts), _dp(np.ascontiguousarray(pts)), _dp(ref_xy), _ip(ref_ok))
        xy, ok = sc.project_points(0, pts, root_mode=root_mode)
        assert (ok != 0).mean() > 0.5
        same = (ok != 0) == (ref_ok != 0)
        assert same.mean() >= 0.995, same.mean()  # acceptance at the +-1e-3 band edges may differ by rounding
        both = (ok != 0) & (ref_ok != 0)
        err = np.abs(xy[both] - ref_xy[both]).max(axis=1)
        assert np.quantile(err, 0.995) <= 1e-8, np.quantile(err, 0.995)
        # a different accepted root is only possible inside the |y| < 1e-3 acceptance band
        assert (err > 1e-6).mean() <= 0.005


def test_camera_from_projection_matrix_matches_reference():
    """Camera::setP -> updateOthers (camera.cpp:251-288): RQ factorisation as the host code does it."""
    from stereoreconstruction_b200 import types as T
    for c in _cams()[:2]:
        K = np.array(c.K[:]).reshape(3, 3)
        R = np.array(c.R[:]).reshape(3, 3)
        P = K @ np.hstack([R, np.array(c.t[:])[:, None]])
        got = O.OrcCamera()
        REF.ref_camera_from_P(_dp(np.ascontiguousarray(P * 3.7)), C.byref(got))
        mine = T.camera_from_P(P * 3.7)
        for f in ("K", "R", "t", "C", "prin_dir"):
            a, b = np.array(getattr(got, f)[:]), np.array(getattr(mine, f)[:])
            assert np.allclose(a, b, rtol=1e-9, atol=1e-9), (f, a, b)
        # (the reference divides P by the SQUARED norm of its third row, :252: a scaled P gives a scaled K)
        assert np.allclose(np.array(got.K[:]) / got.K[8], np.array(c.K[:]), rtol=1e-9, atol=1e-7)


# ---- stereo/multiviewstereo.cpp: the reference's own MultiViewStereo, end to end ---------------------
@pytest.mark.parametrize("adaptive", [False, True])
@pytest.mark.parametrize("interface,distortion", [(True, True), (False, True), (False, False)])
def test_mvs_end_to_end_matches_reference(interface, distortion, adaptive):
    """initialize() -> runTask() of the reference's own class (neighbour rule, rasterised epipolar curves,
    weighted NCC, K = 9 peak lists, selection, cross-check; multiviewstereo.cpp:193-247,325-475,524-810) on a
    masked 4-view scene — refractive + lens-distorted, air + distorted, air + pinhole: the oracle reproduces
    every output BIT FOR BIT."""
    import golden_cases as G
    from stereoreconstruction_b200 import types as T
    from scene_util import refractive_arc_scene
    cams, imgs, ms, _ = refractive_arc_scene(V=4, w=96, h=64, masks=True, interface=interface, distortion=distortion)
    mind, maxd, D, cross = G.REF_MVS_CASES["arc"]
    # adaptive: the reference file built with `typedef AdaptiveWeight WeightFunc` (BASELINE configs[1])
    ref = O.RefMVS(cams, imgs, ms, mind, maxd, D, cross, adaptive=adaptive)
    after, nb = ref.run()
    sc = O.Scene(cams, imgs, ms)
    P = T.default_params(True, mind, maxd, D, weight_kind=T.SR_WEIGHT_ADAPTIVE if adaptive else T.SR_WEIGHT_GEODESIC)
    assert nb == [[int(v) for v in r] for r in sc.select_neighbours(3)]
    before = []
    for v in range(len(cams)):
        d, pk = ref.initial_estimate(v)
        od, _, ob, _, op = sc.mvs_view(P, v, nb[v], curve_mode=True, root_mode=0, want_peaks=True)
        before.append(od)
        assert ((d == od) | (np.isnan(d) & np.isnan(od))).all()
        white = ms[v] == 255
        assert (pk[white] == op[white]).all()
        assert (np.isfinite(od) & (od > 0)).mean() > 0.3
    want = sc.crosscheck_mvs(P, before, cross)
    for v in range(len(cams)):
        assert ((after[v] == want[v]) | (np.isnan(after[v]) & np.isnan(want[v]))).all()
        assert 0.2 < (np.isfinite(want[v]) & (want[v] > 0)).mean() < (np.isfinite(before[v]) & (before[v] > 0)).mean()
    # the file-local cost function and the public curve of single pixels
    rng = np.random.RandomState(8)
    for _ in range(40):
        x, y = int(rng.randint(4, 92)), int(rng.randint(4, 60))
        curve = ref.curve(1, 2, x, y)
        mine = sc.epipolar_curve(P, 1, 2, x, y, True, 0)
        assert curve.shape == mine.shape and (curve == mine).all()
        for (px, py) in curve[:: max(1, len(curve) // 5)]:
            assert ref.cost_ncc(1, 2, x, y, px, py) == sc.cost(P, 1, 2, x, y, int(px), int(py))
    ref.close()


# ---- stereo/twoviewstereo.cpp: the reference's own TwoViewStereo, end to end ----------------------------
@pytest.mark.parametrize("interface,distortion", [(True, True), (False, False)])
def test_twoview_end_to_end_matches_reference(interface, distortion):
    """constructor -> computeCostVolumes -> crossCheck of the reference's own class (live rasterised-curve
    search with the second-best test in both directions, cross-check; twoviewstereo.cpp:89-124,233-500,
    596-672) and its two cost functions (cost_ncc :909-977, cost_sad :864-905, windows over borders and
    masked pixels included): the oracle reproduces every output BIT FOR BIT.  (The reference's unqualified
    abs(sum1) is the floating overload of its author's toolchain: ref_shim/precompiled_shim.hpp.)"""
    import golden_cases as G
    from stereoreconstruction_b200 import types as T
    from scene_util import refractive_arc_scene
    cams, imgs, ms, _ = refractive_arc_scene(V=4, w=96, h=64, masks=True, interface=interface, distortion=distortion)
    a, b, mind, maxd, D = G.REF_TWO_CASES["arc"]
    ref = O.RefTwoView(cams[a], cams[b], imgs[a], imgs[b], ms[a], ms[b], mind, maxd, D)
    bl, br = ref.search()
    sc = O.Scene(cams, imgs, ms)
    P = T.default_params(False, mind, maxd, D)
    ol, _, cl = sc.twoview_curve(P, a, b, root_mode=0)
    orr, _, _ = sc.twoview_curve(P, b, a, root_mode=0)

    def same(x, y):
        return ((x == y) | (np.isnan(x) & np.isnan(y))).all()
    assert same(bl, ol) and same(br, orr)
    assert np.isfinite(ol).mean() > 0.3 and np.isinf(ol).any() and cl.max() > 5
    al, ar = ref.run()
    wl, wr = sc.crosscheck_two(P, a, b, ol, orr, 1.0, root_mode=0)
    assert same(al, wl) and same(ar, wr)
    rng = np.random.RandomState(1)
    for _ in range(150):
        x1, y1 = int(rng.randint(0, 96)), int(rng.randint(0, 64))
        x2, y2 = int(rng.randint(-3, 99)), int(rng.randint(-3, 67))
        for sad, kind in ((0, T.SR_COST_NCC_TWOVIEW), (1, T.SR_COST_SAD_TWOVIEW)):
            P2 = T.default_params(False, mind, maxd, D, cost_kind=kind)
            r_, o_ = ref.cost(sad, 0, x1, y1, x2, y2), sc.cost(P2, a, b, x1, y1, x2, y2)
            assert r_ == o_ or (np.isnan(r_) and np.isnan(o_)), (sad, x1, y1, x2, y2, r_, o_)
    ref.close()


def test_label_mode_is_the_reference_pieces_composed():
    """The depth-label sweep (north_star's cost volume + WTA) is compiled out of the reference
    (twoviewstereo.cpp:283,308-329), so it cannot be run there.  Here it is COMPOSED from the reference's own
    compiled pieces — Camera::unproject, intersect (pointFromDepth), Camera::project, the MultiViewStereo
    cost_ncc — with only the loop and the selection rule (multiviewstereo.cpp:589-602,654-660) written in
    this test, for 120 random pixels: index, depth and winning cost equal the oracle's label mode."""
    import golden_cases as G
    from stereoreconstruction_b200 import types as T
    cams, imgs, ms = G.arc_scene()
    mind, maxd, D, cross = G.REF_MVS_CASES["arc"]
    P = T.default_params(True, mind, maxd, D)
    sc = O.Scene(cams, imgs, ms)
    ref_view = 1
    nb = [int(v) for v in sc.select_neighbours(3)[ref_view]]
    od, oi, ob, _, _ = sc.mvs_view(P, ref_view, nb, root_mode=0)
    ref = O.RefMVS(cams, imgs, ms, mind, maxd, D, cross)
    L = O.lib()
    L.orc_depth_from_label.restype = C.c_double
    pp = O.as_params(P)
    depths = np.array([L.orc_depth_from_label(C.byref(pp), d) for d in range(D)])
    cam = cams[ref_view]
    n = np.array(cam.prin_dir[:])
    n = n / np.sqrt(n.dot(n))
    Cc = np.array(cam.C[:])
    rng = np.random.RandomState(12)
    ys, xs = np.nonzero(ms[ref_view] == 255)
    pick = rng.choice(len(ys), 120, replace=False)
    h, w = ms[0].shape
    checked = labelled = 0
    for y, x in zip(ys[pick], xs[pick]):
        ray = np.empty(6)
        REF.ref_camera_unproject(_pod(cam), 1, _dp(np.array([x + 0.5, y + 0.5])), _dp(ray))
        best = None  # (cost, depth, label)
        for j in nb:
            pts, ok = np.empty((D, 3)), np.zeros(D, bool)
            for d in range(D):
                x0 = Cc + n * depths[d]
                out3 = np.empty(3)
                ok[d] = REF.ref_intersect(_dp(ray[:3].copy()), _dp(ray[3:].copy()), _dp(n.copy()), C.c_double(n.dot(x0)), _dp(out3)) != 0
                pts[d] = out3
            xy, pok = np.empty((D, 2)), np.empty(D, np.int32)
            REF.ref_camera_project(_pod(cams[j]), D, _dp(np.ascontiguousarray(pts)), _dp(xy), _ip(pok))
            for d in range(D):
                if not ok[d] or not pok[d]:
                    continue
                tx, ty = int(xy[d, 0]), int(xy[d, 1])  # truncation, multiviewstereo.cpp:773-775
                if not (0 <= tx < w and 0 <= ty < h) or ms[j][ty, tx] != 255:
                    continue
                cost = ref.cost_ncc(ref_view, j, int(x), int(y), tx, ty)
                if cost > 0.95 and (best is None or (cost, depths[d]) > best[:2]):
                    best = (cost, depths[d], d)
        checked += 1
        if best is None:
            assert oi[y, x] == -1 and od[y, x] == -1
        else:
            labelled += 1
            assert oi[y, x] == best[2] and od[y, x] == best[1] and ob[y, x] == best[0], (x, y, best, oi[y, x], ob[y, x])
    assert checked == 120 and labelled > 30
    ref.close()


def test_twoview_label_mode_is_the_reference_pieces_composed():
    """The same for the two-view label sweep (twoviewstereo.cpp:308-329 — the compiled-out branch — with the
    live path's second-best test :304-305): Camera::unproject, intersect, Camera::project and
    TwoViewStereo::cost_ncc are the reference's own compiled code; the loop, the tap rule
    (x*scale - 0.5, truncated by the callee's int parameters) and the selection are written here."""
    import golden_cases as G
    from stereoreconstruction_b200 import types as T
    cams, imgs, ms = G.arc_scene()
    a, b, mind, maxd, D = G.REF_TWO_CASES["arc"]
    P = T.default_params(False, mind, maxd, D)
    sc = O.Scene(cams, imgs, ms)
    od, oi, ob, _ = sc.twoview_label(P, a, b, root_mode=0)
    ref = O.RefTwoView(cams[a], cams[b], imgs[a], imgs[b], ms[a], ms[b], mind, maxd, D)
    L = O.lib()
    L.orc_depth_from_label.restype = C.c_double
    pp = O.as_params(P)
    depths = np.array([L.orc_depth_from_label(C.byref(pp), d) for d in range(D)])
    cam = cams[a]
    n = np.array(cam.prin_dir[:])
    n = n / np.sqrt(n.dot(n))
    Cc = np.array(cam.C[:])
    rng = np.random.RandomState(13)
    ys, xs = np.nonzero(ms[a] == 255)
    pick = rng.choice(len(ys), 60, replace=False)
    finite = 0
    for y, x in zip(ys[pick], xs[pick]):
        ray = np.empty(6)
        REF.ref_camera_unproject(_pod(cam), 1, _dp(np.array([x + 0.5, y + 0.5])), _dp(ray))
        pts, ok = np.empty((D, 3)), np.zeros(D, bool)
        for d in range(D):
            out3 = np.empty(3)
            ok[d] = REF.ref_intersect(_dp(ray[:3].copy()), _dp(ray[3:].copy()), _dp(n.copy()),
                                      C.c_double(n.dot(Cc + n * depths[d])), _dp(out3)) != 0
            pts[d] = out3
        xy, pok = np.empty((D, 2)), np.empty(D, np.int32)
        REF.ref_camera_project(_pod(cams[b]), D, _dp(np.ascontiguousarray(pts)), _dp(xy), _ip(pok))
        min_cost, second, depth, index = np.inf, np.inf, np.nan, -1
        for d in range(D):
            if not ok[d] or not pok[d]:
                continue
            cost = ref.cost(0, 0, int(x), int(y), int(xy[d, 0] - 0.5), int(xy[d, 1] - 0.5))
            if cost + 1e-10 < min_cost:
                second, min_cost, depth, index = min_cost, cost, depths[d], d
        if index >= 0 and min_cost > 0.95 * second:
            depth, index = np.inf, -3
        assert oi[y, x] == index, (x, y, index, oi[y, x])
        assert (od[y, x] == depth) or (np.isnan(depth) and np.isnan(od[y, x]))
        if index != -1:
            assert ob[y, x] == min_cost
        finite += np.isfinite(depth)
    assert finite > 10
    ref.close()
