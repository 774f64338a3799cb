"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same seeded
inputs.  Bars (BASELINE.json north_star): cost volumes within 1e-4 relative (absolute floor 1e-6
on the 0..120 / -1..1 cost scales, FP32 storage of the volume adds 6e-8 relative); integer
depth-index maps bit-exact except at ties/truncation flips, whose rate is asserted < 1e-4."""
import numpy as np
import pytest

from oracle import oracle_api as O
from stereoreconstruction_b200 import capi, types as T
from scene_util import refractive_arc_scene, rectified_scene, cost_close

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def arc():
    cams, imgs, ms, surf = refractive_arc_scene(V=4, w=96, h=64, masks=True)
    return cams, imgs, ms, O.Scene(cams, imgs, ms)


@pytest.fixture(scope="module")
def ctx():
    c = capi.Context(0)
    yield c
    c.close()


def test_unproject_grid_matches_oracle(arc, ctx):
    cams, imgs, ms, sc = arc
    ctx.set_views(cams, imgs, ms)
    ctx.set_params(T.default_params(True, 420.0, 580.0, 16))
    for v in range(len(cams)):
        g = ctx.unproject_grid(v)
        o = sc.unproject_grid(v)
        assert np.allclose(g, o, rtol=0, atol=1e-11), np.abs(g - o).max()


def test_project_points_matches_oracle(arc, ctx):
    cams, imgs, ms, sc = arc
    ctx.set_views(cams, imgs, ms)
    rng = np.random.RandomState(7)
    rays = sc.unproject_grid(0)
    t = rng.uniform(380.0, 620.0, size=rays.shape[:2] + (1,))
    pts = (rays[..., :3] + t * rays[..., 3:]).reshape(-1, 3)
    for v in range(len(cams)):
        gxy, gok = ctx.project_points(v, pts)
        oxy, ook = sc.project_points(v, pts, root_mode=1)
        assert (gok == ook).all()
        assert np.abs(gxy - oxy)[ook == 1].max() < 1e-9
        # and against the reference-style quartic root selection
        qxy, qok = sc.project_points(v, pts, root_mode=0)
        assert (qok == ook).all()
        assert np.abs(gxy - qxy)[ook == 1].max() < 1e-7


@pytest.mark.parametrize("kind", [T.SR_WEIGHT_ADAPTIVE, T.SR_WEIGHT_GEODESIC])
@pytest.mark.parametrize("radius", [2, 5])
def test_weights_match_oracle(arc, ctx, kind, radius):
    cams, imgs, ms, sc = arc
    ctx.set_views(cams, imgs, ms)
    rng = np.random.RandomState(3)
    cx = np.concatenate([rng.randint(0, 96, 40), [0, 95, 0, 95, 1]]).astype(np.int32)
    cy = np.concatenate([rng.randint(0, 64, 40), [0, 0, 63, 63, 62]]).astype(np.int32)
    g = ctx.weights(1, kind, radius, cx, cy)
    o = sc.weights(1, kind, radius, cx, cy)
    assert np.allclose(g, o, rtol=1e-13, atol=1e-300), np.abs(g - o).max()


def _run_twoview(ctx, sc, P, a, b):
    P.keep_cost_volume = 1
    ctx.set_params(P)
    ctx.run_view(a, [b])
    gi, gd, gb = ctx.depth_index(a), ctx.depth(a), ctx.best_cost(a)
    gv = ctx.cost_volume(1)[0]  # [D][rows][w]
    od, oi, ob, ov = sc.twoview_label(P, a, b, root_mode=1, want_volume=True)
    return (gi, gd, gb, gv), (oi, od, ob, np.transpose(ov, (2, 0, 1)))


def _check(g, o, max_flip_rate=1e-4):
    gi, gd, gb, gv = g
    oi, od, ob, ov = o
    bad_cost = cost_close(gv, ov)
    # a cost may differ only where the integer tap flipped (projection within ~1e-9 px of an
    # integer): count, do not tolerate silently
    rate = bad_cost.mean()
    assert rate <= max_flip_rate, f"cost-volume mismatch rate {rate}"
    mism = gi != oi
    assert mism.mean() <= max_flip_rate, f"index mismatch rate {mism.mean()}"
    same = ~mism
    both_nan = np.isnan(gd) & np.isnan(od)
    assert (both_nan | (gd == od))[same].all()
    assert not cost_close(gb[same], ob[same]).any()
    return rate, mism.mean()


@pytest.mark.parametrize("weight", [T.SR_WEIGHT_ADAPTIVE, T.SR_WEIGHT_GEODESIC])
@pytest.mark.parametrize("cost", [T.SR_COST_NCC_TWOVIEW, T.SR_COST_SAD_TWOVIEW])
@pytest.mark.parametrize("radius", [2, 5])
def test_twoview_label_refractive(arc, ctx, weight, cost, radius):
    cams, imgs, ms, sc = arc
    ctx.set_views(cams, imgs, ms)
    P = T.default_params(False, 420.0, 580.0, 24, radius=radius, weight_kind=weight, cost_kind=cost)
    g, o = _run_twoview(ctx, sc, P, 1, 2)
    _check(g, o)
    assert (g[0] >= 0).mean() > 0.2  # the test is not vacuous


def test_twoview_label_rectified_r16_adaptive(ctx):
    """cfg3 in miniature: 33x33 AdaptiveWeight window, uniform-disparity labels, no mask."""
    cams, imgs, ms, surf = rectified_scene(w=112, h=40)
    sc = O.Scene(cams, imgs)
    ctx.set_views(cams, imgs, None)
    P = T.default_params(False, 100.0, 500.0, 32, radius=16, weight_kind=T.SR_WEIGHT_ADAPTIVE)
    g, o = _run_twoview(ctx, sc, P, 0, 1)
    _check(g, o)
    g, o = _run_twoview(ctx, sc, P, 1, 0)
    _check(g, o)


@pytest.mark.parametrize("weight", [T.SR_WEIGHT_ADAPTIVE, T.SR_WEIGHT_GEODESIC])
@pytest.mark.parametrize("radius", [2, 3])
def test_mvs_label_refractive(arc, ctx, weight, radius):
    cams, imgs, ms, sc = arc
    ctx.set_views(cams, imgs, ms)
    nb_g = ctx.select_neighbours(3)
    nb_o = sc.select_neighbours(3)
    assert nb_g == [[int(v) for v in r] for r in nb_o]
    P = T.default_params(True, 420.0, 580.0, 24, radius=radius, weight_kind=weight, keep_cost_volume=1)
    ctx.set_params(P)
    for ref in (0, 2):
        ctx.run_view(ref, nb_g[ref])
        gi, gd, gb = ctx.depth_index(ref), ctx.depth(ref), ctx.best_cost(ref)
        gv = ctx.cost_volume(len(nb_g[ref]))
        od, oi, ob, ov, _ = sc.mvs_view(P, ref, nb_o[ref], want_volume=True)
        ov = np.transpose(ov, (0, 3, 1, 2))
        _check((gi, gd, gb, gv), (oi, od, ob, ov))
        assert (gi >= 0).mean() > 0.1


def test_row_sharding_is_bit_invisible(arc, ctx):
    """Rows [a,b) computed alone equal the same rows of the full run."""
    cams, imgs, ms, sc = arc
    ctx.set_views(cams, imgs, ms)
    P = T.default_params(True, 420.0, 580.0, 16)
    ctx.set_params(P)
    ctx.run_view(1, [0, 2, 3])
    full_i, full_d = ctx.depth_index(1).copy(), ctx.depth(1).copy()
    ctx.set_views(cams, imgs, ms)
    for (r0, r1) in ((0, 20), (20, 41), (41, 64)):
        P.row_begin, P.row_end = r0, r1
        ctx.set_params(P)
        ctx.run_view(1, [0, 2, 3])
    assert (ctx.depth_index(1) == full_i).all()
    d = ctx.depth(1)
    assert ((d == full_d) | (np.isnan(d) & np.isnan(full_d))).all()


def test_tap_budget_banding_is_invisible(arc, monkeypatch):
    cams, imgs, ms, sc = arc
    monkeypatch.setenv("SR_TAP_BUDGET_MB", "1")
    c2 = capi.Context(0)
    c2.set_views(cams, imgs, ms)
    P = T.default_params(True, 420.0, 580.0, 32)
    c2.set_params(P)
    c2.run_view(1, [0, 2, 3])
    a = c2.depth_index(1).copy()
    c2.close()
    monkeypatch.delenv("SR_TAP_BUDGET_MB")
    c3 = capi.Context(0)
    c3.set_views(cams, imgs, ms)
    c3.set_params(P)
    c3.run_view(1, [0, 2, 3])
    assert (c3.depth_index(1) == a).all()
    c3.close()


def test_cross_check_two_view(arc, ctx):
    cams, imgs, ms, sc = arc
    two = [cams[1], cams[2]]
    sc2 = O.Scene(two, imgs[1:3], ms[1:3])
    ctx.set_views(two, imgs[1:3], ms[1:3])
    P = T.default_params(False, 420.0, 580.0, 24, radius=2)
    ctx.set_params(P)
    ctx.run_view(0, [1])
    ctx.run_view(1, [0])
    dl, dr = ctx.depth(0).copy(), ctx.depth(1).copy()
    ctx.cross_check(True, 30.0)
    gl, gr = ctx.depth(0), ctx.depth(1)
    ol, orr = sc2.crosscheck_two(P, 0, 1, dl, dr, thresh=30.0)
    for g, o in ((gl, ol), (gr, orr)):
        assert ((g == o) | (np.isnan(g) & np.isnan(o))).mean() > 1 - 1e-4
    assert np.isfinite(gl).sum() > 0 and (np.isinf(gl) & np.isfinite(dl)).sum() > 0


def test_cross_check_mvs(arc, ctx):
    cams, imgs, ms, sc = arc
    ctx.set_views(cams, imgs, ms)
    P = T.default_params(True, 420.0, 580.0, 24)
    ctx.set_params(P)
    nb = ctx.select_neighbours(3)
    for v in range(4):
        ctx.run_view(v, nb[v])
    before = [ctx.depth(v).copy() for v in range(4)]
    ctx.cross_check(False, 12.0)
    after = [ctx.depth(v) for v in range(4)]
    want = sc.crosscheck_mvs(P, before, 12.0)
    for g, o in zip(after, want):
        assert ((g == o) | (np.isnan(g) & np.isnan(o))).mean() > 1 - 1e-4
    assert sum(np.isfinite(a).sum() for a in after) > 0


@pytest.mark.parametrize("weight", [T.SR_WEIGHT_ADAPTIVE, T.SR_WEIGHT_GEODESIC])
@pytest.mark.parametrize("radius", [2, 3, 5])
def test_mvs_screen_path_matches_oracle(arc, ctx, weight, radius):
    """The production MVS path (no kept volume): FP32 screen + FP64 verify (sr_match_screen.cuh).
    The verify step is the reference's exact tap filter, so index, depth AND winning cost must
    equal the oracle's (cost to FP64 rounding of the geodesic/adaptive exp(), 1e-12)."""
    cams, imgs, ms, sc = arc
    ctx.set_views(cams, imgs, ms)
    nb = ctx.select_neighbours(3)
    P = T.default_params(True, 420.0, 580.0, 48, radius=radius, weight_kind=weight)
    ctx.set_params(P)
    for ref in (1, 3):
        ctx.run_view(ref, nb[ref])
        gi, gd, gb = ctx.depth_index(ref), ctx.depth(ref), ctx.best_cost(ref)
        od, oi, ob, _, _ = sc.mvs_view(P, ref, nb[ref])
        mism = gi != oi
        assert mism.mean() <= 1e-4, f"index mismatch rate {mism.mean()}"
        same = ~mism
        assert ((gd == od) | (np.isnan(gd) & np.isnan(od)))[same].all()
        lab = same & (gi >= 0)
        assert lab.mean() > 0.1
        assert np.abs(gb[lab] - ob[lab]).max() <= 1e-12


def test_mvs_screen_equals_all_fp64_kernel(monkeypatch):
    """A/B on a larger unmasked scene: the screened path and the all-FP64 match_kernel pick the
    same label at every pixel (the screen only prunes labels that cannot win)."""
    cams, imgs, ms, surf = refractive_arc_scene(V=4, w=320, h=200, arc_deg=25.0, cell=6.0)
    P = T.default_params(True, 420.0, 580.0, 96)
    out = []
    for flag in ("0", "1"):
        monkeypatch.setenv("SR_MATCH_SCREEN", flag)
        c = capi.Context(0)
        c.set_views(cams, imgs, None)
        c.set_params(P)
        nb = c.select_neighbours(3)
        c.run_view(1, nb[1])
        out.append((c.depth_index(1).copy(), c.best_cost(1).copy()))
        c.close()
    (i0, b0), (i1, b1) = out
    assert (i0 >= 0).mean() > 0.3
    assert (i0 == i1).all(), f"{(i0 != i1).sum()} pixels differ"
    lab = i0 >= 0
    assert np.abs(b0[lab] - b1[lab]).max() <= 1e-9


def test_build_interpolation_self_check(monkeypatch):
    """The refractive build projects every 4th label exactly and reads the labels in between off a
    cubic through the anchors, guarded by the distance to the nearest pixel boundary.  With the
    self-check on, every interpolated label is also projected exactly: no tap may differ."""
    monkeypatch.setenv("SR_MATCH_STATS", "1")
    cams, imgs, ms, surf = refractive_arc_scene(V=4, w=320, h=200, arc_deg=25.0, cell=6.0)
    c = capi.Context(0)
    c.set_views(cams, imgs, None)
    for D in (96, 37, 7):  # also label counts that are not multiples of the anchor stride
        c.set_params(T.default_params(True, 420.0, 580.0, D))
        nb = c.select_neighbours(3)
        c.run_view(1, nb[1])
    st = c.build_stats()
    c.close()
    assert st["interpolated"] > 1e6
    assert st["tap_mismatches"] == 0
    assert st["guard_fallbacks"] < 0.01 * st["interpolated"]


def test_calibration_residuals_match_oracle(arc, ctx):
    """RefractiveCalibrationFunction::diff (stereo/refractioncalibration.cpp:175-201) on the GPU."""
    cams, imgs, ms, sc = arc
    rng = np.random.RandomState(21)
    n = 5000
    pairs = np.stack([rng.randint(0, 4, n), rng.randint(0, 4, n)], axis=1)
    pairs[pairs[:, 0] == pairs[:, 1], 1] = (pairs[pairs[:, 0] == pairs[:, 1], 0] + 1) % 4
    # true correspondences (projections of points on the scene surface) plus pixel noise
    rays = sc.unproject_grid(0)
    pts = (rays[..., :3] + rng.uniform(450, 550, rays.shape[:2] + (1,)) * rays[..., 3:]).reshape(-1, 3)
    pts = pts[rng.randint(0, pts.shape[0], n)]
    pix = np.empty((n, 4))
    for v in range(4):
        xy, ok = sc.project_points(v, pts)
        for side in (0, 1):
            m = pairs[:, side] == v
            pix[m, 2 * side:2 * side + 2] = xy[m]
    pix += rng.normal(scale=0.3, size=pix.shape)
    g = ctx.calibration_residuals(cams, pairs, pix)
    o = O.calibration_residuals(cams, pairs, pix)
    fin = np.isfinite(o)
    assert fin.mean() > 0.9 and (np.isfinite(g) == fin).all()
    assert np.allclose(g[fin], o[fin], rtol=1e-10, atol=1e-12)
    assert np.median(o[fin]) < 2.0  # sub-pixel noise gives pixel-scale residuals: the metric is what it claims


@pytest.mark.parametrize("radius", [2, 3])
def test_mvs_peak_lists_match_oracle(arc, ctx, radius):
    """CostFunction::peakPairs (multiviewstereo.cpp:479-482,589-602): the 9 largest (ncc, depth) pairs."""
    cams, imgs, ms, sc = arc
    ctx.set_views(cams, imgs, ms)
    nb = ctx.select_neighbours(3)
    P = T.default_params(True, 420.0, 580.0, 40, radius=radius, keep_cost_volume=2)
    ctx.set_params(P)
    ctx.run_view(1, nb[1])
    gp = ctx.peaks(1)
    gi = ctx.depth_index(1)
    od, oi, ob, _, op = sc.mvs_view(P, 1, nb[1], want_peaks=True)
    assert (gi != oi).mean() <= 1e-4
    white = ms[1] == 255
    assert (gp[..., 1][white] == op[..., 1][white]).mean() > 1 - 1e-3  # depths of the kept pairs
    assert not cost_close(gp[..., 0][white], op[..., 0][white]).any() or \
        cost_close(gp[..., 0][white], op[..., 0][white]).mean() <= 1e-3
    assert (gp[..., 0][white][:, -1] > 0.95).mean() > 0.1  # the last entry is the winner
    P.keep_cost_volume = 0
    ctx.set_params(P)
    with pytest.raises(capi.SrError):
        ctx.run_view(1, nb[1])
        ctx.peaks(1)
