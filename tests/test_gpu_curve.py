"""Curve-mode search on the GPU (the reference's LIVE formulation, stereo/multiviewstereo.cpp:574-602,
754-810 and stereo/twoviewstereo.cpp:285-305,999-1054) against the oracle's curve mode."""
import numpy as np
import pytest

from oracle import oracle_api as O
from stereoreconstruction_b200 import capi, types as T
from scene_util import refractive_arc_scene, cost_close

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = capi.Context(0)
    yield c
    c.close()


def _depth_equal(g, o):
    with np.errstate(invalid="ignore"):
        return (g == o) | (np.isnan(g) & np.isnan(o)) | (np.abs(g - o) <= 1e-9 * np.abs(o))


@pytest.mark.parametrize("interface,distortion", [(True, True), (False, True), (False, False)])
@pytest.mark.parametrize("weight", [T.SR_WEIGHT_GEODESIC, T.SR_WEIGHT_ADAPTIVE])
def test_mvs_curve_matches_oracle(ctx, interface, distortion, weight):
    cams, imgs, ms, _ = refractive_arc_scene(V=4, w=96, h=64, masks=True, interface=interface, distortion=distortion)
    ctx.set_views(cams, imgs, ms)
    P = T.default_params(True, 420.0, 580.0, 40, weight_kind=weight)
    ctx.set_params(P)
    nb = ctx.select_neighbours(3)
    sc = O.Scene(cams, imgs, ms)
    for ref in (0, 2):
        ctx.run_view_curve(ref, nb[ref])
        gd, gb = ctx.depth(ref), ctx.best_cost(ref)
        od, _, ob, _, _ = sc.mvs_view(P, ref, nb[ref], curve_mode=True)
        ok = _depth_equal(gd, od)
        assert ok.mean() >= 1 - 1e-4, f"depth mismatch rate {1 - ok.mean()}"
        have = ok & (od > 0) & np.isfinite(od)
        assert have.mean() > 0.1
        assert np.abs(gb[have] - ob[have]).max() <= 1e-12
        assert (ctx.depth_index(ref)[have] >= 0).all()


@pytest.mark.parametrize("radius,cost", [(2, T.SR_COST_NCC_TWOVIEW), (5, T.SR_COST_NCC_TWOVIEW), (2, T.SR_COST_SAD_TWOVIEW)])
def test_twoview_curve_matches_oracle(ctx, radius, cost):
    cams, imgs, ms, _ = refractive_arc_scene(V=4, w=96, h=64, masks=True)
    ctx.set_views(cams, imgs, ms)
    P = T.default_params(False, 420.0, 580.0, 32, radius=radius, cost_kind=cost)
    ctx.set_params(P)
    sc = O.Scene(cams, imgs, ms)
    for (a, b) in ((1, 2), (2, 1)):
        ctx.run_view_curve(a, [b])
        gd, gb = ctx.depth(a), ctx.best_cost(a)
        od, ob, cnt = sc.twoview_curve(P, a, b)
        ok = _depth_equal(gd, od)
        assert ok.mean() >= 1 - 1e-4, f"depth mismatch rate {1 - ok.mean()}"
        fin = ok & np.isfinite(od)
        assert fin.mean() > 0.05 and cnt.max() > 10
        assert not cost_close(gb[fin], ob[fin]).any()


def test_curve_banding_is_invisible(monkeypatch):
    cams, imgs, ms, _ = refractive_arc_scene(V=4, w=96, h=64, masks=True)
    P = T.default_params(True, 420.0, 580.0, 40)
    outs = []
    for budget in (None, "1"):
        if budget:
            monkeypatch.setenv("SR_TAP_BUDGET_MB", budget)
        c = capi.Context(0)
        c.set_views(cams, imgs, ms)
        c.set_params(P)
        c.run_view_curve(1, [0, 2, 3])
        outs.append((c.depth(1).copy(), c.best_cost(1).copy()))
        c.close()
    (d0, b0), (d1, b1) = outs
    assert ((d0 == d1) | (np.isnan(d0) & np.isnan(d1))).all()
    assert ((b0 == b1) | (np.isnan(b0) & np.isnan(b1))).all()


@pytest.mark.parametrize("interface", [True, False])
def test_mvs_curve_peak_lists_match_oracle(ctx, interface):
    """CostFunction::peakPairs in the reference's live (curve) formulation (multiviewstereo.cpp:479-482,
    583-602): the 9 largest (ncc, z) pairs over the curve pixels of all neighbours, ascending, padded
    with (0, -1); the last pair is the pixel's result.  keep_cost_volume & 2 routes the view through
    the all-FP64 kernel, which must also agree with the screened kernel's winner."""
    cams, imgs, ms, _ = refractive_arc_scene(V=4, w=96, h=64, masks=True, interface=interface)
    ctx.set_views(cams, imgs, ms)
    P = T.default_params(True, 420.0, 580.0, 40)
    ctx.set_params(P)
    nb = ctx.select_neighbours(3)
    ctx.run_view_curve(1, nb[1])
    sd, sb = ctx.depth(1).copy(), ctx.best_cost(1).copy()  # screened kernel
    P.keep_cost_volume = 2
    ctx.set_params(P)
    ctx.run_view_curve(1, nb[1])
    gp, gd, gb = ctx.peaks(1), ctx.depth(1), ctx.best_cost(1)
    # winner of the all-FP64 kernel == winner of the screened kernel (costs of the streaming FP64 form
    # and of the exact two-pass form differ at the 1e-13 level, so an exact tie may flip: rate bound)
    assert _depth_equal(gd, sd).mean() >= 1 - 1e-3
    same = _depth_equal(gd, sd) & np.isfinite(sb) & (sd > 0)
    assert np.abs(gb[same] - sb[same]).max() <= 1e-9
    sc = O.Scene(cams, imgs, ms)
    od, _, ob, _, op = sc.mvs_view(P, 1, nb[1], curve_mode=True, want_peaks=True)
    white = ms[1] == 255
    gz, oz = gp[..., 1][white], op[..., 1][white]
    assert _depth_equal(gz, oz).mean() >= 1 - 1e-3                    # depths of the kept pairs
    assert cost_close(gp[..., 0][white], op[..., 0][white], rel=1e-9, abs_floor=1e-12).mean() <= 1e-3
    assert (gp[..., 0][white][:, -1] > 0.95).mean() > 0.1            # the last entry is the winner
    assert (np.diff(gp[..., 0][white], axis=-1) >= 0).all()          # ascending
    filled = (gp[..., 0][white] > 0).sum(axis=-1)
    assert filled.max() == 9 and (filled == 0).any()                 # full lists and empty lists both occur
    assert (filled == (op[..., 0][white] > 0).sum(axis=-1)).mean() >= 1 - 1e-3
    lab = white & (gd > 0) & np.isfinite(gd)
    assert (gp[..., -1, 1][lab] == gd[lab]).all() and (gp[..., -1, 0][lab] == gb[lab]).all()


def test_curve_mode_rejects_cost_volume(ctx):
    cams, imgs, ms, _ = refractive_arc_scene(V=3, w=32, h=24, masks=False)
    ctx.set_views(cams, imgs, None)
    ctx.set_params(T.default_params(True, 420.0, 580.0, 8, keep_cost_volume=1))
    with pytest.raises(capi.SrError):
        ctx.run_view_curve(0, [1, 2])
