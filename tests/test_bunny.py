"""BASELINE.json configs[0] and [1]: the reference's own `bunny` example (real calibrated,
lens-distorted cameras; object masks from the PNG alpha), reduced to 1/4 size (tests/golden/bunny),
with and without the injected refractive interface.  CPU: the fixture and the oracle on it.
GPU: two-view (cfg1) and 8-camera multi-view (cfg2) against the oracle."""
import numpy as np
import pytest

from oracle import oracle_api as O
from stereoreconstruction_b200 import types as T
from bunny_util import load
from scene_util import cost_close


def test_bunny_fixture_and_cameras():
    cams, imgs, masks, scale = load()
    assert len(cams) == 8 and imgs[0].shape == (192, 256, 4) and scale == 0.25
    f = np.array([c.K[0] for c in cams])
    assert (f > 1700).all() and (f < 1900).all()  # f ~ 1756-1820 px
    assert all(c.is_distorted and not c.is_refractive for c in cams)
    C = np.array([list(c.C) for c in cams])
    d01 = np.linalg.norm(C[0] - C[1])
    assert 15 < d01 < 25  # 7310085 - 7310087: 19.0 units
    frac = np.mean([(m == 255).mean() for m in masks])
    assert 0.2 < frac < 0.5  # ~31-41 % of the pixels are object
    rcams, _, _, _ = load(refractive=True)
    assert all(c.is_refractive for c in rcams)
    # K R [I | -C] reproduces the XML's projection matrix up to the scale Camera::setP divides by
    for c in cams:
        K, R, t = np.array(c.K).reshape(3, 3), np.array(c.R).reshape(3, 3), np.array(c.t)
        assert np.allclose(R @ R.T, np.eye(3), atol=1e-12) and abs(np.linalg.det(R) - 1) < 1e-9
        assert np.allclose(-R.T @ t, np.array(c.C), atol=1e-9)


def test_oracle_reconstructs_the_bunny():
    """Label-mode MVS of one view (100 levels over the depth range that brackets the object for the cameras of
    example/project.xml, see golden_cases.py): nearly all of the object gets a depth, and the depths lie where
    the object is (the cameras sit on a semi-circle of radius ~42 around it)."""
    cams, imgs, masks, scale = load()
    sc = O.Scene(cams, imgs, masks)
    nb = sc.select_neighbours(3)
    P = T.default_params(True, 30.0, 55.0, 100, image_scale=scale)
    od, oi, ob, _, _ = sc.mvs_view(P, 0, nb[0])
    obj = masks[0] == 255
    have = obj & (oi >= 0)
    assert have.sum() > 0.9 * obj.sum()
    assert 35 <= np.median(od[have]) <= 50
    assert np.isinf(od[~obj]).all()


@pytest.fixture(scope="module")
def gpu_ctx():
    from stereoreconstruction_b200 import capi
    c = capi.Context(0)
    yield c
    c.close()


def _same_maps(gi, gd, gb, oi, od, ob, best_rel):
    mism = gi != oi
    assert mism.mean() <= 1e-4, f"index mismatch rate {mism.mean()}"
    same = ~mism
    assert ((gd == od) | (np.isnan(gd) & np.isnan(od)))[same].all()
    lab = same & (oi >= 0)
    if lab.any():
        assert not cost_close(gb[lab], ob[lab], rel=best_rel, abs_floor=best_rel).any()
    return int(lab.sum())


@pytest.mark.gpu
@pytest.mark.parametrize("refractive", [False, True])
def test_cfg1_bunny_two_view(gpu_ctx, refractive):
    """cfg1: cameras 7310085 + 7310087, r = 5 GeodesicWeight NCC, both directions + cross-check."""
    cams, imgs, masks, scale = load(refractive)
    two, im2, ms2 = cams[:2], imgs[:2], masks[:2]
    sc = O.Scene(two, im2, ms2)
    gpu_ctx.set_views(two, im2, ms2)
    P = T.default_params(False, 30.0, 55.0, 100, image_scale=scale)
    gpu_ctx.set_params(P)
    depths = []
    for (a, b) in ((0, 1), (1, 0)):
        gpu_ctx.run_view(a, [b])
        od, oi, ob, _ = sc.twoview_label(P, a, b, root_mode=1)
        assert _same_maps(gpu_ctx.depth_index(a), gpu_ctx.depth(a), gpu_ctx.best_cost(a), oi, od, ob, 1e-4) > 1000
        depths.append(od)
    gpu_ctx.cross_check(True, 5.0)
    ol, orr = sc.crosscheck_two(P, 0, 1, depths[0], depths[1], thresh=5.0)
    for v, o in ((0, ol), (1, orr)):
        g = gpu_ctx.depth(v)
        assert ((g == o) | (np.isnan(g) & np.isnan(o))).mean() > 1 - 1e-4


@pytest.mark.gpu
@pytest.mark.parametrize("weight", [T.SR_WEIGHT_ADAPTIVE, T.SR_WEIGHT_GEODESIC])
def test_cfg2_bunny_multi_view(gpu_ctx, weight):
    """cfg2: all 8 cameras, the reference's neighbour rule, r = 2, all reference views, cross-check 5,
    refractive interface injected; label mode and the reference's live curve mode."""
    cams, imgs, masks, scale = load(refractive=True)
    sc = O.Scene(cams, imgs, masks)
    gpu_ctx.set_views(cams, imgs, masks)
    P = T.default_params(True, 30.0, 55.0, 100, image_scale=scale, weight_kind=weight)
    gpu_ctx.set_params(P)
    nb = gpu_ctx.select_neighbours(3)
    assert nb == [[int(v) for v in r] for r in sc.select_neighbours(3)]
    before, labelled = [], 0
    for v in range(8):
        gpu_ctx.run_view(v, nb[v])
        od, oi, ob, _, _ = sc.mvs_view(P, v, nb[v])
        labelled += _same_maps(gpu_ctx.depth_index(v), gpu_ctx.depth(v), gpu_ctx.best_cost(v), oi, od, ob, 1e-12)
        before.append(od)
    assert labelled > 5000  # some views see little of the object through the injected interface
    gpu_ctx.cross_check(False, 5.0)
    want = sc.crosscheck_mvs(P, before, 5.0)
    for v in range(8):
        g = gpu_ctx.depth(v)
        assert ((g == want[v]) | (np.isnan(g) & np.isnan(want[v]))).mean() > 1 - 1e-4
    if weight == T.SR_WEIGHT_GEODESIC:  # the reference's default functor: also its live search
        for v in (0, 5):
            gpu_ctx.run_view_curve(v, nb[v])
            od, _, ob, _, _ = sc.mvs_view(P, v, nb[v], curve_mode=True)
            gd = gpu_ctx.depth(v)
            with np.errstate(invalid="ignore"):
                ok = (gd == od) | (np.isnan(gd) & np.isnan(od)) | (np.abs(gd - od) <= 1e-9 * np.abs(od))
            assert ok.mean() >= 1 - 1e-4
