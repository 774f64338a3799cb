"""Seeded random configurations (camera arc, interface tilt / distance / index, lens distortion,
masks, label count, window radius, weight functor, depth range) on the GPU against the oracle, for
every mode: label-mode MVS (screened), label-mode two-view, curve-mode MVS and two-view."""
import numpy as np
import pytest

from oracle import oracle_api as O
from stereoreconstruction_b200 import capi, scenes, types as T
from scene_util import cost_close

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = capi.Context(0)
    yield c
    c.close()


def random_case(seed):
    rng = np.random.RandomState(seed)
    w, h = int(rng.choice([48, 64, 80])), int(rng.choice([36, 48]))
    V = int(rng.choice([3, 4, 5]))
    refr = rng.rand() < 0.7
    dist = None
    if rng.rand() < 0.7:
        dist = (rng.uniform(-0.2, 0.1), rng.uniform(-0.1, 0.1), rng.uniform(-0.003, 0.003), rng.uniform(-0.003, 0.003),
                rng.uniform(-0.05, 0.05))
    cams = scenes.arc_cameras(V, w, h, arc_deg=rng.uniform(10, 40), radius=rng.uniform(400, 600), distortion=dist,
                              interface=refr, tilt_px=(rng.uniform(-150, 150), rng.uniform(-100, 100)),
                              plane_d=rng.uniform(5, 60), n=rng.choice([1.333, 1.5, 0.75]))
    dummy = [np.zeros((h, w, 4), np.uint8)] * V
    sc0 = O.Scene(cams, dummy)
    surf = scenes.HeightField(z0=0.0, amp=rng.uniform(5, 25), lx=rng.uniform(40, 90), ly=rng.uniform(30, 70))
    imgs = scenes.render_views(V, lambda v: sc0.unproject_grid(v), surf, seed=seed, cell=rng.uniform(6, 20))
    masks = None
    if rng.rand() < 0.6:
        masks = [np.where(rng.rand(h, w) > 0.03, 255, 0).astype(np.uint8) for _ in range(V)]
        yy, xx = np.mgrid[0:h, 0:w]
        for m in masks:
            m[((xx - w * rng.uniform(0.3, 0.7)) / (0.5 * w)) ** 2 + ((yy - h * rng.uniform(0.3, 0.7)) / (0.5 * h)) ** 2 > 1] = 0
    lo, hi = 500 - rng.uniform(40, 120), 500 + rng.uniform(40, 120)
    if rng.rand() < 0.2:
        lo, hi = hi, lo
    kw = dict(radius=int(rng.choice([1, 2, 3, 4, 5])), weight_kind=int(rng.choice([0, 1])))
    D = int(rng.choice([7, 16, 33, 50]))
    return cams, imgs, masks, (lo, hi, D), kw, rng


@pytest.mark.parametrize("seed", range(100, 112))
def test_random_configuration(ctx, seed):
    cams, imgs, masks, (lo, hi, D), kw, rng = random_case(seed)
    V = len(cams)
    sc = O.Scene(cams, imgs, masks)
    ctx.set_views(cams, imgs, masks)
    ref = int(rng.randint(V))
    nbrs = [v for v in range(V) if v != ref][:3]

    # label-mode MVS (screened) and curve-mode MVS
    P = T.default_params(True, lo, hi, D, **kw)
    ctx.set_params(P)
    ctx.run_view(ref, nbrs)
    gi, gd, gb = ctx.depth_index(ref), ctx.depth(ref), ctx.best_cost(ref)
    od, oi, ob, _, _ = sc.mvs_view(P, ref, nbrs)
    mism = gi != oi
    assert mism.mean() <= 2e-4, (seed, mism.mean())
    assert ((gd == od) | (np.isnan(gd) & np.isnan(od)))[~mism].all()
    lab = ~mism & (oi >= 0)
    if lab.any():
        assert np.abs(gb[lab] - ob[lab]).max() <= 1e-12
    ctx.run_view_curve(ref, nbrs)
    gd = ctx.depth(ref)
    od, _, ob, _, _ = sc.mvs_view(P, ref, nbrs, curve_mode=True)
    with np.errstate(invalid="ignore"):
        ok = (gd == od) | (np.isnan(gd) & np.isnan(od)) | (np.abs(gd - od) <= 1e-9 * np.abs(od))
    assert ok.mean() >= 1 - 2e-4, (seed, 1 - ok.mean())

    # label-mode and curve-mode two-view
    cost = int(rng.choice([T.SR_COST_NCC_TWOVIEW, T.SR_COST_SAD_TWOVIEW]))
    P2 = T.default_params(False, lo, hi, D, cost_kind=cost, **kw)
    ctx.set_params(P2)
    b = nbrs[0]
    ctx.run_view(ref, [b])
    gi, gd, gb = ctx.depth_index(ref), ctx.depth(ref), ctx.best_cost(ref)
    od, oi, ob, _ = sc.twoview_label(P2, ref, b, root_mode=1)
    mism = gi != oi
    assert mism.mean() <= 2e-4, (seed, mism.mean())
    assert ((gd == od) | (np.isnan(gd) & np.isnan(od)))[~mism].all()
    lab = ~mism & np.isfinite(ob)
    assert not cost_close(gb[lab], ob[lab]).any()
    ctx.run_view_curve(ref, [b])
    gd = ctx.depth(ref)
    od, ob, cnt = sc.twoview_curve(P2, ref, b)
    with np.errstate(invalid="ignore"):
        ok = (gd == od) | (np.isnan(gd) & np.isnan(od)) | (np.abs(gd - od) <= 1e-9 * np.abs(od))
    assert ok.mean() >= 1 - 2e-4, (seed, 1 - ok.mean())
