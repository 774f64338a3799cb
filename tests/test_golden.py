"""Golden-vector tests (tests/golden/*.npz, written by tests/golden/make_golden.py).

CPU (`-m "not gpu"`): the oracle reproduces the outputs of the reference's own compiled sources
(ref_leaves.npz) and its own committed scene outputs bit for bit.
GPU (`-m gpu`): the CUDA path, through the C ABI, reproduces the committed scene outputs."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import oracle_api as O
from stereoreconstruction_b200 import types as T
import golden_cases as G
from scene_util import cost_close

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def leaves():
    return np.load(os.path.join(GOLD, "ref_leaves.npz"))


@pytest.fixture(scope="module")
def scenes_gold():
    return np.load(os.path.join(GOLD, "oracle_scenes.npz"))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _same(a, b):
    a, b = np.asarray(a), np.asarray(b)
    if a.dtype.kind == "f":
        return bool(((a == b) | (np.isnan(a) & np.isnan(b))).all())
    return bool((a == b).all())


# ------------------------------------------------------------------ oracle vs reference leaves
def test_oracle_lines_match_reference_golden(leaves):
    L = O.lib()
    cases, w, h = G.line_cases()
    counts, pts, off, k = leaves["line_counts"], leaves["line_points"], 0, 0
    for i, (x0, y0, x1, y1) in enumerate(cases):
        for clip in (0, 1):
            buf = np.empty((4096, 2), np.int32)
            n = L.orc_line(x0, y0, x1, y1, clip, w, h, _ip(buf), 4096)
            assert n == counts[k], (x0, y0, x1, y1, clip)
            assert (buf[:n] == pts[off:off + n]).all(), (x0, y0, x1, y1, clip)
            off += n
            k += 1
        a = np.array([x0, y0, x1, y1], np.int32)
        ok = L.orc_clip_line(_ip(a), w, h)
        assert ok == leaves["clip_ok"][i]
        if ok:
            assert (a == leaves["clip_out"][i]).all()


def test_oracle_ray_ops_match_reference_golden(leaves):
    L = O.lib()
    for i, (s1, d1, s2, d2, pn, pd, n) in enumerate(G.ray_cases()):
        a = np.full(3, np.nan)
        ok = L.orc_intersect(_dp(s1), _dp(d1), _dp(pn), C.c_double(pd), _dp(a))
        assert ok == leaves["intersect_ok"][i]
        if ok:
            assert _same(a, leaves["intersect"][i])
        b = np.empty(6)
        ok = L.orc_refract(_dp(s1), _dp(d1), _dp(pn), C.c_double(pd), C.c_double(n), _dp(b))
        assert ok == leaves["refract_ok"][i] and _same(b, leaves["refract"][i])
        c = np.empty(6)
        L.orc_closest_points(_dp(s1), _dp(d1), _dp(s2), _dp(d2), _dp(c))
        assert _same(c, leaves["closest"][i])


def test_oracle_sample_and_weights_match_reference_golden(leaves):
    img = G.leaf_image()
    h, w = img.shape[:2]
    sc = O.Scene([T.make_camera(np.eye(3), np.eye(3), np.zeros(3))], [img])
    L = O.lib()
    a = np.empty(4)
    for i, (x, y) in enumerate(G.sample_points(w, h)):
        L.orc_sample(sc.ptr, 0, C.c_double(x), C.c_double(y), _dp(a))
        assert _same(a, leaves["sample"][i]), (x, y)
    cx, cy = G.weight_centres(w, h)
    for kind in (0, 1):
        for radius in G.WEIGHT_RADII:
            mine = np.asarray(sc.weights(0, kind, radius, cx, cy)).reshape(cx.size, -1)
            assert _same(mine, leaves[f"weights_k{kind}_r{radius}"]), (kind, radius)


# ------------------------------------------------------------------ oracle vs its committed scene outputs
def test_oracle_reproduces_scene_golden(scenes_gold):
    cams, imgs, ms = G.arc_scene()
    sc = O.Scene(cams, imgs, ms)
    nb = sc.select_neighbours(3)
    assert _same(np.array([[int(v) for v in r] for r in nb], np.int32), scenes_gold["arc_neighbours"])
    name, P = "arc_mvs_geo_r2", G.arc_mvs_params()["arc_mvs_geo_r2"]
    depths = []
    for ref in range(len(cams)):
        od, oi, ob, _, _ = sc.mvs_view(P, ref, nb[ref])
        assert _same(oi, scenes_gold[f"{name}_v{ref}_index"])
        assert _same(od, scenes_gold[f"{name}_v{ref}_depth"])
        assert _same(ob, scenes_gold[f"{name}_v{ref}_best"])
        depths.append(od)
    for v, d in enumerate(sc.crosscheck_mvs(P, depths, G.ARC_CROSS_THRESH)):
        assert _same(d, scenes_gold[f"{name}_v{v}_crosschecked"])
    name = "arc_two_sad_ada_r2"
    P, a, b = G.arc_twoview_params()[name]
    od, oi, ob, ov = sc.twoview_label(P, a, b, root_mode=1, want_volume=True)
    assert _same(oi, scenes_gold[f"{name}_index"]) and _same(od, scenes_gold[f"{name}_depth"])
    assert _same(np.transpose(ov, (2, 0, 1)).astype(np.float32), scenes_gold[f"{name}_volume"])


# ------------------------------------------------------------------ CUDA path vs the committed scene outputs
@pytest.fixture(scope="module")
def gpu_ctx():
    from stereoreconstruction_b200 import capi
    c = capi.Context(0)
    yield c
    c.close()


def _check_maps(gi, gd, gb, oi, od, ob, best_tol):
    mism = gi != oi
    assert mism.mean() <= 1e-4, f"index mismatch rate {mism.mean()}"
    same = ~mism
    assert ((gd == od) | (np.isnan(gd) & np.isnan(od)))[same].all()
    lab = same & (oi >= 0)
    assert lab.any()
    assert not cost_close(gb[lab], ob[lab], rel=best_tol, abs_floor=best_tol).any()


# ------------------------------------------------------------------ the reference's MultiViewStereo, end to end
@pytest.fixture(scope="module")
def ref_mvs_gold():
    return np.load(os.path.join(GOLD, "ref_mvs.npz"))


def _depth_close(a, b, rel):
    """NaN == NaN, inf == inf, -1 == -1, otherwise |a - b| <= rel * |b|."""
    with np.errstate(invalid="ignore"):
        return (a == b) | (np.isnan(a) & np.isnan(b)) | (np.abs(a - b) <= rel * np.abs(b))


@pytest.mark.parametrize("name", list(G.REF_MVS_CASES))
def test_oracle_matches_reference_mvs_golden(ref_mvs_gold, name):
    """The oracle against the committed END-TO-END outputs of the reference's own MultiViewStereo class
    (tests/golden/make_golden.py: ref_mvs) — the arc scene bit for bit; the bunny cases to the last-ulp
    differences of the two cameras whose Gram-Schmidt re-orthonormalisation (Camera::set) does not settle."""
    g = ref_mvs_gold
    mind, maxd, D, cross = G.REF_MVS_CASES[name]
    _, imgs, ms, scale = G.ref_mvs_inputs(name)
    cams = G.cams_from_bytes(g[f"{name}_cams"])
    sc = O.Scene(cams, imgs, ms)
    P = T.default_params(True, mind, maxd, D, image_scale=scale,
                         weight_kind=T.SR_WEIGHT_ADAPTIVE if G.ref_mvs_adaptive(name) else T.SR_WEIGHT_GEODESIC)
    nb = [[int(v) for v in r if v >= 0] for r in g[f"{name}_neighbours"]]
    assert nb == [[int(v) for v in r] for r in sc.select_neighbours(3)]
    rel = 0.0 if name.startswith("arc") else 1e-12
    views = [int(v) for v in g[f"{name}_views"]]
    found = 0
    for k, v in enumerate(views):
        od, _, _, _, op = sc.mvs_view(P, v, nb[v], curve_mode=True, root_mode=0, want_peaks=(v == 1))
        found += int((np.isfinite(od) & (od > 0)).sum())
        ok = _depth_close(od, g[f"{name}_before"][k], rel)
        assert ok.mean() >= 1 - 1e-4, (v, 1 - ok.mean())
        if v == 1:
            r0, r1 = G.ref_mvs_peak_rows(name, od.shape[0])
            white = ms[v][r0:r1] == 255
            assert _depth_close(op[r0:r1][white], g[f"{name}_peaks_v1"][white], rel).mean() >= 1 - 1e-4
    assert found > 1000  # not vacuous: the reference found depths
    # the cross-check of the reference's own depths: bit for bit in every case
    after = G.ref_mvs_after(g, name, None)
    if after is not None:
        want = sc.crosscheck_mvs(P, [g[f"{name}_before"][v].copy() for v in range(len(cams))], cross)
        for v in range(len(cams)):
            assert _same(want[v], after[v])


@pytest.mark.gpu
@pytest.mark.parametrize("name", [n for n in G.REF_MVS_CASES if n != "bunny_refr_ada"])  # (that one: CPU replay above)
def test_gpu_curve_mode_matches_reference_mvs_golden(ref_mvs_gold, gpu_ctx, name):
    """The CUDA path (sr_run_view_curve for every view, then sr_cross_check) against the END-TO-END
    outputs of the reference's own MultiViewStereo::runTask."""
    g = ref_mvs_gold
    mind, maxd, D, cross = G.REF_MVS_CASES[name]
    _, imgs, ms, scale = G.ref_mvs_inputs(name)
    cams = G.cams_from_bytes(g[f"{name}_cams"])
    gpu_ctx.set_views(cams, imgs, ms)
    P = T.default_params(True, mind, maxd, D, image_scale=scale,
                         weight_kind=T.SR_WEIGHT_ADAPTIVE if G.ref_mvs_adaptive(name) else T.SR_WEIGHT_GEODESIC)
    gpu_ctx.set_params(P)
    nb = gpu_ctx.select_neighbours(3)
    assert nb == [[int(v) for v in r if v >= 0] for r in g[f"{name}_neighbours"]]
    after = G.ref_mvs_after(g, name, None)
    for v in range(len(cams)):
        gpu_ctx.run_view_curve(v, nb[v])
        ok = _depth_close(gpu_ctx.depth(v), g[f"{name}_before"][v], 1e-9)
        assert ok.mean() >= 1 - 1e-4, (v, 1 - ok.mean())
    gpu_ctx.cross_check(False, cross)
    for v in range(len(cams)):
        ok = _depth_close(gpu_ctx.depth(v), after[v], 1e-9)
        assert ok.mean() >= 1 - 2e-4, (v, 1 - ok.mean())


@pytest.fixture(scope="module")
def ref_two_gold():
    return np.load(os.path.join(GOLD, "ref_two.npz"))


@pytest.mark.parametrize("name", list(G.REF_TWO_CASES))
def test_oracle_matches_reference_twoview_golden(ref_two_gold, name):
    """The oracle against the committed END-TO-END outputs of the reference's own TwoViewStereo class
    (both directions of the live curve search before and after the cross-check)."""
    g = ref_two_gold
    a, b, mind, maxd, D = G.REF_TWO_CASES[name]
    _, imgs, ms, scale = G.ref_mvs_inputs(name)
    cams = G.cams_from_bytes(g[f"{name}_cams"])
    sc = O.Scene(cams, [imgs[a], imgs[b]], [ms[a], ms[b]])
    P = T.default_params(False, mind, maxd, D, image_scale=scale)
    rel = 0.0 if name == "arc" else 1e-12
    ol, _, _ = sc.twoview_curve(P, 0, 1, root_mode=0)
    orr, _, _ = sc.twoview_curve(P, 1, 0, root_mode=0)
    assert _depth_close(ol, g[f"{name}_before"][0], rel).mean() >= 1 - 1e-4
    assert _depth_close(orr, g[f"{name}_before"][1], rel).mean() >= 1 - 1e-4
    wl, wr = sc.crosscheck_two(P, 0, 1, g[f"{name}_before"][0], g[f"{name}_before"][1], 1.0, root_mode=0)
    assert _same(wl, g[f"{name}_after"][0]) and _same(wr, g[f"{name}_after"][1])


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(G.REF_TWO_CASES))
def test_gpu_twoview_curve_matches_reference_golden(ref_two_gold, gpu_ctx, name):
    """The CUDA path (sr_run_view_curve in both directions with the two-view selection, then sr_cross_check)
    against the END-TO-END outputs of the reference's own TwoViewStereo::computeDepthMaps."""
    g = ref_two_gold
    a, b, mind, maxd, D = G.REF_TWO_CASES[name]
    _, imgs, ms, scale = G.ref_mvs_inputs(name)
    cams = G.cams_from_bytes(g[f"{name}_cams"])
    gpu_ctx.set_views(cams, [imgs[a], imgs[b]], [ms[a], ms[b]])
    gpu_ctx.set_params(T.default_params(False, mind, maxd, D, image_scale=scale))
    gpu_ctx.run_view_curve(0, [1])
    gpu_ctx.run_view_curve(1, [0])
    for v in (0, 1):
        ok = _depth_close(gpu_ctx.depth(v), g[f"{name}_before"][v], 1e-9)
        assert ok.mean() >= 1 - 1e-4, (v, 1 - ok.mean())
    gpu_ctx.cross_check(True, 1.0)
    for v in (0, 1):
        ok = _depth_close(gpu_ctx.depth(v), g[f"{name}_after"][v], 1e-9)
        assert ok.mean() >= 1 - 2e-4, (v, 1 - ok.mean())


def _camera_golden_check(leaves, i, rays, xy, ok, tol_ray=1e-12, tol_px=1e-8):
    gr, gxy, gok = leaves[f"cam{i}_rays"], leaves[f"cam{i}_xy"], leaves[f"cam{i}_ok"]
    assert np.abs(rays - gr).max() <= tol_ray * max(1.0, np.abs(gr).max())
    same = (ok != 0) == (gok != 0)
    assert same.mean() >= 0.995  # the +-1e-3 acceptance band of camera.cpp:121-134 is decided by roundings
    both = (ok != 0) & (gok != 0)
    assert both.mean() > 0.5
    err = np.abs(xy[both] - gxy[both]).max(axis=1)
    assert np.quantile(err, 0.995) <= tol_px and (err > 1e-6).mean() <= 0.005


def test_oracle_camera_matches_reference_golden(leaves):
    """Camera::unproject / Camera::project of the reference's own project/camera.cpp (golden) vs the oracle."""
    cams, pix, pts = G.camera_cases()
    L = O.lib()
    for i, c in enumerate(cams):
        sc = O.Scene([c], [np.zeros((480, 640, 4), np.uint8)])
        rays = np.empty((len(pix), 6))
        for k, (x, y) in enumerate(pix):
            L.orc_unproject(O.as_cam_array([c]), C.c_double(x), C.c_double(y), _dp(rays[k]))
        for root_mode in (0, 1):
            xy, ok = sc.project_points(0, pts[i], root_mode=root_mode)
            _camera_golden_check(leaves, i, rays, xy, ok)


@pytest.mark.gpu
def test_gpu_camera_matches_reference_golden(leaves, gpu_ctx):
    """sr_project_points and the ray table of sr_unproject_grid against the reference's own Camera."""
    cams, pix, pts = G.camera_cases()
    for i, c in enumerate(cams):
        gpu_ctx.set_views([c], [np.zeros((480, 640, 4), np.uint8)], None)
        gpu_ctx.set_params(T.default_params(True, 10.0, 100.0, 8))  # image_scale 1: the grid's pixel centres are (x + 0.5, y + 0.5)
        xy, ok = gpu_ctx.project_points(0, pts[i])
        # the grid holds pixel centres: compare the golden rays of the centres that are in it
        grid = gpu_ctx.unproject_grid(0)
        centre = np.array([[319.5, 239.5], [0.5, 0.5], [639.5, 479.5]])
        gr = leaves[f"cam{i}_rays"][-3:]
        got = np.array([grid[int(y), int(x)] for x, y in centre])
        assert np.abs(got - gr).max() <= 1e-11 * max(1.0, np.abs(gr).max())
        gxy, gok = leaves[f"cam{i}_xy"], leaves[f"cam{i}_ok"]
        same = (ok != 0) == (gok != 0)
        assert same.mean() >= 0.995
        both = (ok != 0) & (gok != 0)
        err = np.abs(xy - gxy).max(axis=1)
        mag = np.maximum(1.0, np.abs(gxy).max(axis=1))
        # points in view (the first 400): 1e-9 px; points anywhere (distorted coordinates reach 1e7 px): relative
        assert err[:400][both[:400]].max() <= 1e-9
        rel = (err / mag)[both]
        assert np.quantile(rel, 0.995) <= 1e-10 and (rel > 1e-8).mean() <= 0.005


@pytest.mark.gpu
def test_gpu_weights_match_reference_golden(leaves, gpu_ctx):
    """AdaptiveWeight / GeodesicWeight on the GPU against the reference's own compiled sources."""
    img = G.leaf_image()
    h, w = img.shape[:2]
    gpu_ctx.set_views([T.make_camera(np.eye(3), np.eye(3), np.zeros(3))], [img], None)
    cx, cy = G.weight_centres(w, h)
    for kind in (0, 1):
        for radius in G.WEIGHT_RADII:
            g = np.asarray(gpu_ctx.weights(0, kind, radius, cx, cy)).reshape(cx.size, -1)
            o = leaves[f"weights_k{kind}_r{radius}"]
            assert np.allclose(g, o, rtol=1e-13, atol=1e-300), (kind, radius, np.abs(g - o).max())


@pytest.mark.gpu
def test_gpu_mvs_matches_scene_golden(scenes_gold, gpu_ctx):
    cams, imgs, ms = G.arc_scene()
    gpu_ctx.set_views(cams, imgs, ms)
    nb = gpu_ctx.select_neighbours(3)
    assert _same(np.array(nb, np.int32), scenes_gold["arc_neighbours"])
    for name, P in G.arc_mvs_params().items():
        gpu_ctx.set_views(cams, imgs, ms)
        gpu_ctx.set_params(P)
        for ref in range(len(cams)):
            gpu_ctx.run_view(ref, nb[ref])
            _check_maps(gpu_ctx.depth_index(ref), gpu_ctx.depth(ref), gpu_ctx.best_cost(ref),
                        scenes_gold[f"{name}_v{ref}_index"], scenes_gold[f"{name}_v{ref}_depth"],
                        scenes_gold[f"{name}_v{ref}_best"], best_tol=1e-12)
        if name == "arc_mvs_geo_r2":
            gpu_ctx.cross_check(False, G.ARC_CROSS_THRESH)
            for v in range(len(cams)):
                g, o = gpu_ctx.depth(v), scenes_gold[f"{name}_v{v}_crosschecked"]
                assert ((g == o) | (np.isnan(g) & np.isnan(o))).mean() > 1 - 1e-4


@pytest.mark.gpu
def test_gpu_twoview_matches_scene_golden(scenes_gold, gpu_ctx):
    cams, imgs, ms = G.arc_scene()
    gpu_ctx.set_views(cams, imgs, ms)
    for name, (P, a, b) in G.arc_twoview_params().items():
        P.keep_cost_volume = 1
        gpu_ctx.set_params(P)
        gpu_ctx.run_view(a, [b])
        _check_maps(gpu_ctx.depth_index(a), gpu_ctx.depth(a), gpu_ctx.best_cost(a), scenes_gold[f"{name}_index"],
                    scenes_gold[f"{name}_depth"], scenes_gold[f"{name}_best"], best_tol=1e-4)
        gv = gpu_ctx.cost_volume(1)[0]
        assert cost_close(gv, scenes_gold[f"{name}_volume"]).mean() <= 1e-4
    cams2, imgs2 = G.rectified_scene()
    gpu_ctx.set_views(cams2, imgs2, None)
    P = G.rectified_params()
    gpu_ctx.set_params(P)
    for (a, b) in ((0, 1), (1, 0)):
        gpu_ctx.run_view(a, [b])
        _check_maps(gpu_ctx.depth_index(a), gpu_ctx.depth(a), gpu_ctx.best_cost(a), scenes_gold[f"rect_{a}{b}_index"],
                    scenes_gold[f"rect_{a}{b}_depth"], scenes_gold[f"rect_{a}{b}_best"], best_tol=1e-4)
