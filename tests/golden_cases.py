"""Seeded inputs shared by tests/golden/make_golden.py (which writes the fixtures) and
tests/test_golden.py (which replays them through the oracle and the CUDA path)."""
import numpy as np

from stereoreconstruction_b200 import scenes, types as T
from scene_util import refractive_arc_scene, rectified_scene as _rect

WEIGHT_RADII = (1, 2, 5)
ARC_CROSS_THRESH = 12.0


def line_cases():
    w, h = 64, 48
    rng = np.random.RandomState(11)
    cases = [(0, 0, 10, 3), (10, 3, 0, 0), (5, 5, 5, 5), (0, 0, 0, 9), (7, 2, 7, -6), (-5, -5, 70, 60),
             (-20, 10, 90, 12), (30, -40, 31, 90), (63, 47, 0, 0), (64, 48, -1, -1), (100, 100, 200, 150)]
    for _ in range(200):
        cases.append(tuple(int(v) for v in rng.randint(-40, 110, 4)))
    return cases, w, h


def ray_cases():
    rng = np.random.RandomState(5)
    out = []
    for _ in range(300):
        s1, d1, s2, d2, pn = (rng.normal(size=3) for _ in range(5))
        out.append((s1, d1, s2, d2, pn, float(rng.normal() * 3), float(rng.uniform(0.6, 1.7))))
    return out


def leaf_image():
    rng = np.random.RandomState(2)
    h, w = 37, 53
    img = rng.randint(0, 256, size=(h, w, 4)).astype(np.uint8)
    img[..., 3] = 255
    img[5:9, 7:12] = 255  # a WHITE patch
    img[20:30, 30:45, :3] //= 8  # a dark, low-contrast patch: long geodesic paths
    return img


def sample_points(w, h):
    rng = np.random.RandomState(9)
    pts = [(float(rng.uniform(-2, w + 1)), float(rng.uniform(-2, h + 1))) for _ in range(500)]
    pts += [(float(x), float(y)) for x in (-1, 0, 1, w - 2, w - 1, w) for y in (-1, 0, 1, h - 2, h - 1, h)]
    return pts


def weight_centres(w, h):
    rng = np.random.RandomState(4)
    cx = np.concatenate([rng.randint(0, w, 30), [0, w - 1, 0, w - 1, 1]]).astype(np.int32)
    cy = np.concatenate([rng.randint(0, h, 30), [0, 0, h - 1, h - 1, h - 2]]).astype(np.int32)
    return cx, cy


def arc_scene():
    cams, imgs, ms, _ = refractive_arc_scene(V=4, w=96, h=64, masks=True)
    return cams, imgs, ms


def arc_mvs_params():
    return {
        "arc_mvs_geo_r2": T.default_params(True, 420.0, 580.0, 40),
        "arc_mvs_ada_r2": T.default_params(True, 420.0, 580.0, 40, weight_kind=T.SR_WEIGHT_ADAPTIVE),
        "arc_mvs_geo_r3": T.default_params(True, 420.0, 580.0, 24, radius=3),
    }


def arc_twoview_params():
    return {
        "arc_two_ncc_geo_r5": (T.default_params(False, 420.0, 580.0, 24, radius=5), 1, 2),
        "arc_two_sad_ada_r2": (T.default_params(False, 420.0, 580.0, 24, radius=2, weight_kind=T.SR_WEIGHT_ADAPTIVE,
                                                cost_kind=T.SR_COST_SAD_TWOVIEW), 2, 1),
    }


def rectified_scene():
    cams, imgs, _, _ = _rect(w=112, h=40)
    return cams, imgs


def rectified_params():
    return T.default_params(False, 100.0, 500.0, 32, radius=16, weight_kind=T.SR_WEIGHT_ADAPTIVE)


def camera_cases():
    """Cameras covering every branch of project/camera.cpp (refractive x distorted), pixels to
    unproject and global points to project for each."""
    cams = []
    for interface, dist in ((True, (-0.1, 0.05, 0.001, 0.001, 0.0)), (True, (0.0,) * 5),
                            (False, (-0.1, 0.05, 0.001, 0.001, 0.0)), (False, (0.0,) * 5)):
        cams += scenes.arc_cameras(4, 640, 480, arc_deg=40.0, interface=interface, distortion=dist)[1:3]
    rng = np.random.RandomState(31)
    pix = np.concatenate([rng.uniform(-20, 660, (300, 1)), rng.uniform(-20, 500, (300, 1))], axis=1)
    pix = np.ascontiguousarray(np.concatenate([pix, [[319.5, 239.5], [0.5, 0.5], [639.5, 479.5]]]))
    # points: pixel rays of an ideal pinhole at the camera, pushed to depths 200..900, plus a box of points anywhere
    pts = []
    for c in cams:
        Kinv = np.array(c.Kinv[:]).reshape(3, 3)
        Rinv = np.array(c.Rinv[:]).reshape(3, 3)
        Cc = np.array(c.C[:])
        uv = np.concatenate([rng.uniform(0, 640, (400, 1)), rng.uniform(0, 480, (400, 1)), np.ones((400, 1))], axis=1)
        d = (Rinv @ (Kinv @ uv.T)).T
        p = Cc + d * rng.uniform(200, 900, (400, 1))
        pts.append(np.ascontiguousarray(np.concatenate([p, rng.uniform(-400, 400, (100, 3)) + [0, 0, 100]])))
    return cams, pix, pts


def settled_cameras(cams):
    """Cameras passed through the reference's own Camera::set (which re-orthonormalises R and re-derives
    Kinv, Rinv, C) until the derived state reproduces itself, so that the reference and the code
    under test start from bit-identical camera state.  Needs oracle/_ref (the build container)."""
    import ctypes as C
    from oracle import oracle_api as O
    REF = O.ref_lib()
    out = []
    for c in cams:
        cur = O.as_cam_array([c])
        for _ in range(6):
            nxt = O.OrcCamera()
            REF.ref_camera_derived(cur, C.byref(nxt))
            done = bytes(cur[0]) == bytes(nxt)
            cur = O.as_cam_array([nxt])
            if done:
                break
        s = T.SrCamera()
        C.memmove(C.byref(s), C.byref(cur[0]), C.sizeof(s))
        out.append(s)
    return out


def cams_to_bytes(cams):
    import ctypes as C
    return np.frombuffer(b"".join(bytes(c) for c in cams), dtype=np.uint8).copy()


def cams_from_bytes(buf):
    import ctypes as C
    n = C.sizeof(T.SrCamera)
    out = []
    for i in range(len(buf) // n):
        c = T.SrCamera()
        C.memmove(C.byref(c), bytes(buf[i * n:(i + 1) * n]), n)
        out.append(c)
    return out


# Depth range of the bunny cases: README.md:103-112 quotes "min and max depth of 300 and 800" for cameras
# re-calibrated in the GUI (its own units).  With the projection matrices example/project.xml SHIPS, the eight
# cameras sit on a semi-circle of radius ~42 around the object and the object's depth is 35..50: at 300..800 no
# candidate ever reaches ncc > 0.95 and every map is the "nothing found" sentinel (round 1 compared exactly
# that).  30..55 brackets the object: > 99 % of the in-mask pixels get a depth.
REF_MVS_CASES = {
    # name: (min depth, max depth, levels, cross-check threshold)
    "arc": (420.0, 580.0, 40, 12.0),
    "bunny_refr": (30.0, 55.0, 100, 5.0),   # SURVEY 8d cfg2 with the injected interface (depth range: see above)
    # the same reference file built with `typedef AdaptiveWeight WeightFunc` (BASELINE configs[1]: adaptive-weight aggregation)
    "arc_ada": (420.0, 580.0, 40, 12.0),
    "bunny_refr_ada": (30.0, 55.0, 100, 5.0),
}


def ref_mvs_kept_views(name, V):
    """Views whose depth maps before the cross-check are kept in ref_mvs.npz.  (With a depth range that finds
    the object the maps are real-valued and do not compress: the AdaptiveWeight bunny case keeps two views.)"""
    return [1, 2] if name == "bunny_refr_ada" else list(range(V))


def ref_mvs_peak_rows(name, h):
    """Rows of view 1 whose K = 9 peak lists are kept."""
    return (0, h) if name.startswith("arc") else (h // 2 - 16, h // 2 + 16)


def ref_mvs_after(g, name, shape):
    """The depth maps after the cross-check, rebuilt from the maps before it and the mask of the pixels it
    invalidated (None when the case keeps a subset of the views)."""
    if f"{name}_after_nan" not in g:
        return None
    before = g[f"{name}_before"]
    nan_after = np.unpackbits(g[f"{name}_after_nan"])[:before.size].reshape(before.shape).astype(bool)
    return np.where(nan_after, np.nan, before)


def ref_mvs_adaptive(name):
    return name.endswith("_ada")


def ref_mvs_inputs(name):
    """(cams, images, masks, image_scale) of a reference end-to-end case (cameras NOT yet settled)."""
    if name.endswith("_ada"):
        name = name[:-4]
    if name == "arc":
        cams, imgs, ms = arc_scene()
        return cams, imgs, ms, 1.0
    from bunny_util import load
    cams, imgs, ms, scale = load(refractive=(name == "bunny_refr"))
    return cams, imgs, ms, scale


REF_TWO_CASES = {
    # name: (left view, right view, min depth, max depth, levels); r = 5 GeodesicWeight NCC, cross-check 1 (the class's constants)
    "arc": (1, 2, 420.0, 580.0, 32),
    "bunny": (0, 1, 30.0, 55.0, 100),       # cameras 7310085 + 7310087
    "bunny_refr": (0, 1, 30.0, 55.0, 100),
}
