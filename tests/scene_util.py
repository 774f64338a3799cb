"""Small seeded scenes shared by the oracle tests (CPU) and the parity tests (GPU)."""
import numpy as np

from oracle import oracle_api as O
from stereoreconstruction_b200 import scenes, types as T


def refractive_arc_scene(V=4, w=96, h=64, arc_deg=25.0, seed=4321, masks=False, distortion=True,
                         interface=True, cell=14.0, amp=15.0):
    """cfg4 in miniature: V cameras on an arc behind tilted interfaces, photo-consistent images
    rendered through the oracle's unproject."""
    dist = (-0.1, 0.05, 0.001, 0.001, 0.0) if distortion else None
    cams = scenes.arc_cameras(V, w, h, arc_deg=arc_deg, distortion=dist, interface=interface)
    dummy = [np.zeros((h, w, 4), np.uint8)] * V
    sc0 = O.Scene(cams, dummy)
    surf = scenes.HeightField(z0=0.0, amp=amp, lx=60.0, ly=45.0)
    imgs = scenes.render_views(V, lambda v: sc0.unproject_grid(v), surf, seed=seed, cell=cell)
    ms = None
    if masks:
        rng = np.random.RandomState(seed)
        ms = []
        yy, xx = np.mgrid[0:h, 0:w]
        for v in range(V):
            cx, cy = rng.uniform(0.35, 0.65) * w, rng.uniform(0.35, 0.65) * h
            m = (((xx - cx) / (0.48 * w)) ** 2 + ((yy - cy) / (0.46 * h)) ** 2 < 1.0)
            m &= rng.rand(h, w) > 0.01  # a few isolated holes
            ms.append(np.where(m, 255, 0).astype(np.uint8))
    return cams, imgs, ms, surf


def rectified_scene(w=128, h=48, seed=1234, cell=None):
    """cfg3 in miniature: rectified pair, uniform-disparity labels when max_depth = 5*min_depth."""
    cams, B = scenes.rectified_pair(w, h, z0=100.0)
    dummy = [np.zeros((h, w, 4), np.uint8)] * 2
    sc0 = O.Scene(cams, dummy)
    f = cams[0].K[0]
    surf = scenes.HeightField(z0=167.0, amp=20.0, lx=25.0, ly=18.0)
    cell = cell if cell is not None else 3.5 * 167.0 / f
    imgs = scenes.render_views(2, lambda v: sc0.unproject_grid(v), surf, seed=seed, cell=cell)
    return cams, imgs, None, surf


def cost_close(a, b, rel=1e-4, abs_floor=1e-6):
    """|a-b| <= rel*max(|b|, abs_floor/rel...) with NaN == NaN; returns the mismatch mask."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    both_nan = np.isnan(a) & np.isnan(b)
    same_inf = np.isinf(a) & np.isinf(b) & (np.sign(a) == np.sign(b))
    with np.errstate(invalid="ignore"):
        ok = np.abs(a - b) <= rel * np.abs(b) + abs_floor
    return ~(ok | both_nan | same_inf)
