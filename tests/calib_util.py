"""A synthetic refractive-interface calibration problem (stereo/refractioncalibration.cpp): V arc
cameras behind tilted interfaces, correspondences = projections of scene points into view pairs
(+ pixel noise), and a model [n, (px, py, dist) x V] that starts away from the truth."""
import copy

import numpy as np

from stereoreconstruction_b200 import scenes


def model_of(cams, n=None):
    """The model vector the GUI builds (gui/widgets/stereowidget.cpp:573-592): pixel = K * normal / z."""
    m = [cams[0].n if n is None else n]
    for c in cams:
        K = np.array(c.K[:]).reshape(3, 3)
        p = K @ np.array(c.plane_n[:])
        p /= p[2]
        m += [p[0], p[1], c.plane_d]
    return np.array(m)


def calibration_problem(oracle_scene_cls, V=3, n=240, noise=0.05, seed=5, w=640, h=480):
    rng = np.random.RandomState(seed)
    cams = scenes.arc_cameras(V, w, h, arc_deg=30.0)
    sc = oracle_scene_cls(cams, [np.zeros((h, w, 4), np.uint8)] * V)
    rays = sc.unproject_grid(0)
    pts = (rays[..., :3] + rng.uniform(450, 550, rays.shape[:2] + (1,)) * rays[..., 3:]).reshape(-1, 3)
    pts = pts[rng.randint(0, pts.shape[0], n)]
    pairs = np.array([(a, b) for a in range(V) for b in range(a + 1, V)], dtype=np.int32)[rng.randint(0, V * (V - 1) // 2, n)]
    pix = np.empty((n, 4))
    okall = np.ones(n, bool)
    for v in range(V):
        xy, ok = sc.project_points(v, pts)
        for side in (0, 1):
            m = pairs[:, side] == v
            pix[m, 2 * side:2 * side + 2] = xy[m]
            okall[m] &= ok[m] != 0
    pix += rng.normal(scale=noise, size=pix.shape)
    pairs, pix = pairs[okall], pix[okall]
    truth = model_of(cams)
    start = truth.copy()
    for v in range(V):  # a wrong first guess of every interface
        start[3 * v + 1] += rng.uniform(-25, 25)
        start[3 * v + 2] += rng.uniform(-25, 25)
        start[3 * v + 3] *= rng.uniform(0.8, 1.25)
    # cameras as the caller would hold them before calibration: no interface knowledge yet
    return [copy.copy(c) for c in cams], pairs, pix, truth, start
