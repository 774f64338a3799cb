"""Synthetic refractive scenes of the shapes BASELINE.json names.

Host-side input construction only (numpy): cameras on an arc behind tilted planar interfaces,
and photo-consistent images rendered by casting each pixel's (refracted) ray onto a textured
height field.  The per-pixel rays come from a caller-supplied `rays_fn(view_index) -> (h,w,6)`
(the GPU's sr_unproject_grid in bench.py, the oracle's in tests), so that this module stays
free of both the CUDA library and the oracle.
"""
import numpy as np

from .types import make_camera


def _look_at(C, target, up=(0.0, 1.0, 0.0)):
    """World->camera rotation with +z towards target, +x right, +y down-ish."""
    z = np.asarray(target, float) - np.asarray(C, float)
    z /= np.linalg.norm(z)
    x = np.cross(np.asarray(up, float), z)
    x /= np.linalg.norm(x)
    y = np.cross(z, x)
    return np.stack([x, y, z])  # rows


def arc_cameras(V, w, h, f=None, radius=500.0, arc_deg=60.0, distortion=(-0.1, 0.05, 0.001, 0.001, 0.0),
                interface=True, tilt_px=(60.0, -35.0), plane_d=25.0, n=1.333):
    """cfg4/cfg5 geometry: V cameras on an arc of `arc_deg` degrees and radius `radius`, all
    looking at the origin; K=[[f,0,(w-1)/2],[0,f,(h-1)/2],[0,0,1]]; interface normal
    K^-1 (cx+tilt_x, cy+tilt_y, 1), distance plane_d, index ratio n."""
    if f is None:
        f = 1600.0 * w / 1920.0
    cx, cy = (w - 1) / 2.0, (h - 1) / 2.0
    K = np.array([[f, 0, cx], [0, f, cy], [0, 0, 1.0]])
    cams = []
    for i in range(V):
        a = np.deg2rad(-arc_deg / 2 + arc_deg * (i / max(V - 1, 1)))
        Cc = np.array([radius * np.sin(a), 0.0, -radius * np.cos(a)])
        R = _look_at(Cc, (0, 0, 0))
        t = -R @ Cc
        pn = None
        if interface:
            s = w / 1920.0
            pn = np.linalg.inv(K) @ np.array([cx + tilt_px[0] * s, cy + tilt_px[1] * s, 1.0])
        cams.append(make_camera(K, R, t, dist=distortion, plane_normal=pn,
                                plane_d=plane_d if interface else 0.0, n=n if interface else 1.0))
    return cams


def rectified_pair(w, h, f=None, z0=100.0, max_disp=318.75):
    """cfg3 geometry: K=[[f,0,(w-1)/2],[0,f,(h-1)/2],[0,0,1]], R=I, baseline along +x with
    f*B/z0 = max_disp, no distortion, n=1.  With min_depth=z0, max_depth=5*z0 and the two-view
    depthFromLabel the D labels are uniform in disparity (max_disp .. max_disp/5)."""
    if f is None:
        f = 1600.0 * w / 1920.0
    md = max_disp * w / 1920.0
    B = md * z0 / f
    K = np.array([[f, 0, (w - 1) / 2.0], [0, f, (h - 1) / 2.0], [0, 0, 1.0]])
    left = make_camera(K, np.eye(3), np.zeros(3))
    right = make_camera(K, np.eye(3), np.array([-B, 0.0, 0.0]))  # C = (B,0,0)
    return [left, right], B


def _hash_noise(ix, iy, seed):
    v = (ix.astype(np.int64) * 73856093) ^ (iy.astype(np.int64) * 19349663) ^ (seed * 83492791)
    v = (v ^ (v >> 13)) * 1274126177
    v = v ^ (v >> 16)
    return (v & 0xFFFF).astype(np.float64) / 65535.0


def texture(u, v, seed, cell=1.7):
    """Procedural RGB texture in [0,255]: 6 sinusoids + bilinearly interpolated hash noise."""
    rng = np.random.RandomState(seed)
    out = np.zeros(u.shape + (3,))
    for c in range(3):
        acc = np.zeros_like(u)
        for _ in range(6):
            fx, fy = rng.uniform(0.02, 0.6, 2) * (1.7 / cell)
            ph = rng.uniform(0, 2 * np.pi)
            acc += np.sin(fx * u + fy * v + ph)
        g = cell  # noise cell size in scene units (keep it >= ~3 pixel footprints)
        gu, gv = u / g, v / g
        iu, iv = np.floor(gu), np.floor(gv)
        du, dv = gu - iu, gv - iv
        n00 = _hash_noise(iu, iv, seed + c)
        n10 = _hash_noise(iu + 1, iv, seed + c)
        n01 = _hash_noise(iu, iv + 1, seed + c)
        n11 = _hash_noise(iu + 1, iv + 1, seed + c)
        nz = (n00 * (1 - du) + n10 * du) * (1 - dv) + (n01 * (1 - du) + n11 * du) * dv
        out[..., c] = 127.5 + 18.0 * acc + 110.0 * (nz - 0.5)
    return np.clip(np.rint(out), 0, 255).astype(np.uint8)


class HeightField:
    """Surface z = z0 + amp * sin(x/lx) * cos(y/ly) in world coordinates."""

    def __init__(self, z0=0.0, amp=25.0, lx=90.0, ly=70.0):
        self.z0, self.amp, self.lx, self.ly = z0, amp, lx, ly

    def z(self, x, y):
        return self.z0 + self.amp * np.sin(x / self.lx) * np.cos(y / self.ly)

    def hit(self, rays):
        """rays (h,w,6) -> hit points (h,w,3) by fixed-point iteration on the ray parameter."""
        s, d = rays[..., :3], rays[..., 3:]
        t = (self.z0 - s[..., 2]) / d[..., 2]
        for _ in range(30):
            p = s + t[..., None] * d
            t = (self.z(p[..., 0], p[..., 1]) - s[..., 2]) / d[..., 2]
        return s + t[..., None] * d


def render_views(num_views, rays_fn, surface, seed, cell=1.7):
    """RGBA8 images (alpha 255) of a textured height field seen along per-pixel rays."""
    imgs = []
    for v in range(num_views):
        rays = rays_fn(v)
        p = surface.hit(rays)
        rgb = texture(p[..., 0], p[..., 1], seed, cell)
        bad = ~np.isfinite(p).all(axis=-1)
        rgb[bad] = 0
        img = np.empty(rgb.shape[:2] + (4,), dtype=np.uint8)
        img[..., :3] = rgb
        img[..., 3] = 255
        imgs.append(img)
    return imgs


def noise_images(num_views, w, h, seed):
    """Band-limited noise images (uniform uint8, 3x3 box blur) — content is irrelevant for
    throughput; used when photo-consistency is not needed."""
    imgs = []
    for v in range(num_views):
        rng = np.random.RandomState(seed + v)
        a = rng.randint(0, 256, size=(h + 2, w + 2, 3)).astype(np.float64)
        acc = np.zeros((h, w, 3))
        for dy in range(3):
            for dx in range(3):
                acc += a[dy:dy + h, dx:dx + w]
        img = np.empty((h, w, 4), dtype=np.uint8)
        img[..., :3] = np.clip(np.rint(acc / 9.0), 0, 255).astype(np.uint8)
        img[..., 3] = 255
        imgs.append(img)
    return imgs


def nearest_neighbours(cams, max_n=3):
    """MultiViewStereo::runTask neighbour rule (stereo/multiviewstereo.cpp:335-360): views whose
    principal directions satisfy |d_i . d_j| > 0.2, the max_n closest by squared centre
    distance (all of them, unsorted, if there are <= max_n)."""
    V = len(cams)
    out = []
    for i in range(V):
        near = []
        di, Ci = np.array(cams[i].prin_dir[:]), np.array(cams[i].C[:])
        for j in range(V):
            if i == j:
                continue
            dj, Cj = np.array(cams[j].prin_dir[:]), np.array(cams[j].C[:])
            if abs(di.dot(dj)) > 0.2:
                near.append((float((Ci - Cj).dot(Ci - Cj)), j))
        if max_n < len(near):
            near.sort()
            near = near[:max_n]
        out.append([j for _, j in near])
    return out
