"""Pins oracle/oracle.cpp against the REFERENCE'S OWN leaf sources compiled from /root/reference
(oracle/_ref/libref.so: util/lineiter.cpp, util/ray.cpp, util/vectorimage.cpp,
stereo/adaptiveweight.cpp, stereo/geodesicweight.cpp — see oracle/Makefile).  Bit-exact."""
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
