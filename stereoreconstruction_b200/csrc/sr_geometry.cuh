// sr_geometry.cuh — device-side camera geometry of the dense-matching path (FP64).
//
// What each function computes follows the reference (citations are into the reference tree);
// how it is computed is organised for the GPU: everything that is invariant per pixel or per
// (pixel, neighbour view) is hoisted, and the refractive projection replaces the reference's
// quartic eigen-solve (project/camera.cpp:68-138, GSL) by a safeguarded Newton iteration on the
// un-squared Snell equation, which has exactly one root on [0, r].
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/sr_b200.h"

namespace sr {

struct d3 {
    double x, y, z;
};
__device__ __forceinline__ d3 operator+(d3 a, d3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ d3 operator-(d3 a, d3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ d3 operator*(double s, d3 a) { return {s * a.x, s * a.y, s * a.z}; }
__device__ __forceinline__ double dot(d3 a, d3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ d3 normalized(d3 a) {
    double n = sqrt(dot(a, a));
    return {a.x / n, a.y / n, a.z / n};
}
__device__ __forceinline__ d3 mul3(const double *M, d3 v) {
    return {M[0] * v.x + M[1] * v.y + M[2] * v.z, M[3] * v.x + M[4] * v.y + M[5] * v.z,
            M[6] * v.x + M[7] * v.y + M[8] * v.z};
}
__device__ __forceinline__ d3 ld3(const double *p) { return {p[0], p[1], p[2]}; }

// util/ray.cpp:78-88 (intersect) with the plane given as (unit normal, distance).
__device__ __forceinline__ bool ray_plane(d3 src, d3 dir, d3 n, double dist, d3 &p) {
    double nd = dot(n, dir);
    if (fabs(nd) < 1e-10) return false;
    d3 x0 = dist * n;
    double t = dot(n, x0 - src) / nd;
    if (t < 1e-10) return false;
    p = src + t * dir;
    return true;
}

// Camera::unproject (project/camera.cpp:423-459): pixel -> global ray (source, unit direction).
__device__ inline void cam_unproject(const sr_camera &c, double px, double py, d3 &src, d3 &dir) {
    double x = px, y = py;
    if (c.is_distorted) {
        const double cx = c.K[2], cy = c.K[5];
        const double ifx = 1.0 / c.K[0], ify = 1.0 / c.K[4];
        const double x0 = x = (x - cx) * ifx;
        const double y0 = y = (y - cy) * ify;
        const double *k = c.dist;
#pragma unroll 1
        for (int j = 0; j < 5; j++) {
            const double r2 = x * x + y * y;
            const double icdist = 1.0 / (1 + ((k[4] * r2 + k[1]) * r2 + k[0]) * r2);
            const double deltaX = 2 * k[2] * x * y + k[3] * (r2 + 2 * x * x);
            const double deltaY = k[2] * (r2 + 2 * y * y) + 2 * k[3] * x * y;
            x = (x0 - deltaX) * icdist;
            y = (y0 - deltaY) * icdist;
        }
        x /= ifx;
        y /= ify;
        x += cx;
        y += cy;
    }
    d3 s = {0, 0, 0};
    d3 d = normalized(mul3(c.Kinv, d3{x, y, 1.0}));  // Ray3d ctor normalises
    if (c.is_refractive) {  // refract(), util/ray.cpp:92-106; failure leaves the ray unrefracted
        d3 N = ld3(c.plane_n);
        d3 hit;
        if (ray_plane(s, d, N, c.plane_d, hit)) {
            double cosI = -dot(N, d);
            double cosT2 = 1.0 - (1.0 - cosI * cosI) / (c.n * c.n);
            if (cosT2 > 0.0) {
                double sign = (cosI > 0.0 ? -1.0 : 1.0);
                s = hit;
                d = normalized(d + (cosI + c.n * sign * sqrt(cosT2)) * N);
            }
        }
    }
    // fromLocalToGlobal (camera.cpp:372-376); Ray3d ctor normalises again
    dir = normalized(mul3(c.Rinv, d));
    src = mul3(c.Rinv, s - ld3(c.t));
}

// Unique root in [0,r] of g(x) = x/sqrt(x^2+d^2) - n (r-x)/sqrt((r-x)^2+h^2).
// `guess` is a starting ratio x/r in (0,1) (warm start from the previous depth label) or < 0 for
// the paraxial start.  Safeguarded Newton in FP64; converges to ~1 ulp of the root.
__device__ __forceinline__ double snell_root(double r, double d, double h, double n, double guess) {
    const double dd = d * d, hh = h * h;
    double lo = 0.0, hi = r;
    double x = (guess > 0.0) ? guess * r : n * fabs(d) * r / (fabs(h) + n * fabs(d) + 1e-300);
    if (!(x > lo && x < hi)) x = 0.5 * r;
#pragma unroll 1
    for (int it = 0; it < 100; ++it) {
        const double rx = r - x;
        const double ia = rsqrt(x * x + dd), ib = rsqrt(rx * rx + hh);
        const double g = x * ia - n * rx * ib;
        if (g == 0.0) break;
        if (g < 0.0) lo = x; else hi = x;
        const double gp = dd * ia * ia * ia + n * hh * ib * ib * ib;
        double xn = x - g / gp;
        if (!(xn > lo && xn < hi)) xn = 0.5 * (lo + hi);
        const double dx = fabs(xn - x);
        x = xn;
        if (dx <= 4e-16 * r || hi - lo <= 4e-16 * r) break;
    }
    return x;
}

// Per (reference pixel, target view) invariants of Camera::project applied to points
// P(t) = src + t*dir on one ray: local(t) = Ls + t*Ld (camera.cpp:346-348 hoisted).
struct RayInView {
    d3 Ls, Ld;
};
__device__ __forceinline__ RayInView ray_in_view(const sr_camera &c, d3 src, d3 dir) {
    RayInView r;
    r.Ls = mul3(c.R, src) + ld3(c.t);
    r.Ld = mul3(c.R, dir);
    return r;
}

// Camera::project (project/camera.cpp:380-419) of the camera-local point `local`.
// `warm` carries x/r between consecutive depth labels (set < 0 before the first call).
__device__ __forceinline__ bool cam_project_local(const sr_camera &c, d3 local, double &warm, double &u, double &v) {
    d3 point = local;
    if (c.is_refractive) {  // projectRefraction, camera.cpp:95-138
        const d3 N = ld3(c.plane_n);
        const double a = dot(N, local);
        const d3 proj = a * N;
        const d3 radv = local - proj;
        const double rr = dot(radv, radv);
        const double r = sqrt(rr);
        const double z = fabs(a);  // |proj|
        if (!(r > 0.0)) return false;  // dir = radv/r is NaN: no root is accepted
        const double x = snell_root(r, c.plane_d, z - c.plane_d, c.n, warm);
        if (!(x == x)) return false;
        warm = x / r;
        // root in [0,r] always passes the y-component acceptance test of camera.cpp:119-135
        point = warm * radv + c.plane_d * N;
    }
    d3 p = mul3(c.K, point);
    double x = p.x / p.z, y = p.y / p.z;
    if (c.is_distorted) {
        const double cx = c.K[2], cy = c.K[5], fx = c.K[0], fy = c.K[4];
        x = (x - cx) / fx;
        y = (y - cy) / fy;
        const double *k = c.dist;
        const double r2 = x * x + y * y;
        const double cdist = 1 + ((k[4] * r2 + k[1]) * r2 + k[0]) * r2;
        const double xo = x, yo = y;
        x = xo * cdist + 2 * k[2] * xo * yo + k[3] * (r2 + 2 * xo * xo);
        // camera.cpp:411-412: y's tangential term uses the already-distorted x
        y = yo * cdist + k[2] * (r2 + 2 * yo * yo) + 2 * k[3] * x * yo;
        x = fx * x + cx;
        y = fy * y + cy;
    }
    u = x;
    v = y;
    return true;
}

// Camera::project of a global point.
__device__ __forceinline__ bool cam_project(const sr_camera &c, d3 p, double &u, double &v) {
    double warm = -1.0;
    return cam_project_local(c, mul3(c.R, p) + ld3(c.t), warm, u, v);
}

// The implicit double->int conversion of projected coordinates (util/lineiter.hpp:34 ctor
// arguments, cost_ncc's int x2,y2): truncation toward zero; out of range / NaN -> INT_MIN as
// x86 cvttsd2si does (the oracle fixes the same behaviour).
__device__ __forceinline__ int to_int_x86(double v) {
    if (!(v > -2147483649.0 && v < 2147483648.0)) return INT32_MIN;
    return __double2int_rz(v);
}

}  // namespace sr
