// sr_geometry.cuh — device-side camera geometry of the dense-matching path (FP64).
//
// What each function computes follows the reference (citations are into the reference tree);
// how it is computed is organised for the GPU: everything that is invariant per pixel or per
// (pixel, neighbour view) is hoisted, and the refractive projection replaces the reference's
// quartic eigen-solve (project/camera.cpp:68-138, GSL) by Newton on the un-squared Snell
// equation, which has exactly one root on [0, r].
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
// The library is compiled with -fmad=false so that plain expressions round exactly like the
// CPU oracle (IEEE, no contraction).  Where fusion is wanted (the refractive fast path, whose
// root solve is not bit-comparable anyway) it is written explicitly:
__device__ __forceinline__ double fdot(d3 a, d3 b) { return fma(a.x, b.x, fma(a.y, b.y, a.z * b.z)); }
__device__ __forceinline__ d3 faxpy(double s, d3 a, d3 b) { return {fma(s, a.x, b.x), fma(s, a.y, b.y), fma(s, a.z, b.z)}; }
__device__ __forceinline__ d3 fmul3(const double *M, d3 v) {
    return {fma(M[0], v.x, fma(M[1], v.y, M[2] * v.z)), fma(M[3], v.x, fma(M[4], v.y, M[5] * v.z)),
            fma(M[6], v.x, fma(M[7], v.y, M[8] * v.z))};
}

// Branch-free FP64 reciprocal / reciprocal square root for the refractive fast path: MUFU seed
// (rcp/rsqrt.approx.ftz.f64, ~2^-22 relative) + two Newton steps -> ~1 ulp.  Arguments there are
// finite, positive and far from the denormal range, so the IEEE special-case handling that makes
// the library versions ~25 instructions each is not needed.
__device__ __forceinline__ double fast_rcp(double a) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a));
    double e = fma(-a, r, 1.0);
    r = fma(r, e, r);
    e = fma(-a, r, 1.0);
    return fma(r, e, r);
}
__device__ __forceinline__ double fast_rsqrt(double a) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
    const double ha = 0.5 * a;
    double e = fma(-ha * y, y, 0.5);
    y = fma(y, e, y);
    e = fma(-ha * y, y, 0.5);
    return fma(y, e, y);
}

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

// Unique root in [0,r] of g(x) = x/sqrt(x^2+d^2) - n (r-x)/sqrt((r-x)^2+h^2)   (Snell, un-squared).
// Robust FP64 solver: Newton with a bracketing safeguard.  `guess` is a starting ratio x/r or < 0
// for the paraxial start.  Convergence is tested on the Newton step BEFORE the safeguard (a step
// that rounds to zero must terminate, not bisect), and because convergence is quadratic a step
// below 2e-9*r already leaves an error below 1e-17*r once it is applied.
__device__ __noinline__ double snell_root_robust(double r, double d, double h, double n, double guess) {
    const double dd = d * d, hh = h * h;
    double lo = 0.0, hi = r;
    double x = (guess > 0.0) ? guess * r : n * fabs(d) * r / (fabs(h) + n * fabs(d) + 1e-300);
    if (!(x >= lo && x <= hi)) x = 0.5 * r;
#pragma unroll 1
    for (int it = 0; it < 80; ++it) {
        const double rx = r - x;
        const double ia = rsqrt(x * x + dd), ib = rsqrt(rx * rx + hh);
        const double g = x * ia - n * rx * ib;
        if (g == 0.0) break;
        if (g < 0.0) lo = x; else hi = x;
        const double gp = dd * ia * ia * ia + n * hh * ib * ib * ib;
        const double step = g / gp;
        double xn = x - step;
        if (fabs(step) <= 2e-9 * r) {
            x = xn;
            break;
        }
        if (!(xn >= lo && xn <= hi)) xn = 0.5 * (lo + hi);
        x = xn;
        if (hi - lo <= 4e-16 * r) break;
    }
    return x;
}

// Fast path of the same solve, the one the build kernel runs per (pixel, label, view):
// FP32 Newton from a warm start (the ratio x/r moves smoothly along the depth axis; w0,w1 hold
// the ratios of the two previous labels and are linearly extrapolated), a fixed number of
// iterations so the warp does not diverge, then ONE Newton step in FP64: with an FP32-converged
// start (|e0| ~ 1e-6) the quadratic error after the polish is ~|g''/2g'| e0^2 < 1e-14, i.e.
// ~1e-12 px.  If the polish step is not tiny the robust FP64 loop takes over.
__device__ __forceinline__ double snell_root_fast(double r, double d, double h, double n, float &w0, float &w1) {
    const float rf = (float)r, hf = (float)h, df = (float)d, nf = (float)n;
    const float ddf = df * df, hhf = hf * hf;
    const bool warm = w1 > 0.0f;
    float x;
    if (warm) x = (w0 > 0.0f ? fmaf(2.0f, w1, -w0) : w1) * rf;
    else x = nf * fabsf(df) * rf / (fabsf(hf) + nf * fabsf(df));
    x = fminf(fmaxf(x, 0.0f), rf);
    // cold start (first label of a chunk): 5 iterations; warm start with one previous ratio: 2;
    // with the linear extrapolation from two previous ratios the start is already within ~1e-5
    // of the root and one FP32 iteration reaches FP32 accuracy.
    const int iters = warm ? (w0 > 0.0f ? 1 : 2) : 5;
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
        const float rx = rf - x;
        const float ia = rsqrtf(fmaf(x, x, ddf)), ib = rsqrtf(fmaf(rx, rx, hhf));
        const float g = fmaf(x, ia, -(nf * rx) * ib);
        const float gp = fmaf(ddf * ia, ia * ia, (nf * hhf) * ib * (ib * ib));
        x = fminf(fmaxf(x - __fdividef(g, gp), 0.0f), rf);
    }
    double X = (double)x;
    {
        const double dd = d * d, hh = h * h;
        const double rx = r - X;
        const double ia = fast_rsqrt(fma(X, X, dd)), ib = fast_rsqrt(fma(rx, rx, hh));
        const double g = fma(X, ia, -(n * rx) * ib);
        const float gpf = fmaf(ddf * (float)ia, (float)(ia * ia), (nf * hhf) * (float)ib * (float)(ib * ib));
        // the step is ~1e-6: an FP32-accurate 1/g' (1e-7 relative) changes x by ~1e-13, i.e. ~1e-11 px
        const double step = g * (double)__frcp_rn(gpf);
        X -= step;
        if (!(fabs(step) <= 2e-5 * r) || !(X >= 0.0 && X <= r)) X = snell_root_robust(r, d, h, n, -1.0);
    }
    w0 = w1;
    w1 = __fdividef(x, rf);
    return X;
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
// w0,w1: warm-start ratios of the two previous depth labels (set < 0 before the first call).
// Divisions by per-camera constants become multiplications with reciprocals computed once per
// thread (<= 1 ulp apart from the reference's divisions).
struct ProjConsts {
    double inv_fx, inv_fy;
};
__device__ __forceinline__ ProjConsts proj_consts(const sr_camera &c) { return {1.0 / c.K[0], 1.0 / c.K[4]}; }

__device__ __forceinline__ bool cam_project_local(const sr_camera &c, const ProjConsts &pc, d3 local, float &w0,
                                                  float &w1, double &u, double &v) {
    d3 point = local;
    if (c.is_refractive) {  // projectRefraction, camera.cpp:95-138
        const d3 N = ld3(c.plane_n);
        const double a = fdot(N, local);
        const d3 radv = faxpy(-a, N, local);
        const double rr = fdot(radv, radv);
        if (!(rr > 0.0)) return false;  // dir = radv/r is NaN in the reference: no root is accepted
        const double ir = fast_rsqrt(rr);
        const double r = rr * ir;
        const double x = snell_root_fast(r, c.plane_d, fabs(a) - c.plane_d, c.n, w0, w1);
        if (!(x == x)) return false;
        // a root in [0,r] always passes the y-component acceptance test of camera.cpp:119-135
        point = faxpy(x * ir, radv, c.plane_d * N);
    }
    const d3 p = fmul3(c.K, point);
    const double iz = c.is_refractive ? fast_rcp(p.z) : 1.0 / p.z;
    double x = p.x * iz, y = p.y * iz;
    if (c.is_distorted) {
        const double cx = c.K[2], cy = c.K[5], fx = c.K[0], fy = c.K[4];
        x = (x - cx) * pc.inv_fx;
        y = (y - cy) * pc.inv_fy;
        const double *k = c.dist;
        const double r2 = fma(x, x, y * y);
        const double cdist = fma(fma(fma(k[4], r2, k[1]), r2, k[0]), r2, 1.0);
        const double xo = x, yo = y;
        x = fma(xo, cdist, fma(2 * k[2] * xo, yo, k[3] * fma(2 * xo, xo, r2)));
        // camera.cpp:411-412: y's tangential term uses the already-distorted x
        y = fma(yo, cdist, fma(k[2], fma(2 * yo, yo, r2), 2 * k[3] * x * yo));
        x = fma(fx, x, cx);
        y = fma(fy, y, cy);
    }
    u = x;
    v = y;
    return true;
}

// Camera::project (project/camera.cpp:380-419) of a GLOBAL point for a NON-refractive camera,
// operation for operation as the reference evaluates it (and as oracle.cpp restates it): with
// -fmad=false every +,*,/ below rounds exactly as on the CPU, so the integer truncation of the
// projected coordinate agrees even when the projection lands exactly on a pixel boundary
// (rectified pairs: y2 = v - 0.5 is an exact integer in exact arithmetic).
__device__ __forceinline__ void cam_project_exact(const sr_camera &c, d3 pg, double &u, double &v) {
    const d3 point = mul3(c.R, pg) + ld3(c.t);
    const d3 p = mul3(c.K, point);
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
        y = yo * cdist + k[2] * (r2 + 2 * yo * yo) + 2 * k[3] * x * yo;
        x = fx * x + cx;
        y = fy * y + cy;
    }
    u = x;
    v = y;
}

// Camera::project of a global point.
__device__ __forceinline__ bool cam_project(const sr_camera &c, d3 p, double &u, double &v) {
    if (!c.is_refractive) {
        cam_project_exact(c, p, u, v);
        return true;
    }
    float w0 = -1.0f, w1 = -1.0f;
    return cam_project_local(c, proj_consts(c), mul3(c.R, p) + ld3(c.t), w0, w1, u, v);
}

// The implicit double->int conversion of projected coordinates (util/lineiter.hpp:34 ctor
// arguments, cost_ncc's int x2,y2): truncation toward zero; out of range / NaN -> INT_MIN as
// x86 cvttsd2si does (the oracle fixes the same behaviour).
__device__ __forceinline__ int to_int_x86(double v) {
    if (!(v > -2147483649.0 && v < 2147483648.0)) return INT32_MIN;
    return __double2int_rz(v);
}

}  // namespace sr
