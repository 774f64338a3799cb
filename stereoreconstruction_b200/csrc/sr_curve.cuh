// sr_curve.cuh — curve-mode search, stage (1): the rasterised (refractive) epipolar curve.
//
// This is the reference's LIVE formulation (stereo/multiviewstereo.cpp:574-596,754-810 and
// stereo/twoviewstereo.cpp:285-302,999-1054): per reference pixel and neighbour view, the D depth
// labels are projected into the neighbour, consecutive projections at least one pixel apart are
// joined by Bresenham segments (util/lineiter.hpp:32-118; clipped to the mask rectangle in the
// multi-view class, util/lineiter.cpp:44-88; unclipped in the two-view class), every segment pixel
// whose neighbour-mask is WHITE becomes a candidate, and — multi-view class only — consecutive
// duplicates are removed (:800-807).
//
// The kernel emits the candidates in the SAME packed-tap layout the label-mode build uses, one
// plane per candidate ordinal: taps[ordinal][rows][w], TAP_NONE after the end of a pixel's curve.
// The match kernels then stream it exactly like a label volume; what differs downstream is only
// where the depth of a candidate comes from (closest approach of the two viewing rays, S2/S3).
// Curve lengths vary per pixel: the host fills a volume of 2 D + 64 planes (curves are rarely longer
// than the label count), reads back the longest curve, and only if that exceeds the capacity
// sizes the volume to it and runs the kernel again.
#pragma once
#include "sr_build_refr.cuh"

namespace sr {

struct CurveArgs {
    sr_camera nbr;               // target view
    double Kn[9];                // see BuildRefrArgs (refractive targets)
    double fxs, cxs, fys, cys;   // curve mode: pixel*scale, no -0.5 shift (twoviewstereo.cpp:1019, multiviewstereo.cpp:774)
    double prin[3], C[3];        // reference view principal direction and centre
    double scale;
    const double *rays;          // [6][h][w] of the reference view
    const double *depth_table;   // [D]
    const uint8_t *ref_mask;
    const uint8_t *nbr_mask;     // a plane of 255 when the view was given no mask (then the launch uses HAS_MASK = false)
    int32_t *taps;               // [capacity][rows][w] for this neighbour (null: only measure the curve lengths)
    int32_t *max_count;          // device scalar, atomicMax of the curve lengths over all pixels
    int w, h, row0, rows, D, capacity;
    int mvs;                     // 1: clipped iterator + consecutive-duplicate removal; 0: two-view flavour
};

constexpr int CS_LEFT = 1, CS_RIGHT = 2, CS_BOTTOM = 4, CS_TOP = 8;

__device__ __forceinline__ int cs_outcode(int x, int y, int w, int h) {
    int code = 0;
    if (x < 0) code |= CS_LEFT;
    else if (x > w) code |= CS_RIGHT;
    if (y < 0) code |= CS_BOTTOM;
    else if (y > h) code |= CS_TOP;
    return code;
}
// int arithmetic of the reference wraps on x86; unsigned arithmetic states that explicitly
__device__ __forceinline__ int wsub(int a, int b) { return (int)((unsigned)a - (unsigned)b); }
__device__ __forceinline__ int wadd(int a, int b) { return (int)((unsigned)a + (unsigned)b); }
__device__ __forceinline__ int wmul(int a, int b) { return (int)((unsigned)a * (unsigned)b); }
__device__ __forceinline__ int wdiv(int a, int b) {
    if (b == 0 || (a == INT32_MIN && b == -1)) return 0;  // x86 would trap; unreachable for finite clips
    return a / b;
}

// Cohen-Sutherland clipLine, util/lineiter.cpp:44-88 (w, h are decremented there: inclusive bounds)
__device__ inline bool clip_line(int &x0, int &y0, int &x1, int &y1, int w, int h) {
    w--;
    h--;
    int oc0 = cs_outcode(x0, y0, w, h), oc1 = cs_outcode(x1, y1, w, h);
    for (int guard = 0;; ++guard) {
        if (!(oc0 | oc1)) return true;
        if (oc0 & oc1) return false;
        if (guard >= 64) return false;
        int x = 0, y = 0;
        const int oc = oc0 ? oc0 : oc1;
        if (oc & CS_TOP) {
            x = wadd(x0, wdiv(wmul(wsub(x1, x0), wsub(h, y0)), wsub(y1, y0)));
            y = h;
        } else if (oc & CS_BOTTOM) {
            x = wadd(x0, wdiv(wmul(wsub(x1, x0), wsub(0, y0)), wsub(y1, y0)));
            y = 0;
        } else if (oc & CS_RIGHT) {
            y = wadd(y0, wdiv(wmul(wsub(y1, y0), wsub(w, x0)), wsub(x1, x0)));
            x = w;
        } else if (oc & CS_LEFT) {
            y = wadd(y0, wdiv(wmul(wsub(y1, y0), wsub(0, x0)), wsub(x1, x0)));
            x = 0;
        }
        if (oc == oc0) {
            x0 = x;
            y0 = y;
            oc0 = cs_outcode(x0, y0, w, h);
        } else {
            x1 = x;
            y1 = y;
            oc1 = cs_outcode(x1, y1, w, h);
        }
    }
}

// HAS_MASK = false: the neighbour was given no mask (every pixel WHITE): the rasteriser's per-step mask
// gather — a dependent global load in a serial loop, the latency that bounds this kernel — is compiled out.
template <bool REFR, bool HAS_MASK>
__global__ void __launch_bounds__(128) curve_build_kernel(const __grid_constant__ CurveArgs a) {
    const int pid = blockIdx.x * blockDim.x + threadIdx.x;
    if (pid >= a.rows * a.w) return;
    const int x = pid % a.w, y = a.row0 + pid / a.w;
    const size_t pix = (size_t)y * a.w + x;
    const size_t plane = (size_t)a.rows * a.w;
    int count = 0;
    if (a.ref_mask[pix] == 255) {
        const size_t n = (size_t)a.w * a.h;
        const d3 src = {a.rays[pix], a.rays[n + pix], a.rays[2 * n + pix]};
        const d3 dir = {a.rays[3 * n + pix], a.rays[4 * n + pix], a.rays[5 * n + pix]};
        RefrProjector pj;
        if (REFR) pj.init(a.nbr, a.Kn, a.prin, a.C, src, dir, a.depth_table, a.fxs, a.cxs, a.fys, a.cys);
        // non-refractive targets: the reference's operations in its own order (cam_project_exact)
        const d3 prin = ld3(a.prin);
        const d3 nrm = normalized(prin);
        const double nd = dot(nrm, dir);
        const bool ray_ok = !(fabs(nd) < 1e-10);
        double r1 = 0.0, r2 = 0.0, r3 = 0.0;  // rho history (warm start of the refractive root)
        int nhist = 0;
        bool have = false;
        double x1 = 0.0, y1 = 0.0;
        int32_t last = TAP_NONE;
        const int cap = a.taps ? a.capacity : 0;
#pragma unroll 1
        for (int d = 0; d < a.D; ++d) {
            double x2, y2;
            bool ok;
            if (REFR) {
                double g = -1.0, rho = 0.0;
                if (nhist >= 3) g = fma(3.0, r1 - r2, r3);
                else if (nhist == 2) g = fma(2.0, r1, -r2);
                else if (nhist == 1) g = r1;
                ok = pj.project(d, g, x2, y2, rho);
                if (ok) {
                    r3 = r2;
                    r2 = r1;
                    r1 = rho;
                    ++nhist;
                } else {
                    nhist = 0;
                }
            } else {
                ok = false;
                const d3 x0 = ld3(a.C) + a.depth_table[d] * prin;
                const double dist = dot(nrm, x0);
                if (ray_ok) {
                    const double t = dot(nrm, dist * nrm - src) / nd;
                    if (!(t < 1e-10)) {
                        double u, v;
                        cam_project_exact(a.nbr, src + t * dir, u, v);
                        x2 = u * a.scale;
                        y2 = v * a.scale;
                        ok = true;
                    }
                }
            }
            if (!ok) continue;
            if (!have) {
                have = true;
                x1 = x2;
                y1 = y2;
                continue;
            }
            const double dx = x2 - x1, dy = y2 - y1;
            if (!(dx * dx + dy * dy >= 1)) continue;
            // LineIterator(x1, y1, x2, y2[, w, h]): the doubles convert to the ctor's int parameters
            int ax0 = to_int_x86(x1), ay0 = to_int_x86(y1), ax1 = to_int_x86(x2), ay1 = to_int_x86(y2);
            x1 = x2;
            y1 = y2;
            if (a.mvs && !clip_line(ax0, ay0, ax1, ay1, a.w, a.h)) continue;
            const bool steep = llabs((long long)ay1 - ay0) > llabs((long long)ax1 - ax0);
            if (steep) {
                int t = ax0; ax0 = ay0; ay0 = t;
                t = ax1; ax1 = ay1; ay1 = t;
            }
            if (ax0 > ax1) {
                int t = ax0; ax0 = ax1; ax1 = t;
                t = ay0; ay0 = ay1; ay1 = t;
            }
            const int deltax = wsub(ax1, ax0), deltay = abs(wsub(ay1, ay0)), ystep = (ay0 < ay1) ? 1 : -1;
            int error = deltax / 2, ly = ay0;
            long long guard = 0;
            for (int lx = ax0; lx <= ax1; ++lx) {
                const int tx = steep ? ly : lx, ty = steep ? lx : ly;
                if (tx >= 0 && ty >= 0 && tx < a.w && ty < a.h && (!HAS_MASK || a.nbr_mask[(size_t)ty * a.w + tx] == 255)) {
                    const int32_t tap = (int32_t)(((uint32_t)ty << 16) | (uint32_t)tx);
                    if (!(a.mvs && tap == last)) {  // multiviewstereo.cpp:800-807
                        if (count < cap) a.taps[(size_t)count * plane + pid] = tap;
                        ++count;
                        last = tap;
                    }
                }
                error -= deltay;
                if (error < 0) {
                    ly += ystep;
                    error += deltax;
                }
                if (lx == INT32_MAX || ++guard > (1LL << 22)) break;  // absurd unclipped segments
            }
        }
    }
    if (a.taps)
        for (int l = min(count, a.capacity); l < a.capacity; ++l) a.taps[(size_t)l * plane + pid] = TAP_NONE;
    if (count > 0) atomicMax(a.max_count, count);  // the longest curve: > capacity means "run again, larger"
}

// Depth hypothesis of a curve candidate (multiviewstereo.cpp:584-593, twoviewstereo.cpp:286-299):
// closest approach of the reference ray and the ray of the candidate's pixel centre in the
// neighbour (Ray3d::closestPoints, util/ray.cpp:53-74), z of the midpoint in the reference frame.
__device__ inline double curve_depth(const double *__restrict__ raysA, const double *__restrict__ raysB, size_t n,
                                     size_t pixA, size_t pixB, const double *R, const double *t) {
    const d3 As = {raysA[pixA], raysA[n + pixA], raysA[2 * n + pixA]};
    const d3 Ad = {raysA[3 * n + pixA], raysA[4 * n + pixA], raysA[5 * n + pixA]};
    const d3 Bs = {raysB[pixB], raysB[n + pixB], raysB[2 * n + pixB]};
    const d3 Bd = {raysB[3 * n + pixB], raysB[4 * n + pixB], raysB[5 * n + pixB]};
    const d3 w0 = As - Bs;
    const double aa = dot(Ad, Ad), bb = dot(Ad, Bd), cc = dot(Bd, Bd), dd = dot(Ad, w0), ee = dot(Bd, w0);
    const double den = 1.0 / (aa * cc - bb * bb);
    const double tl = (bb * ee - cc * dd) * den, tr = (aa * ee - bb * dd) * den;
    d3 p1 = As, p2 = Bs;
    if (tl > 0) p1 = p1 + tl * Ad;
    if (tr > 0) p2 = p2 + tr * Bd;
    p1 = p1 + p2;
    p1 = 0.5 * p1;
    return (mul3(R, p1) + ld3(t)).z;
}

}  // namespace sr
