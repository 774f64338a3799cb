// sr_build_refr.cuh — stage (1) for refractive target views: tap-volume build with the
// reprojection organised around the label sweep.
//
// What is computed per (reference pixel, depth label, neighbour view) is Camera::project of
// pointFromDepth (stereo/multiviewstereo.cpp:740-750,768-775; project/camera.cpp:95-138,380-419).
// How: every quantity is an affine function of the ray parameter t(depth) and is hoisted per
// (pixel, view):
//     t      = depth*tA + tB                       (Plane3d(normal, C + normal*depth) + intersect)
//     a      = a0 + t*a1                           (N . local,  local = R*P + t_cam = Ls + t*Ld)
//     radv   = R0 + t*R1                           (local - a*N, the in-plane offset from the axis)
//     Kradv  = KR0 + t*KR1                         (K applied to radv, K normalised when distorted)
// The reference's quartic (camera.cpp:110-116) in the ratio rho = x/r needs only r^2:
//     G(rho) = rho^2 (s^2 rr + hh) - n^2 s^2 (rho^2 rr + dd),   s = 1 - rho,  rr = |radv|^2
// It has exactly one root in [0,1] (the physical one).  rho moves smoothly along
// the depth axis, so the start is the quadratic extrapolation of the three previous labels
// (error ~1e-8) and ONE FP64 Newton step on G lands at ~1e-15; the loop repeats the step until it
// is below 3e-8 (cold start: the first labels of a chunk), a safeguarded bisection/Newton on the
// un-squared equation is the fallback.  No square root, no FP32 excursion, no FP64<->FP32
// conversions (16/clk/SM on B200) on the label path: ~65 FP64 instructions + 2 MUFU per label.
#pragma once
#include "sr_kernels.cuh"

namespace sr {

struct BuildRefrArgs {
    sr_camera nbr;               // target view
    double Kn[9];                // K, or for a distorted view K with rows 0/1 normalised:
                                 //   (K.row0 - cx K.row2)/fx, (K.row1 - cy K.row2)/fy, K.row2
    double fxs, cxs, fys, cys;   // pixel = (fxs*xd + cxs, fys*yd + cys): scale and shift folded in
    double prin[3], C[3];        // reference view principal direction and centre
    const double *rays;          // [6][h][w] of the reference view
    const double *depth_table;   // [D]
    const uint8_t *ref_mask;
    const uint8_t *nbr_mask;
    int32_t *taps;               // [D][rows][w] for this neighbour
    int w, h, row0, rows, D, d_chunk;
    int mvs;
};

__device__ __forceinline__ double rcp_approx(double a) {  // ~2^-20 relative (MUFU.RCP64H)
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a));
    return r;
}

__global__ void __launch_bounds__(128) build_refr_kernel(const __grid_constant__ BuildRefrArgs a) {
    const int pid = blockIdx.x * blockDim.x + threadIdx.x;
    if (pid >= a.rows * a.w) return;
    const int x = pid % a.w, y = a.row0 + pid / a.w;
    const size_t pix = (size_t)y * a.w + x;
    if (a.ref_mask[pix] != 255) return;  // the match kernels never read taps of masked-out pixels
    const size_t n = (size_t)a.w * a.h;
    const d3 src = {a.rays[pix], a.rays[n + pix], a.rays[2 * n + pix]};
    const d3 dir = {a.rays[3 * n + pix], a.rays[4 * n + pix], a.rays[5 * n + pix]};
    // pointFromDepth: plane through C + prin*depth with unit normal nrm; t = (dist*nn - ns)/nd
    const d3 prin = ld3(a.prin);
    const d3 nrm = normalized(prin);
    const double nd = dot(nrm, dir);
    const bool ray_ok = !(fabs(nd) < 1e-10);
    const double inv_nd = 1.0 / nd;
    const double nC = dot(nrm, ld3(a.C)), npn = dot(nrm, prin), nn = dot(nrm, nrm), ns = dot(nrm, src);
    const double tA = npn * nn * inv_nd, tB = (nC * nn - ns) * inv_nd;
    // camera-local ray and its decomposition about the interface normal
    const d3 Ls = fmul3(a.nbr.R, src) + ld3(a.nbr.t), Ld = fmul3(a.nbr.R, dir);
    const d3 N = ld3(a.nbr.plane_n);
    const double a0 = fdot(N, Ls), a1 = fdot(N, Ld);
    const d3 R0 = faxpy(-a0, N, Ls), R1 = faxpy(-a1, N, Ld);
    const d3 KR0 = fmul3(a.Kn, R0), KR1 = fmul3(a.Kn, R1);
    const double pd = a.nbr.plane_d;
    const d3 KdN = fmul3(a.Kn, pd * N);
    const double dd = pd * pd, n1 = a.nbr.n, n2 = n1 * n1;
    const bool distorted = a.nbr.is_distorted != 0;
    const double *k = a.nbr.dist;

    const int d0 = blockIdx.y * a.d_chunk;
    const int d1 = min(d0 + a.d_chunk, a.D);
    const size_t plane = (size_t)a.rows * a.w;
    double r1 = 0.0, r2 = 0.0, r3 = 0.0;  // rho of the previous three labels
    int nhist = 0;
#pragma unroll 1
    for (int d = d0; d < d1; ++d) {
        int32_t tap = TAP_NONE;
        const double t = fma(a.depth_table[d], tA, tB);
        bool ok = ray_ok && !(t < 1e-10);
        double rho = 0.0;
        const d3 radv = faxpy(t, R1, R0);
        const double rr = fdot(radv, radv);
        if (ok) {
            const double av = fma(t, a1, a0);
            const double h = fabs(av) - pd, hh = h * h;
            ok = rr > 0.0;  // on the axis dir = radv/r is NaN in the reference: no root is accepted
            if (ok) {
                if (nhist >= 3) rho = fma(3.0, r1 - r2, r3);
                else if (nhist == 2) rho = fma(2.0, r1, -r2);
                else if (nhist == 1) rho = r1;
                else rho = n1 * fabs(pd) / (fabs(h) + n1 * fabs(pd) + 1e-300);  // paraxial
                rho = fmin(fmax(rho, 0.0), 1.0);
                bool conv = false;
#pragma unroll 1
                for (int it = 0; it < 8; ++it) {
                    const double s = 1.0 - rho;
                    const double p2 = rho * rho, s2 = s * s;
                    const double A = fma(p2, rr, dd), B = fma(s2, rr, hh);
                    const double G = fma(p2, B, -((n2 * s2) * A));
                    const double u = fma(n2, s, rho);
                    // G'/2 = rho*B + n^2 s A - rho s rr (rho + n^2 s)
                    const double g2 = fma(n2 * s, A, fma(rho, B, -(((rho * s) * rr) * u)));
                    const double step = (0.5 * G) * rcp_approx(g2);
                    rho -= step;
                    if (fabs(step) <= 3e-8) {
                        conv = true;
                        break;
                    }
                    rho = fmin(fmax(rho, 0.0), 1.0);
                }
                if (!conv || !(rho >= 0.0 && rho <= 1.0)) {
                    const double r = sqrt(rr);
                    rho = snell_root_robust(r, pd, h, n1, -1.0) / r;
                }
                ok = rho == rho;
            }
        }
        if (ok) {
            r3 = r2;
            r2 = r1;
            r1 = rho;
            ++nhist;
            // point on the interface = rho*radv + d*N (camera.cpp:127); K*point hoisted
            const d3 Kr = faxpy(t, KR1, KR0);
            const d3 p = faxpy(rho, Kr, KdN);
            const double iz = fast_rcp(p.z);
            double xn = p.x * iz, yn = p.y * iz;
            if (distorted) {  // camera.cpp:395-416 on normalised coordinates
                const double q2 = fma(xn, xn, yn * yn);
                const double cdist = fma(fma(fma(k[4], q2, k[1]), q2, k[0]), q2, 1.0);
                const double xo = xn, yo = yn;
                xn = fma(xo, cdist, fma(2 * k[2] * xo, yo, k[3] * fma(2 * xo, xo, q2)));
                // camera.cpp:411-412: y's tangential term uses the already-distorted x
                yn = fma(yo, cdist, fma(k[2], fma(2 * yo, yo, q2), 2 * k[3] * xn * yo));
            }
            int tx = to_int_x86(fma(a.fxs, xn, a.cxs));
            int ty = to_int_x86(fma(a.fys, yn, a.cys));
            bool keep = true;
            if (a.mvs)  // multiviewstereo.cpp:787: only WHITE neighbour-mask pixels are candidates
                keep = tx >= 0 && ty >= 0 && tx < a.w && ty < a.h && a.nbr_mask[(size_t)ty * a.w + tx] == 255;
            if (keep) {
                tx = max(-TAP_CLAMP, min(TAP_CLAMP, tx));
                ty = max(-TAP_CLAMP, min(TAP_CLAMP, ty));
                tap = (int32_t)(((uint32_t)(ty & 0xffff) << 16) | (uint32_t)(tx & 0xffff));
            }
        } else {
            nhist = 0;
        }
        a.taps[(size_t)d * plane + pid] = tap;
    }
}

}  // namespace sr
