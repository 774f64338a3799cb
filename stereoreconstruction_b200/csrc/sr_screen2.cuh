// sr_screen2.cuh — the thread-per-pixel (r <= 2) form of the FP32 screen / FP64 verify selection of
// sr_match_screen.cuh, reorganised around a 16-label chunk so that the label loop is nothing but
// the window arithmetic:
//
//   * ONE code path for every pixel.  A reference pixel with inactive taps (outside the image, support
//     weight <= 1e-10: stereo/multiviewstereo.cpp:137-141) carries w = dl = 0 for them, so their
//     t_i = w_i g_i - meanR = -meanR exactly and s3 = sum_all t_i^2 - n_inactive * meanR^2.  The
//     correction is one FFMA for everybody (n_inactive = 0 mostly) instead of a masked variant of the
//     whole window that mixed warps executed IN ADDITION to the plain one.  The cancellation is bounded:
//     the label is FORCEd to FP64 when the correction exceeds 100 x the corrected s3, below that the
//     relative error of s3 is <= 15u * 101 ~ 9e-5, i.e. <= 4.5e-5 of ncc, inside SCREEN_EPS_LOOSE (such
//     pixels never use the tight bar).
//   * The neighbour window is 5 aligned LDG.128 + 5 LDG.32 instead of 25 LDG.32: the screen reads FOUR
//     copies of the FP32 gray plane, copy s shifted right by s floats, and each lane picks the copy in
//     which its window's left edge tx - 2 is 16-byte aligned.
//   * Candidate handling is deferred to the end of the chunk: the loop stores ncc32 to a shared column,
//     raises the lower bound and sets a bit in a per-lane mask; one warp vote per CHUNK (not per label)
//     decides whether anybody has to look at the queue at all.  Queue overflow is resolved there too
//     (flush, then continue with the lane's remaining bits), eviction of entries a risen bound has
//     made hopeless happens when a queue fills and before a flush.
//
// The rule that makes this exact is unchanged (sr_match_screen.cuh): a label is dropped only if its
// upper bound ncc32 + eps is below lower32, a proven lower bound of the winning FP64 cost; every label
// that survives is evaluated by the reference's own two-pass filter in FP64 and the reference's
// selection rule (stereo/multiviewstereo.cpp:589-602,654-660) picks the winner.
#pragma once
#include "sr_match_screen.cuh"

namespace sr {

#ifndef SR_SCREEN2_VEC4
#define SR_SCREEN2_VEC4 1  // 0: A/B against 25 scalar loads from the single FP32 plane
#endif
constexpr float SCREEN_CORR_MAX = 100.0f;  // n_inactive * meanR^2 <= this * s3, else FP64 decides
constexpr float SCREEN_SKIP = -2.0f;       // "no value": below every possible lower bound

// Shared memory of one screening warp.
struct Screen2Smem {
    int32_t tap_ring[2][TAP_CHUNK][32];
    float c32_ring[TAP_CHUNK][32];   // ncc32 + eps of the chunk's labels (SCREEN_FORCE: FP64 decides)
    int32_t q_lab[SCREEN_QCAP][32];  // (eps class << 30) | (neighbour << 16) | label
    int32_t q_tap[SCREEN_QCAP][32];
    float q_c32[SCREEN_QCAP][32];    // upper bound ncc32 + eps; after verification: high word of the FP64 cost
    double px_meanL[32], px_totW[32], px_s2[32], px_bestC[32], px_bestZ[32];
    int px_bestIdx[32];
    unsigned short v_ent[SCREEN_QCAP * 32];
    unsigned char px_flags[32];  // bit 0: all_slow, bit 1: has_inactive
    int st_verified[32], st_viol[32];  // SR_MATCH_STATS only
    float st_maxerr[32];
};

// Register slot i of the window arrays holds window tap k = slot_tap(i) (row-major k).  With the
// vectorised loads a row's first four taps come from one LDG.128 and pair up as (0,1),(2,3); the
// fifth taps of the five rows follow.
template <int R>
__host__ __device__ constexpr int screen2_slot_tap(int i) {
    return (R == 2 && SR_SCREEN2_VEC4) ? (i < 20 ? (i / 4) * 5 + (i % 4) : (i - 20) * 5 + 4) : i;
}

__device__ __forceinline__ int32_t lds_b32(unsigned addr) {
    int32_t v;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts_f32(unsigned addr, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v)); }

// ---- the rare paths: free functions with by-value arguments, so that the screening warp's register
// arrays (weights, dl) never have to be addressable ------------------------------------------------
struct Screen2Cold {
    int qn;
    float lower32;
};

// pixel of lane `o` (clamped into the band: out-of-band lanes shadow the last pixel, no writes)
__device__ __forceinline__ int screen2_pid(const MatchArgs &a, int first_pid, int o) { return min(first_pid + o, a.rows * a.w - 1); }

// exact evaluation of one queued candidate of lane `o` (FP64, the reference's filter)
template <int R>
__device__ __forceinline__ double screen2_exact_cost(const MatchArgs &a, const Screen2Smem &sm, int first_pid, int o, int lab, int tap) {
    const int j = (lab >> 16) & 0xff;
    const int tx = tap & 0xffff, ty = (int)((uint32_t)tap >> 16);
    const int opid = screen2_pid(a, first_pid, o);
    const int ox = opid % a.w, oy = a.row0 + opid / a.w;
    const bool inside = (unsigned)(tx - R) < (unsigned)a.win_w && (unsigned)(ty - R) < (unsigned)a.win_h;
    return (inside && sm.px_flags[o] == 0)
               ? verify_cost_mvs<R, 1>(a, a.grayR[j], ox, oy, tx, ty, opid, 0, 0xffffffffu, sm.px_meanL[o], sm.px_totW[o], sm.px_s2[o])
               : slow_cost<R, 1, SR_COST_NCC_MVS>(a, a.grayR[j], ox, oy, tx, ty, opid, 0, 0xffffffffu);
}

template <bool STATS>
__device__ __forceinline__ void screen2_stats_verified(Screen2Smem &sm, int lane, int o, int q, int lab, double cost) {
    if (!STATS) return;
    ++sm.st_verified[lane];
    if (sm.q_c32[q][o] < 2.0f) {  // |ncc32 - ncc64| relative to its error bar
        const float eb = (lab >> 30) ? SCREEN_EPS_TIGHT : SCREEN_EPS_LOOSE;
        const float c32 = sm.q_c32[q][o] - eb;
        const float err = fabsf((float)(cost - (double)c32));
        sm.st_maxerr[lane] = fmaxf(sm.st_maxerr[lane], err);
        if (err > eb) ++sm.st_viol[lane];
    }
}

// Queued labels whose upper bound has fallen below the (risen) lower bound cannot win.
__device__ __forceinline__ int screen2_evict(Screen2Smem &sm, int lane, int qn, float lower32) {
    int kept = 0;
    for (int q = 0; q < qn; ++q) {
        const float uq = sm.q_c32[q][lane];
        if (uq >= lower32) {
            if (kept != q) {
                sm.q_c32[kept][lane] = uq;
                sm.q_lab[kept][lane] = sm.q_lab[q][lane];
                sm.q_tap[kept][lane] = sm.q_tap[q][lane];
            }
            ++kept;
        }
    }
    return kept;
}

// Verification of the warp's queued candidates; returns the lane's new lower bound (its queue is empty
// afterwards).  Label mode: the entries are compacted into one list and dealt out one per lane, whoever
// owns them; every lane then applies the selection rule (multiviewstereo.cpp:589-602,654-660) to its
// own entries in order.  Curve mode: candidates are (ncc, z) pairs, z = closest approach of the two
// viewing rays (:583-588); each lane walks its own queue.
template <int R, bool STATS>
__device__ __noinline__ float screen2_flush(const MatchArgs &a, Screen2Smem &sm, int lane, int first_pid, int qn, float lower32) {
    constexpr unsigned FULL = 0xffffffffu;
    qn = screen2_evict(sm, lane, qn, lower32);
    double bestC = sm.px_bestC[lane];
    int bestIdx = sm.px_bestIdx[lane];
    if (a.curve) {
        const int w = a.w, h = a.h;
        const int pid = screen2_pid(a, first_pid, lane);
        const size_t pix = (size_t)(a.row0 + pid / w) * w + pid % w;
#pragma unroll 1
        for (int q = 0; q < qn; ++q) {
            const int lab = sm.q_lab[q][lane], tap = sm.q_tap[q][lane];
            const int j = (lab >> 16) & 0xff, d = lab & 0xffff;
            const double cost = screen2_exact_cost<R>(a, sm, first_pid, lane, lab, tap);
            screen2_stats_verified<STATS>(sm, lane, lane, q, lab, cost);
            if (cost > a.ncc_threshold && (bestIdx == SR_INDEX_NONE || cost >= bestC)) {
                const int tx = tap & 0xffff, ty = (int)((uint32_t)tap >> 16);
                const double z = curve_depth(a.raysL, a.raysR[j], (size_t)w * h, pix, (size_t)ty * w + tx, a.camR, a.camT);
                if (bestIdx == SR_INDEX_NONE || cost > bestC || z > sm.px_bestZ[lane]) {
                    bestC = cost;
                    bestIdx = d;
                    sm.px_bestZ[lane] = z;
                }
            }
        }
    } else {
        unsigned short *ent = sm.v_ent;
        int incl = qn;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(FULL, incl, o);
            if (lane >= o) incl += v;
        }
        const int total = __shfl_sync(FULL, incl, 31);
        const int off = incl - qn;
        for (int q = 0; q < qn; ++q) ent[off + q] = (unsigned short)((lane << 8) | q);
        __syncwarp();
#pragma unroll 1
        for (int e = lane; e < total; e += 32) {
            const int o = ent[e] >> 8, q = ent[e] & 0xff;
            const int lab = sm.q_lab[q][o], tap = sm.q_tap[q][o];
            const double cost = screen2_exact_cost<R>(a, sm, first_pid, o, lab, tap);
            screen2_stats_verified<STATS>(sm, lane, o, q, lab, cost);
            sm.q_tap[q][o] = __double2loint(cost);
            sm.q_c32[q][o] = __int_as_float(__double2hiint(cost));
        }
        __syncwarp();
        const bool depth_up = a.depth_up != 0;
        for (int q = 0; q < qn; ++q) {
            const double cost = __hiloint2double(__float_as_int(sm.q_c32[q][lane]), sm.q_tap[q][lane]);
            const int d = sm.q_lab[q][lane] & 0xffff;
            if (cost > a.ncc_threshold) {
                const bool deeper = depth_up ? (d > bestIdx) : (d < bestIdx);
                if (bestIdx == SR_INDEX_NONE || cost > bestC || (cost == bestC && deeper)) {
                    bestC = cost;
                    bestIdx = d;
                }
            }
        }
    }
    sm.px_bestC[lane] = bestC;
    sm.px_bestIdx[lane] = bestIdx;
    // the verified maximum is a valid (and tighter) floor for the screen
    if (bestIdx != SR_INDEX_NONE) lower32 = fmaxf(lower32, (float)bestC - 1e-6f);
    __syncwarp();  // the list and the queues are rewritten by the caller
    return lower32;
}

// End of a chunk: the lanes' marked labels (bits of `pend`, upper bounds in sm.c32_ring) go through the queue; a queue that is full even after eviction makes the warp flush.
template <int R, bool STATS>
__device__ __noinline__ Screen2Cold screen2_process_pending(const MatchArgs &a, Screen2Smem &sm, int lane, int first_pid, int j, int d0,
                                                            int buf, int qn, unsigned pend, unsigned tight, float lower32) {
    constexpr unsigned FULL = 0xffffffffu;
    const bool depth_up = a.depth_up != 0;
#pragma unroll 1
    for (;;) {
#pragma unroll 1
        while (pend) {
            if (qn == SCREEN_QCAP) {
                qn = screen2_evict(sm, lane, qn, lower32);
                if (qn == SCREEN_QCAP) break;  // still full: the warp flushes
            }
            const int l = __ffs(pend) - 1;
            pend &= pend - 1;
            const float ub = sm.c32_ring[l][lane];  // ncc32 + eps, or SCREEN_FORCE
            const bool is_tight = (tight >> l) & 1u;  // (statistics only)
            if (!(ub >= lower32)) continue;  // the bound rose after this label was marked
            const int32_t tap = sm.tap_ring[buf][l][lane];
            const int lab = (is_tight ? (1 << 30) : 0) | (j << 16) | (d0 + l);
            if (qn > 0 && sm.q_tap[qn - 1][lane] == tap && ((sm.q_lab[qn - 1][lane] >> 16) & 0xff) == j) {
                // equal cost by construction: the tie-break picks the deeper label (curve mode: the same
                // pixel is the same (ncc, z) pair, nothing to add)
                if (depth_up && !a.curve) sm.q_lab[qn - 1][lane] = lab;
            } else {
                sm.q_lab[qn][lane] = lab;
                sm.q_tap[qn][lane] = tap;
                sm.q_c32[qn][lane] = ub;
                ++qn;
            }
        }
        if (!__any_sync(FULL, pend != 0u)) break;
        lower32 = screen2_flush<R, STATS>(a, sm, lane, first_pid, qn, lower32);
        qn = 0;
    }
    Screen2Cold r;
    r.qn = qn;
    r.lower32 = lower32;
    return r;
}

// One screening warp: 32 consecutive pixels of the band, one per lane.
template <int R, bool STATS, int PITCH>
struct Screener {
    static constexpr int WS = 2 * R + 1, WN = WS * WS;
    static constexpr bool VEC4 = (R == 2) && (SR_SCREEN2_VEC4 != 0);
    static constexpr unsigned FULL = 0xffffffffu;

    const MatchArgs &a;
    Screen2Smem &sm;
    const int lane;
    float wtf[WN], dlf[WN];
    float inv_totWf, ninact_f, lower32;
    int qn, my_win_w, first_pid;
    unsigned pend, tight;  // tight: SR_MATCH_STATS only
    bool alive, pix_tight;
    int n_forced, n_screened;  // STATS only

    __device__ __forceinline__ Screener(const MatchArgs &a_, Screen2Smem &sm_, int lane_) : a(a_), sm(sm_), lane(lane_) {}

    // ---- per-pixel invariants in FP64, as sr_match_screen.cuh ---------------------------------
    // The warp's pixels are first_pid_ .. first_pid_ + nvalid - 1 of the band (lanes >= nvalid have none).
    __device__ __forceinline__ void init(int first_pid_, int nvalid) {
        first_pid = first_pid_;
        const int w = a.w, h = a.h;
        const int npix_i = a.rows * w;
        const size_t npix = (size_t)npix_i;
        const bool in_band = lane < nvalid;
        const int pid = screen2_pid(a, first_pid, lane);
        const int x = pid % w, y = a.row0 + pid / w;
        const size_t pix = (size_t)y * w + x;
        alive = in_band && a.maskL[pix] == 255;
        if (in_band && !alive) {  // multiviewstereo.cpp:559,565: masked-out pixels stay INF
            a.out_index[pix] = SR_INDEX_MASKED;
            a.out_depth[pix] = dinf();
            a.out_best[pix] = qnan();
        }
        double wt[WN], gl[WN];
        double totW = 0.0, SL = 0.0;
        int ninact = 0;
        // the reference's summation order is row-major k: accumulate in that order, store by slot
#pragma unroll
        for (int k = 0; k < WN; ++k) {
            const int row = k / WS - R, col = k % WS - R;
            const int xl = x + col, yl = y + row;
            double g = qnan(), wv = 0.0;
            if (alive && xl >= 0 && yl >= 0 && xl < w && yl < h) {
                g = a.grayL[(size_t)yl * w + xl];
                wv = a.W[(size_t)k * npix + pid];
            }
            const bool active = (g == g) && (wv > 1e-10);
            wt[k] = active ? wv : 0.0;
            gl[k] = active ? g : 0.0;
            if (active) {
                totW += wv;
                SL += wv * g;
            } else {
                ++ninact;
            }
        }
        const double meanL = SL / totW;
        double s2 = 0.0;
#pragma unroll
        for (int k = 0; k < WN; ++k) {
            const double dl = (wt[k] > 0.0) ? wt[k] * gl[k] - meanL : 0.0;
            s2 += dl * dl;
            gl[k] = dl;
        }
        const bool all_slow = !(totW >= 1e-10) || !(s2 >= (double)WN) || !(s2 < 1e30);
        const double rs2 = all_slow ? 0.0 : 1.0 / sqrt(s2);  // dlf carries 1/sqrt(s2): ncc32 = s1 * rsqrt(s3)
#pragma unroll
        for (int i = 0; i < WN; ++i) {
            const int k = screen2_slot_tap<R>(i);
            wtf[i] = (float)wt[k];
            dlf[i] = (float)(gl[k] * rs2);
        }
        const bool has_inactive = ninact != 0;
        pix_tight = (s2 >= 100.0 * WN) && !has_inactive;
        inv_totWf = (float)(1.0 / totW);
        ninact_f = (float)ninact;
        // lanes that screen nothing (no pixel, or an ill-conditioned reference window that only the exact
        // filter may judge) see an empty interior: every evaluable tap of theirs is FORCEd
        my_win_w = (alive && !all_slow) ? a.win_w : 0;
        sm.px_meanL[lane] = meanL;
        sm.px_totW[lane] = totW;
        sm.px_s2[lane] = s2;
        sm.px_bestC[lane] = 0.0;
        sm.px_bestIdx[lane] = SR_INDEX_NONE;
        sm.px_bestZ[lane] = -1.0;
        sm.px_flags[lane] = (unsigned char)((all_slow ? 1 : 0) | (has_inactive ? 2 : 0));
        if (STATS) {
            sm.st_verified[lane] = sm.st_viol[lane] = 0;
            sm.st_maxerr[lane] = 0.0f;
        }
        asm volatile("" : "+f"(inv_totWf), "+f"(ninact_f));
        qn = 0;
        pend = tight = 0u;
        lower32 = (float)a.ncc_threshold - 1e-6f;
        n_forced = n_screened = 0;
        if (STATS && alive && all_slow && a.stats) atomicAdd(a.stats + 4, 1ull);
        __syncwarp();
    }

    // ---- FP32 screen of one interior tap ---------------------------------------------------------
    // gplane: the neighbour's FP32 gray plane (VEC4: the first of its four shifted copies).
    // Returns the upper bound ncc32 + eps (SCREEN_FORCE when FP64 has to decide) and the lower bound
    // ncc32 - eps (SCREEN_SKIP then); is_tight = the label's error bar is SCREEN_EPS_TIGHT.
    __device__ __forceinline__ float screen_one(const float *__restrict__ gplane, int tx, int ty, float &lb_out, bool &is_tight) const {
        const int fp = PITCH ? PITCH : a.pitch_f;
        float g[WN];
        if (VEC4) {
            // copy s holds pixel x at index x + s; s = (2 - tx) & 3 aligns the window's left edge
            // (e = tx - 2 + s is a multiple of 4); 32-bit index arithmetic: 4 copies < 2^31 floats
            const int s = (2 - tx) & 3;
            const int idx = s * a.plane4_stride + (ty * fp + tx) + s;
            const float *__restrict__ p = gplane + idx;
#pragma unroll
            for (int row = 0; row < WS; ++row) {
                const float4 v = *reinterpret_cast<const float4 *>(p + ((row - R) * fp - 2));
                g[4 * row + 0] = v.x;
                g[4 * row + 1] = v.y;
                g[4 * row + 2] = v.z;
                g[4 * row + 3] = v.w;
                g[20 + row] = p[(row - R) * fp + 2];
            }
        } else {
            const float *__restrict__ base = gplane + (ty * fp + tx);
#pragma unroll
            for (int row = 0; row < WS; ++row)
#pragma unroll
                for (int col = 0; col < WS; ++col) g[row * WS + col] = base[(row - R) * fp + (col - R)];
        }
        constexpr int NP = WN / 2;  // WN is odd: NP pairs + one scalar tap
        float2 S1p[2] = {make_float2(0.0f, 0.0f), make_float2(0.0f, 0.0f)};
#pragma unroll
        for (int p = 0; p < NP; ++p)
            S1p[p & 1] = fma2(make_float2(wtf[2 * p], wtf[2 * p + 1]), make_float2(g[2 * p], g[2 * p + 1]), S1p[p & 1]);
        float S1 = (S1p[0].x + S1p[1].x) + (S1p[0].y + S1p[1].y);
        S1 = fmaf(wtf[WN - 1], g[WN - 1], S1);
        const float mR = S1 * inv_totWf;
        const float2 nm = make_float2(-mR, -mR);
        float2 s3p[2] = {make_float2(0.0f, 0.0f), make_float2(0.0f, 0.0f)};
        float2 s1p[2] = {make_float2(0.0f, 0.0f), make_float2(0.0f, 0.0f)};
#pragma unroll
        for (int p = 0; p < NP; ++p) {
            const float2 t = fma2(make_float2(wtf[2 * p], wtf[2 * p + 1]), make_float2(g[2 * p], g[2 * p + 1]), nm);
            s3p[p & 1] = fma2(t, t, s3p[p & 1]);
            s1p[p & 1] = fma2(make_float2(dlf[2 * p], dlf[2 * p + 1]), t, s1p[p & 1]);
        }
        float s3a = (s3p[0].x + s3p[1].x) + (s3p[0].y + s3p[1].y);
        float s1 = (s1p[0].x + s1p[1].x) + (s1p[0].y + s1p[1].y);
        {
            const float t = fmaf(wtf[WN - 1], g[WN - 1], -mR);
            s3a = fmaf(t, t, s3a);
            s1 = fmaf(dlf[WN - 1], t, s1);
        }
        // inactive taps contributed (-mR)^2 each to s3a and nothing to s1
        const float s3 = fmaf(-ninact_f, mR * mR, s3a);
        is_tight = pix_tight && (s3 >= 100.0f * WN);
        const float eps = is_tight ? SCREEN_EPS_TIGHT : SCREEN_EPS_LOOSE;
        float rs;
        asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(rs) : "f"(s3));
        const float c32 = s1 * rs;
        // ill-conditioned, non-finite or cancellation-dominated neighbour window (correction s3a - s3 above
        // SCREEN_CORR_MAX * s3): FP64 decides
        const bool ok = (s3 >= (float)WN) && (s3 < 1e30f) && (s3a <= (SCREEN_CORR_MAX + 1.0f) * s3);
        lb_out = ok ? c32 - eps : SCREEN_SKIP;
        return ok ? c32 + eps : SCREEN_FORCE;
    }

    // ---- one chunk of labels of neighbour j: taps in sm.tap_ring[buf][0..nl) -----------------------
    __device__ __forceinline__ void chunk(int j, int d0, int nl, int buf, const float *__restrict__ gplane) {
        unsigned ra = (unsigned)__cvta_generic_to_shared(&sm.tap_ring[buf][0][lane]);
        unsigned ca = (unsigned)__cvta_generic_to_shared(&sm.c32_ring[0][lane]);
        unsigned bit = 1u;
        const unsigned win_h = (unsigned)a.win_h;
#pragma unroll 1
        for (int l = 0; l < nl; ++l, ra += 128u, ca += 128u, bit <<= 1) {
            const int32_t tap = lds_b32(ra);
            const int tx = tap & 0xffff, ty = (int)((uint32_t)tap >> 16);  // MVS taps lie inside the image; TAP_NONE -> ty = 32768
            if ((unsigned)(tx - R) < (unsigned)my_win_w && (unsigned)(ty - R) < win_h) {
                float lb;
                bool is_tight;
                const float ub = screen_one(gplane, tx, ty, lb, is_tight);
                sts_f32(ca, ub);
                lower32 = fmaxf(lower32, lb);
                pend |= (ub >= lower32) ? bit : 0u;
                if (STATS) {
                    tight |= is_tight ? bit : 0u;
                    if (ub == SCREEN_FORCE) ++n_forced;
                    else ++n_screened;
                }
            } else if (tap != TAP_NONE && alive) {  // window on the neighbour's border / FP64-only pixel
                sts_f32(ca, SCREEN_FORCE);
                pend |= bit;
                if (STATS) ++n_forced;
            }
        }
        if (__any_sync(FULL, pend != 0u)) {
            const Screen2Cold r = screen2_process_pending<R, STATS>(a, sm, lane, first_pid, j, d0, buf, qn, pend, tight, lower32);
            qn = r.qn;
            lower32 = r.lower32;
            pend = 0u;
        }
        tight = 0u;
    }

    __device__ __forceinline__ void finish() {
        lower32 = screen2_flush<R, STATS>(a, sm, lane, first_pid, qn, lower32);
        qn = 0;
        if (STATS && a.stats && alive) {
            atomicAdd(a.stats + 0, 1ull);
            atomicAdd(a.stats + 1, (unsigned long long)n_screened);
            atomicAdd(a.stats + 2, (unsigned long long)n_forced);
            atomicAdd(a.stats + 6, (unsigned long long)sm.st_viol[lane]);
        }
        if (STATS && a.stats) {  // verifications are counted by the lane that ran them
            atomicAdd(a.stats + 3, (unsigned long long)sm.st_verified[lane]);
            atomicMax(a.stats + 5, (unsigned long long)__float_as_uint(sm.st_maxerr[lane]));  // positive floats order as integers
        }
        if (alive) {
            const int pid = screen2_pid(a, first_pid, lane);
            const size_t pix = (size_t)(a.row0 + pid / a.w) * a.w + pid % a.w;
            const int bestIdx = sm.px_bestIdx[lane];
            a.out_index[pix] = bestIdx;
            a.out_depth[pix] = (bestIdx >= 0) ? (a.curve ? sm.px_bestZ[lane] : a.depth_table[bestIdx]) : -1.0;
            a.out_best[pix] = sm.px_bestC[lane];
        }
    }
};

#ifndef SR_SCREEN2_TILE_ROWS
#define SR_SCREEN2_TILE_ROWS 8  // warps per block: a block owns a 32 x TILE_ROWS tile of reference pixels
#endif
constexpr int SCREEN2_TILE_ROWS = SR_SCREEN2_TILE_ROWS;
constexpr int SCREEN2_WARPS_PER_SM = 16;

// Stand-alone form: the tap volume comes from HBM (label mode: build kernels; curve mode: the curve
// rasteriser) through the two-stage cp.async ring.  The warps of a block are independent (nothing is
// block-wide); they own the rows of a 32-pixel-wide tile so that their neighbour-image footprints overlap
// in L1 (adjacent reference rows project to adjacent neighbour rows).  grid = (ceil(w/32), ceil(rows/TR)).
template <int R, bool STATS, int PITCH>
__global__ void __launch_bounds__(32 * SCREEN2_TILE_ROWS, SCREEN2_WARPS_PER_SM / SCREEN2_TILE_ROWS)
    match_mvs_screen2_kernel(const __grid_constant__ MatchArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    Screen2Smem &sm = reinterpret_cast<Screen2Smem *>(smem_raw)[warp];
    const int x0 = blockIdx.x * 32, row = blockIdx.y * SCREEN2_TILE_ROWS + warp;  // row within the band
    const int nvalid = (row < a.rows) ? min(32, a.w - x0) : 0;
    const int first_pid = min(row, a.rows - 1) * a.w + x0;
    Screener<R, STATS, PITCH> S(a, sm, lane);
    S.init(first_pid, nvalid);
    const int D = a.D;
    const size_t npix = (size_t)a.rows * a.w;
    const int pid = screen2_pid(a, first_pid, lane);
    const uint64_t pol = l2_evict_first_policy();
    // lanes without a pixel never request taps: their columns hold TAP_NONE from here on
    if (!S.alive) {
        for (int l = 0; l < TAP_CHUNK; ++l) sm.tap_ring[0][l][lane] = sm.tap_ring[1][l][lane] = TAP_NONE;
    }
    int jn = 0, dn = 0, bufn = 0;
    auto issue_next = [&]() {
        if (jn < a.num_nbrs && S.alive) {
            const int32_t *src = a.taps + ((size_t)jn * (a.tap_planes ? a.tap_planes : D) + dn) * npix + pid;
            const int nl = min(TAP_CHUNK, D - dn);
            for (int l = 0; l < nl; ++l) cp_async4(&sm.tap_ring[bufn][l][lane], src + (size_t)l * npix, pol);
        }
        cp_async_commit();
        bufn ^= 1;
        dn += TAP_CHUNK;
        if (dn >= D) {
            dn = 0;
            ++jn;
        }
    };
    issue_next();
    int buf = 0;
#pragma unroll 1
    for (int j = 0; j < a.num_nbrs; ++j) {
        const float *__restrict__ gplane = Screener<R, STATS, PITCH>::VEC4 ? a.grayRf4[j] : a.grayRf[j];
#pragma unroll 1
        for (int d0 = 0; d0 < D; d0 += TAP_CHUNK, buf ^= 1) {
            issue_next();
            cp_async_wait<1>();
            S.chunk(j, d0, min(TAP_CHUNK, D - d0), buf, gplane);
        }
    }
    cp_async_wait<0>();
    S.finish();
}

template <int R, bool STATS, int PITCH>
cudaError_t launch_screen2(const MatchArgs &a, cudaStream_t st) {
    auto kern = match_mvs_screen2_kernel<R, STATS, PITCH>;
    constexpr size_t smem = sizeof(Screen2Smem) * SCREEN2_TILE_ROWS;
    if (smem > 48 * 1024) {  // (per device: set on every launch, it is cheap)
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    const dim3 grid((unsigned)((a.w + 31) / 32), (unsigned)((a.rows + SCREEN2_TILE_ROWS - 1) / SCREEN2_TILE_ROWS));
    kern<<<grid, 32 * SCREEN2_TILE_ROWS, smem, st>>>(a);
    return cudaGetLastError();
}

}  // namespace sr
