// sr_pipeline.cuh — stage (1) and stages (2)+(3) of the MultiViewStereo label path as ONE launch.
//
// The refractive tap build is FP64/integer work (Newton on the quartic, guarded cubic interpolation:
// FP64 pipe ~55 % busy, FMA pipe idle); the screen is FFMA2/LSU work (FMA pipe the busiest unit, FP64
// pipe idle).  Run back to back each leaves the other's pipes empty, and two streams do not mix them
// (the block scheduler drains one grid before it places the next).  Here both are roles of one kernel:
//
//   * A block draws a ticket (one atomic) and derives its role from it.  Tickets come in groups of
//     num_nbrs + 1: ticket (g, j < num_nbrs) BUILDS tile g of the reference image against neighbour j,
//     ticket (g, num_nbrs) SCREENS tile g - lag.  A tile is 32 x SCREEN2_TILE_ROWS reference pixels, one
//     warp per row, in both roles.  The block scheduler hands tickets out in order, so every SM holds
//     a mix of build and screen blocks in proportion to their running times, and their instruction
//     streams share the SM's pipes.
//   * The taps of a tile travel through a ring in global memory, [slot][neighbour][row][D][32]: a warp
//     writes 128 contiguous bytes per label, and the screening warp of the same row later streams its
//     labels as ONE contiguous run through the cp.async ring (16-byte L1-bypassing copies: the slot may
//     have held another tile's taps in this SM's L1).  With a short lag the ring is L2-resident: the
//     tap volume no longer makes the 2 x 4 B per (pixel, label, neighbour) round trip through HBM.
//   * Ordering: a screen block waits until built[tile] == num_nbrs (release: __threadfence + barrier +
//     atomicAdd; acquire: ld.acquire.gpu by one thread + barrier); a build block that re-uses a ring
//     slot waits for done[tile - ring_tiles].  A block only ever waits for blocks with SMALLER tickets
//     (lag >= 1, ring_tiles > lag), which are resident or finished: no deadlock, whatever the order in
//     which the hardware starts blocks.
//
// The arithmetic is the stand-alone kernels' (build_refr_sweep, Screener): outputs are bit-identical.
#pragma once
#include "sr_build_refr.cuh"
#include "sr_screen2.cuh"

namespace sr {

struct PipeNbr {
    sr_camera cam;         // target view
    double Kn[9];          // see BuildRefrArgs::Kn
    double fxs, cxs, fys, cys;
    const uint8_t *mask;   // null: every pixel of the neighbour is WHITE
};

struct PipeArgs {
    MatchArgs m;                 // (taps / tap_planes unused)
    PipeNbr nb[SR_MAX_NBRS];
    double prin[3], C[3];        // reference view principal direction and centre
    const double *rays;          // [6][h][w] of the reference view
    int32_t *ring;               // [ring_tiles][num_nbrs][TILE_ROWS][D][32]
    int *built;                  // [ntiles] finished build blocks of the tile (zeroed per launch)
    int *done;                   // [ntiles] 1: the tile's screen block has consumed its taps
    unsigned *ticket;            // zeroed per launch
    int tiles_x, ntiles, ring_tiles, lag;
    unsigned long long *check;   // SR_MATCH_STATS: build self-check counters
};

__device__ __forceinline__ int ld_acquire_gpu(const int *p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(int *p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void cp_async16_cg(void *smem_dst, const void *gmem_src, uint64_t pol) {
    asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)),
                 "l"(gmem_src), "l"(pol));
}

template <int R, bool STATS, int PITCH>
__global__ void __launch_bounds__(32 * SCREEN2_TILE_ROWS, SR_SCREEN2_MINBLOCKS) mvs_pipeline_kernel(const __grid_constant__ PipeArgs a) {
    constexpr int TR = SCREEN2_TILE_ROWS;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ unsigned s_ticket;
    __shared__ double s_uniform[RefrProjector::UC_COUNT];
    if (threadIdx.x == 0) s_ticket = atomicAdd(a.ticket, 1u);
    __syncthreads();
    const int nn = a.m.num_nbrs;
    const int group = (int)(s_ticket / (unsigned)(nn + 1)), slot = (int)(s_ticket % (unsigned)(nn + 1));
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int D = a.m.D, w = a.m.w;
    const size_t row_ints = (size_t)D * 32;  // one warp row of one (tile, neighbour): D labels x 32 lanes

    if (slot < nn) {
        // ---------------------------------------------------------------- build tile `group` against neighbour `slot`
        const int t = group;
        if (t >= a.ntiles) return;
        if (threadIdx.x == 0) {
            const PipeNbr &nb0 = a.nb[slot];
            RefrProjector::fill_uniform(s_uniform, nb0.cam, nb0.Kn, nb0.fxs, nb0.cxs, nb0.fys, nb0.cys);
            // the ring slot still holds tile t - ring_tiles until its screen block is through
            if (t >= a.ring_tiles)
                while (ld_acquire_gpu(a.done + (t - a.ring_tiles)) == 0) __nanosleep(256);
        }
        __syncthreads();
        const int x = (t % a.tiles_x) * 32 + lane, row = (t / a.tiles_x) * TR + warp;  // row within the band
        if (x < w && row < a.m.rows) {
            const int y = a.m.row0 + row;
            if (a.m.maskL[(size_t)y * w + x] == 255) {  // (the screen never reads taps of masked-out pixels)
                const PipeNbr &nb = a.nb[slot];
                BuildSweep s;
                s.nbr = &nb.cam;
                s.Kn = nb.Kn;
                s.fxs = nb.fxs;
                s.cxs = nb.cxs;
                s.fys = nb.fys;
                s.cys = nb.cys;
                s.prin = a.prin;
                s.C = a.C;
                s.rays = a.rays;
                s.depth_table = a.m.depth_table;
                s.nbr_mask = nb.mask;
                s.w = w;
                s.h = a.m.h;
                s.D = D;
                s.check = a.check;
                s.uniform = s_uniform;
                int32_t *__restrict__ out = a.ring + ((size_t)((t % a.ring_tiles) * nn + slot) * TR + warp) * row_ints + lane;
                auto sink = [&](int d, int32_t tap) { out[(size_t)d * 32] = tap; };
                if (nb.mask) build_refr_sweep<true, true, true>(s, x, y, 0, D, sink);
                else build_refr_sweep<true, false, true>(s, x, y, 0, D, sink);
            }
        }
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) atomicAdd(a.built + t, 1);
        return;
    }

    // -------------------------------------------------------------------- screen tile `group - lag`
    const int t = group - a.lag;
    if (t < 0 || t >= a.ntiles) return;
    if (threadIdx.x == 0)
        while (ld_acquire_gpu(a.built + t) < nn) __nanosleep(256);
    __syncthreads();
    {
        Screen2Smem<STATS> &sm = reinterpret_cast<Screen2Smem<STATS> *>(smem_raw)[warp];
        const int x0 = (t % a.tiles_x) * 32, row = (t / a.tiles_x) * TR + warp;
        const int nvalid = (row < a.m.rows) ? min(32, w - x0) : 0;
        const int first_pid = min(row, a.m.rows - 1) * w + x0;
        Screener<R, STATS, PITCH> S(a.m, sm, lane);
        S.init(first_pid, nvalid);
        const uint64_t pol = l2_evict_first_policy();
        // this warp's taps: neighbour j, labels d0.. = one contiguous run of 128-byte rows
        const int32_t *__restrict__ src0 = a.ring + ((size_t)((t % a.ring_tiles) * nn) * TR + warp) * row_ints;
        int jn = 0, dn = 0, bufn = 0;
        auto issue_next = [&]() {
            if (jn < nn) {
                const int4 *src = reinterpret_cast<const int4 *>(src0 + (size_t)jn * TR * row_ints + (size_t)dn * 32);
                int4 *dst = reinterpret_cast<int4 *>(&sm.tap_ring[bufn][0][0]);
                const int pieces = min(TAP_CHUNK, D - dn) * 8;  // 16-byte pieces
                for (int p = lane; p < pieces; p += 32) cp_async16_cg(dst + p, src + p, pol);
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
        for (int j = 0; j < nn; ++j) {
            const float *__restrict__ gplane = a.m.grayRf[j];
#pragma unroll 1
            for (int d0 = 0; d0 < D; d0 += TAP_CHUNK, buf ^= 1) {
                issue_next();
                cp_async_wait<1>();
                __syncwarp();  // the chunk was fetched by all lanes together
                S.chunk(j, d0, min(TAP_CHUNK, D - d0), buf, gplane);
                __syncwarp();  // ... and nobody refills a buffer a slower lane still reads
            }
        }
        cp_async_wait<0>();
        S.finish();
    }
    __syncthreads();
    if (threadIdx.x == 0) st_release_gpu(a.done + t, 1);
}

template <int R, bool STATS, int PITCH>
cudaError_t launch_pipeline_rp(const PipeArgs &a, cudaStream_t st) {
    auto kern = mvs_pipeline_kernel<R, STATS, PITCH>;
    constexpr size_t smem = sizeof(Screen2Smem<STATS>) * SCREEN2_TILE_ROWS;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    const unsigned grid = (unsigned)(a.ntiles + a.lag) * (unsigned)(a.m.num_nbrs + 1);
    kern<<<grid, 32 * SCREEN2_TILE_ROWS, smem, st>>>(a);
    return cudaGetLastError();
}

template <int R>
cudaError_t launch_pipeline_r(const PipeArgs &a, cudaStream_t st) {
    if (a.m.stats) return launch_pipeline_rp<R, true, 0>(a, st);
    switch (a.m.pitch_f) {
        case 1024: return launch_pipeline_rp<R, false, 1024>(a, st);
        case 2048: return launch_pipeline_rp<R, false, 2048>(a, st);
        case 4096: return launch_pipeline_rp<R, false, 4096>(a, st);
        default: return launch_pipeline_rp<R, false, 0>(a, st);
    }
}

inline bool pipeline_supported(int radius) { return radius == 1 || radius == 2; }
inline cudaError_t launch_pipeline(int radius, const PipeArgs &a, cudaStream_t st) {
    return radius == 1 ? launch_pipeline_r<1>(a, st) : launch_pipeline_r<2>(a, st);
}

}  // namespace sr
