// microbench.cu — calibrates the FP64/FP32/LSU issue rates of the B200 that bound the
// dense-matching kernels (DESIGN.md "rooflines").  nvcc -O3 -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cuda_runtime.h>

#define ITER 4096
template <int OP>
__global__ void k(double *out, double a, double b, int n) {
    double x0 = threadIdx.x * 1e-3 + 1.0, x1 = x0 + 0.1, x2 = x0 + 0.2, x3 = x0 + 0.3, x4 = x0 + 0.4, x5 = x0 + 0.5,
           x6 = x0 + 0.6, x7 = x0 + 0.7;
    for (int i = 0; i < n; ++i) {
        if (OP == 0) {  // DFMA
            x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
            x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
        } else if (OP == 1) {  // rsqrt
            x0 = rsqrt(x0 + a); x1 = rsqrt(x1 + a); x2 = rsqrt(x2 + a); x3 = rsqrt(x3 + a);
            x4 = rsqrt(x4 + a); x5 = rsqrt(x5 + a); x6 = rsqrt(x6 + a); x7 = rsqrt(x7 + a);
        } else if (OP == 2) {  // div
            x0 = a / (x0 + b); x1 = a / (x1 + b); x2 = a / (x2 + b); x3 = a / (x3 + b);
            x4 = a / (x4 + b); x5 = a / (x5 + b); x6 = a / (x6 + b); x7 = a / (x7 + b);
        } else if (OP == 3) {  // sqrt
            x0 = sqrt(x0 + a); x1 = sqrt(x1 + a); x2 = sqrt(x2 + a); x3 = sqrt(x3 + a);
            x4 = sqrt(x4 + a); x5 = sqrt(x5 + a); x6 = sqrt(x6 + a); x7 = sqrt(x7 + a);
        } else if (OP == 4) {  // double -> float -> double conversions
            x0 = (double)((float)x0) + a; x1 = (double)((float)x1) + a; x2 = (double)((float)x2) + a; x3 = (double)((float)x3) + a;
            x4 = (double)((float)x4) + a; x5 = (double)((float)x5) + a; x6 = (double)((float)x6) + a; x7 = (double)((float)x7) + a;
        } else if (OP == 5) {  // double -> int (trunc) -> double
            x0 = (double)__double2int_rz(x0) + a; x1 = (double)__double2int_rz(x1) + a; x2 = (double)__double2int_rz(x2) + a; x3 = (double)__double2int_rz(x3) + a;
            x4 = (double)__double2int_rz(x4) + a; x5 = (double)__double2int_rz(x5) + a; x6 = (double)__double2int_rz(x6) + a; x7 = (double)__double2int_rz(x7) + a;
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}
__global__ void kf(float *out, float a, float b, int n) {  // FFMA
    float x0 = threadIdx.x * 1e-3f + 1.0f, x1 = x0 + 0.1f, x2 = x0 + 0.2f, x3 = x0 + 0.3f, x4 = x0 + 0.4f, x5 = x0 + 0.5f, x6 = x0 + 0.6f, x7 = x0 + 0.7f;
    for (int i = 0; i < n; ++i) {
        x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
        x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}
// gather: each lane reads 25 doubles (5x5 window) from an L1/L2-resident plane
template <typename T>
__global__ void kg(const T *plane, int w, int h, double *out, int n) {
    int x = 8 + (blockIdx.x * blockDim.x + threadIdx.x) % (w - 16), y = 8 + (blockIdx.x % (h - 16));
    double acc = 0;
    for (int i = 0; i < n; ++i) {
        const T *b = plane + (size_t)(y + (i & 3)) * w + x + (i & 7);
#pragma unroll
        for (int r = -2; r <= 2; ++r)
#pragma unroll
            for (int c = -2; c <= 2; ++c) acc += (double)b[r * w + c];
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
// dependent-issue latency: one warp per SM, one chain
__global__ void klat(double *out, long long *cyc, double a, double b, int n, int op) {
    double x = threadIdx.x * 1e-3 + 1.0;
    long long t0 = clock64();
    if (op == 0) for (int i = 0; i < n; ++i) x = fma(x, a, b);
    else if (op == 1) for (int i = 0; i < n; ++i) x = x + a;
    else for (int i = 0; i < n; ++i) x = x * a;
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = x;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
// dependent chains with W warps per SMSP and C chains per thread: FP64 issue rate vs parallelism
template <int C>
__global__ void kchains(double *out, double a, double b, int n) {
    double x[C];
#pragma unroll
    for (int c = 0; c < C; ++c) x[c] = threadIdx.x * 1e-3 + c;
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int c = 0; c < C; ++c) x[c] = fma(x[c], a, b);
    }
    double s = 0;
#pragma unroll
    for (int c = 0; c < C; ++c) s += x[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <typename F>
float timeit(F f) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); return ms;
}
int main() {
    const int blocks = 148 * 8, threads = 256;
    double *out; cudaMalloc(&out, blocks * threads * 8);
    const double nops = (double)blocks * threads * ITER * 8;
    const char *names[] = {"DFMA", "rsqrt(double)", "div(double)", "sqrt(double)", "F2F f64<->f32 pair + DADD", "F2I+I2F f64 pair + DADD"};
    float ms;
    ms = timeit([&] { k<0><<<blocks, threads>>>(out, 1.0000001, 1e-9, ITER); }); printf("%-28s %8.3f ms  %8.2f Gop/s\n", names[0], ms, nops / ms / 1e6);
    ms = timeit([&] { k<1><<<blocks, threads>>>(out, 1.0000001, 1e-9, ITER); }); printf("%-28s %8.3f ms  %8.2f Gop/s\n", names[1], ms, nops / ms / 1e6);
    ms = timeit([&] { k<2><<<blocks, threads>>>(out, 1.0000001, 1e-9, ITER); }); printf("%-28s %8.3f ms  %8.2f Gop/s\n", names[2], ms, nops / ms / 1e6);
    ms = timeit([&] { k<3><<<blocks, threads>>>(out, 1.0000001, 1e-9, ITER); }); printf("%-28s %8.3f ms  %8.2f Gop/s\n", names[3], ms, nops / ms / 1e6);
    ms = timeit([&] { k<4><<<blocks, threads>>>(out, 1.0000001, 1e-9, ITER); }); printf("%-28s %8.3f ms  %8.2f Gop/s\n", names[4], ms, nops / ms / 1e6);
    ms = timeit([&] { k<5><<<blocks, threads>>>(out, 1.0000001, 1e-9, ITER); }); printf("%-28s %8.3f ms  %8.2f Gop/s\n", names[5], ms, nops / ms / 1e6);
    ms = timeit([&] { kf<<<blocks, threads>>>((float *)out, 1.0000001f, 1e-9f, ITER); }); printf("%-28s %8.3f ms  %8.2f Gop/s\n", "FFMA", ms, nops / ms / 1e6);
    long long *cyc; cudaMallocManaged(&cyc, 8);
    for (int op = 0; op < 3; ++op) {
        klat<<<1, 32>>>(out, cyc, 1.0000001, 1e-9, 8192, op); cudaDeviceSynchronize();
        printf("dependent %s latency: %.2f cycles\n", op == 0 ? "DFMA" : op == 1 ? "DADD" : "DMUL", (double)*cyc / 8192);
    }
    {
        const int n = 8192;
        float m;
        m = timeit([&] { kchains<1><<<148 * 2, 128>>>(out, 1.0000001, 1e-9, n); }); printf("8 warps/SM x 1 chain : %7.2f Gop/s\n", 148.0 * 2 * 128 * n * 1 / m / 1e6);
        m = timeit([&] { kchains<2><<<148 * 2, 128>>>(out, 1.0000001, 1e-9, n); }); printf("8 warps/SM x 2 chains: %7.2f Gop/s\n", 148.0 * 2 * 128 * n * 2 / m / 1e6);
        m = timeit([&] { kchains<4><<<148 * 2, 128>>>(out, 1.0000001, 1e-9, n); }); printf("8 warps/SM x 4 chains: %7.2f Gop/s\n", 148.0 * 2 * 128 * n * 4 / m / 1e6);
        m = timeit([&] { kchains<8><<<148 * 2, 128>>>(out, 1.0000001, 1e-9, n); }); printf("8 warps/SM x 8 chains: %7.2f Gop/s\n", 148.0 * 2 * 128 * n * 8 / m / 1e6);
        m = timeit([&] { kchains<4><<<148 * 4, 128>>>(out, 1.0000001, 1e-9, n); }); printf("16 warps/SM x 4 chains: %7.2f Gop/s\n", 148.0 * 4 * 128 * n * 4 / m / 1e6);
    }
    const int w = 1920, h = 1080;
    double *pd; float *pf; unsigned short *ps;
    cudaMalloc(&pd, w * h * 8); cudaMalloc(&pf, w * h * 4); cudaMalloc(&ps, w * h * 2);
    cudaMemset(pd, 0, w * h * 8); cudaMemset(pf, 0, w * h * 4); cudaMemset(ps, 0, w * h * 2);
    const int gi = 256; const double nl = (double)blocks * threads * gi * 25;
    ms = timeit([&] { kg<double><<<blocks, threads>>>(pd, w, h, out, gi); }); printf("%-28s %8.3f ms  %8.2f Gload/s\n", "5x5 gather f64 (+conv,DADD)", ms, nl / ms / 1e6);
    ms = timeit([&] { kg<float><<<blocks, threads>>>(pf, w, h, out, gi); }); printf("%-28s %8.3f ms  %8.2f Gload/s\n", "5x5 gather f32 (+F2F,DADD)", ms, nl / ms / 1e6);
    ms = timeit([&] { kg<unsigned short><<<blocks, threads>>>(ps, w, h, out, gi); }); printf("%-28s %8.3f ms  %8.2f Gload/s\n", "5x5 gather u16 (+I2F,DADD)", ms, nl / ms / 1e6);
    return 0;
}
