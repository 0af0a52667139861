// FP32 issue-rate microbenchmark (B200): how many FFMA / FADD / FFMA2 per clock per SM when the
// operands are distinct registers (FFT-like code) vs. reused?
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda.h>

#define ITERS 4096

template <int MODE>
__global__ void __launch_bounds__(256) rate_kernel(float* out, const float* in, long long* cycles) {
    float a[8], x[8], y[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = in[threadIdx.x + i]; x[i] = in[64 + threadIdx.x + i]; y[i] = in[128 + threadIdx.x + i]; }
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) a[i] = fmaf(a[i], x[0], y[0]);                     // FFMA, 2 operands shared by all
            if (MODE == 1) a[i] = fmaf(a[i], x[i], y[(i + 3) & 7]);           // FFMA, all operands distinct
            if (MODE == 2) a[i] = a[i] + x[i];                                // FADD distinct
            if (MODE == 3) a[i] = a[i] * x[i];                                // FMUL distinct
            if (MODE == 4) a[i] = fmaf(a[i], 1.0001f, y[i]);                  // FFMA with immediate
            if (MODE == 5) a[i] = fmaf(x[i], y[(i + 3) & 7], a[(i + 5) & 7]); // FFMA, butterfly-like mixing
        }
    }
    long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

__global__ void __launch_bounds__(256) rate_kernel_packed(float2* out, const float2* in, long long* cycles) {
    float2 a[8], x[8], y[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = in[threadIdx.x + i]; x[i] = in[64 + threadIdx.x + i]; y[i] = in[128 + threadIdx.x + i]; }
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = __ffma2_rn(a[i], x[i], y[(i + 3) & 7]);
    }
    long long t1 = clock64();
    float2 s = make_float2(0, 0);
#pragma unroll
    for (int i = 0; i < 8; ++i) { s.x += a[i].x; s.y += a[i].y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

static double clk_ghz = 1.965;

template <int MODE> void run(const char* name, float* out, float* in, long long* cyc, int warps_per_sm) {
    int blocks = 148 * (warps_per_sm / 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    rate_kernel<MODE><<<blocks, 256>>>(out, in, cyc);
    cudaEventRecord(e0);
    for (int r = 0; r < 5; ++r) rate_kernel<MODE><<<blocks, 256>>>(out, in, cyc);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double ops = 5.0 * ITERS * 8 * 256.0 * blocks;                // lane-ops in total
    printf("%-34s warps/SM %2d: %6.1f lane-ops/clk/SM  (%.3f ms)\n", name, warps_per_sm, ops / (ms * 1e-3) / (clk_ghz * 1e9) / 148.0, ms / 5);
}

int main() {
    float *out, *in; long long* cyc;
    cudaMalloc(&out, 148 * 8 * 256 * sizeof(float2)); cudaMalloc(&in, 4096 * sizeof(float2)); cudaMalloc(&cyc, 4096 * sizeof(long long));
    cudaMemset(in, 0, 4096 * sizeof(float2));
    for (int w : {8, 16, 32, 64}) {
        run<0>("FFMA acc, shared x, shared y", out, in, cyc, w);
        run<1>("FFMA 3 distinct regs", out, in, cyc, w);
        run<5>("FFMA 3 distinct regs (mixing)", out, in, cyc, w);
        run<2>("FADD 2 distinct regs", out, in, cyc, w);
        run<3>("FMUL 2 distinct regs", out, in, cyc, w);
        run<4>("FFMA immediate", out, in, cyc, w);
        int blocks = 148 * (w / 8);
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        rate_kernel_packed<<<blocks, 256>>>((float2*)out, (float2*)in, cyc);
        cudaEventRecord(e0);
        for (int r = 0; r < 5; ++r) rate_kernel_packed<<<blocks, 256>>>((float2*)out, (float2*)in, cyc);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double ops = 5.0 * ITERS * 8 * 256.0 * blocks;
        printf("%-34s warps/SM %2d: %6.1f packed-ops/clk/SM = %6.1f fp32 lane-ops (%.3f ms)\n", "FFMA2 3 distinct pairs", w,
               ops / (ms * 1e-3) / (clk_ghz * 1e9) / 148.0, 2 * ops / (ms * 1e-3) / (clk_ghz * 1e9) / 148.0, ms / 5);
    }
    return 0;
}
