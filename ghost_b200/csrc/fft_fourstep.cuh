// fp64 overlap-save through a four-step FFT whose sub-transforms run in shared memory.
//
// An n-point transform (n = N1 * N2 = 2^11 ... 2^20, N1, N2 <= 1024) is two kernels instead of one launch
// per radix pass: transforms of length N1 down the columns of the N1 x N2 matrix, a twiddle, transforms of
// length N2 along the rows.  The forward transform leaves the spectrum in "transposed" order (bin
// k1 + N1 k2 at position k1 N2 + k2); the filter responses are tabulated in that order and the inverse
// runs rows first, columns second, which returns natural order -- no transposition pass anywhere.  The
// multiplication by the response is fused into the load of the inverse row pass and the epilogue
// (complex / |W| / |W|^2, overlap-save discard, cast to the output type) into the store of the column
// pass, so one (channel, scale) costs: read spectrum + response, write and re-read one intermediate,
// write the coefficients.
#pragma once
#include "plan.h"
#include "common.cuh"

namespace gcwt {

constexpr int kTile = 2048;          // complex128 elements per block tile (32 KB; two tiles ping-pong)
constexpr int kTilePad = 64;         // columns are stored N1 + 1 apart (bank spread): up to 64 columns of 32
constexpr size_t kFourStepSmem = sizeof(double2) * 2 * (kTile + kTilePad);

struct FourStep {
    int p = 0, p1 = 0, p2 = 0;       // n = 2^p, N1 = 2^p1 (columns), N2 = 2^p2 (rows, contiguous)
    int n1 = 0, n2 = 0;
    int rows_per_block = 0;          // kTile / N2
    int cols_per_block = 0;          // kTile / N1
    // (a block tile of kTile elements must not exceed the transform: at n = 1024 it would hold two of them)
    static bool usable(int64_t n) { return n >= kTile && n <= (int64_t(1) << 20) && (n & (n - 1)) == 0; }
    explicit FourStep(int64_t n) {
        p = ilog2_ceil(n);
        p1 = p / 2; p2 = p - p1;
        n1 = 1 << p1; n2 = 1 << p2;
        rows_per_block = kTile / n2; cols_per_block = kTile / n1;
    }
};

__device__ __forceinline__ double2 cmuld(double2 a, double2 b) {
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// e^{SIGN * 2 pi i j / 1024} from the table of e^{-2 pi i j / 1024}
template <int SIGN>
__device__ __forceinline__ double2 tw1k(const double2* __restrict__ tw, int j) {
    const double2 t = __ldg(tw + (j & 1023));
    return SIGN < 0 ? t : make_double2(t.x, -t.y);
}

// e^{SIGN * 2 pi i idx / n}, idx < n <= 2^20: product of a coarse (1024 / (n / 1024) spaced) and a fine table entry
template <int SIGN>
__device__ __forceinline__ double2 tw_big(const double2* __restrict__ tw, const double2* __restrict__ tw_fine, int p, int idx) {
    if (p <= 10) return tw1k<SIGN>(tw, idx << (10 - p));
    const int hi = idx >> 10, lo = idx & 1023;
    const double2 a = tw1k<SIGN>(tw, hi << (20 - p));           // e^{2 pi i hi 1024 / n}
    double2 b = __ldg(tw_fine + lo);                            // e^{-2 pi i lo / n}
    if (SIGN > 0) b.y = -b.y;
    return cmuld(a, b);
}

// `count` transforms of length 2^q, transform f occupying a[f * stride .. + 2^q): radix-4 Stockham passes
// (plus one radix-2 pass when q is odd) between the two tile buffers; returns the buffer holding the result.
template <int SIGN>
__device__ __forceinline__ double2* tile_fft(double2* a, double2* b, int q, int count, int stride,
                                             const double2* __restrict__ tw) {
    const int M = 1 << q;
    int ns = 1, lg = 0;
    for (; lg + 2 <= q; lg += 2) {
        const int quarter = M >> 2, total = count * quarter;
        for (int u = threadIdx.x; u < total; u += blockDim.x) {
            const int f = u >> (q - 2), j = u & (quarter - 1), k = j & (ns - 1);
            const double2* src = a + f * stride + j;
            double2 v0 = src[0], v1 = src[quarter], v2 = src[2 * quarter], v3 = src[3 * quarter];
            if (ns > 1) {
                const int idx = k << (8 - lg);                    // k * 1024 / (4 ns)
                v1 = cmuld(v1, tw1k<SIGN>(tw, idx));
                v2 = cmuld(v2, tw1k<SIGN>(tw, 2 * idx));
                v3 = cmuld(v3, tw1k<SIGN>(tw, 3 * idx));
            }
            dft4<SIGN>(v0, v1, v2, v3);
            double2* dst = b + f * stride + ((j - k) << 2) + k;
            dst[0] = v0; dst[ns] = v1; dst[2 * ns] = v2; dst[3 * ns] = v3;
        }
        __syncthreads();
        double2* t = a; a = b; b = t;
        ns <<= 2;
    }
    if (q & 1) {
        const int half = M >> 1, total = count * half;
        for (int u = threadIdx.x; u < total; u += blockDim.x) {
            const int f = u >> (q - 1), j = u & (half - 1), k = j & (ns - 1);
            const double2* src = a + f * stride + j;
            double2 v0 = src[0], v1 = src[half];
            if (ns > 1) v1 = cmuld(v1, tw1k<SIGN>(tw, k << (9 - lg)));   // k * 1024 / (2 ns)
            double2* dst = b + f * stride + ((j - k) << 1) + k;
            dst[0] = make_double2(v0.x + v1.x, v0.y + v1.y);
            dst[ns] = make_double2(v0.x - v1.x, v0.y - v1.y);
        }
        __syncthreads();
        double2* t = a; a = b; b = t;
    }
    return a;
}

struct FourStepParams {
    int p, p1, p2;
    const double2* tw;        // e^{-2 pi i j / 1024}
    const double2* tw_fine;   // e^{-2 pi i j / n}, j < 1024
    int n_items;              // (channel, chunk) items in this batch
    int64_t item0, n_chunks, hop, offset, n, halo_l, halo_r;
};

// ---- forward, step 1: columns of the zero-padded, mean-removed chunk, then the twiddle -------------
template <typename TIn>
__global__ void __launch_bounds__(256)
fs_forward_cols_kernel(const FourStepParams fp, const TIn* __restrict__ x, int64_t x_stride,
                       const double* __restrict__ means, double2* __restrict__ U) {
    extern __shared__ __align__(16) unsigned char fs_smem[];
    double2* a = (double2*)fs_smem;
    double2* b = a + kTile + kTilePad;
    const int n1 = 1 << fp.p1, n2 = 1 << fp.p2, cw = kTile >> fp.p1, lcw = 11 - fp.p1;
    const int item = blockIdx.x % fp.n_items, c0 = (blockIdx.x / fp.n_items) * cw;
    const int64_t gi = fp.item0 + item, c = gi / fp.n_chunks, q = gi % fp.n_chunks;
    const int64_t t0 = q * fp.hop - fp.offset;
    const TIn* xc = x + c * x_stride;
    const double mu = means[c];
    for (int e = threadIdx.x; e < kTile; e += 256) {
        const int r = e >> lcw, dc = e & (cw - 1);                // row n1-index, column within the block
        const int64_t t = t0 + (int64_t)r * n2 + c0 + dc;
        double v = 0.0;
        if (t >= -fp.halo_l && t < fp.n + fp.halo_r) v = (double)xc[t] - mu;
        a[dc * (n1 + 1) + r] = make_double2(v, 0.0);
    }
    __syncthreads();
    double2* res = tile_fft<-1>(a, b, fp.p1, cw, n1 + 1, fp.tw);
    double2* Ui = U + ((int64_t)item << fp.p);
    for (int e = threadIdx.x; e < kTile; e += 256) {
        const int k1 = e >> lcw, dc = e & (cw - 1), col = c0 + dc;
        const int idx = (int)(((int64_t)col * k1) & ((int64_t(1) << fp.p) - 1));
        Ui[(int64_t)k1 * n2 + col] = cmuld(res[dc * (n1 + 1) + k1], tw_big<-1>(fp.tw, fp.tw_fine, fp.p, idx));
    }
}

// ---- forward, step 2: rows; the spectrum stays in transposed order (bin k1 + N1 k2 at k1 N2 + k2) ---
__global__ void __launch_bounds__(256)
fs_forward_rows_kernel(const FourStepParams fp, const double2* __restrict__ U, double2* __restrict__ Y) {
    extern __shared__ __align__(16) unsigned char fs_smem[];
    double2* a = (double2*)fs_smem;
    double2* b = a + kTile + kTilePad;
    const int n2 = 1 << fp.p2, rb = kTile >> fp.p2;
    const int item = blockIdx.x % fp.n_items, r0 = (blockIdx.x / fp.n_items) * rb;
    const int64_t base = ((int64_t)item << fp.p) + (int64_t)r0 * n2;
    for (int e = threadIdx.x; e < kTile; e += 256) a[e] = U[base + e];
    __syncthreads();
    double2* res = tile_fft<-1>(a, b, fp.p2, rb, n2, fp.tw);
    for (int e = threadIdx.x; e < kTile; e += 256) Y[base + e] = res[e];
}

// ---- inverse, step 1: (spectrum x response) rows, then the twiddle --------------------------------
__global__ void __launch_bounds__(256)
fs_inverse_rows_kernel(const FourStepParams fp, const double2* __restrict__ Y, const double2* __restrict__ H,
                       double2* __restrict__ T, int n_scales) {
    extern __shared__ __align__(16) unsigned char fs_smem[];
    double2* a = (double2*)fs_smem;
    double2* b = a + kTile + kTilePad;
    const int n2 = 1 << fp.p2, rb = kTile >> fp.p2;
    // items vary fastest: the response rows of a scale are shared by consecutive blocks, the spectra of a batch
    // of items stay in L2 across the scales
    const int item = blockIdx.x % fp.n_items, r0 = (blockIdx.x / fp.n_items) * rb, s = blockIdx.y;
    const int64_t row0 = (int64_t)r0 * n2;
    const double2* Yi = Y + ((int64_t)item << fp.p) + row0;
    const double2* Hs = H + ((int64_t)s << fp.p) + row0;
    for (int e = threadIdx.x; e < kTile; e += 256) a[e] = cmuld(Yi[e], __ldg(Hs + e));
    __syncthreads();
    double2* res = tile_fft<+1>(a, b, fp.p2, rb, n2, fp.tw);
    double2* Ti = T + (((int64_t)item * n_scales + s) << fp.p) + row0;
    for (int e = threadIdx.x; e < kTile; e += 256) {
        const int k1 = r0 + (e >> fp.p2), col = e & (n2 - 1);
        const int idx = (int)(((int64_t)col * k1) & ((int64_t(1) << fp.p) - 1));
        Ti[e] = cmuld(res[e], tw_big<+1>(fp.tw, fp.tw_fine, fp.p, idx));
    }
}

// ---- inverse, step 2: columns, natural order out, epilogue fused into the store ---------------------
template <typename TOut, int KIND>
__global__ void __launch_bounds__(256)
fs_inverse_cols_kernel(const FourStepParams fp, const double2* __restrict__ T, int n_scales,
                       const int* __restrict__ ids, void* out, int64_t s_stride, int64_t c_stride) {
    extern __shared__ __align__(16) unsigned char fs_smem[];
    double2* a = (double2*)fs_smem;
    double2* b = a + kTile + kTilePad;
    const int n1 = 1 << fp.p1, n2 = 1 << fp.p2, cw = kTile >> fp.p1, lcw = 11 - fp.p1;
    const int item = blockIdx.x % fp.n_items, c0 = (blockIdx.x / fp.n_items) * cw, s = blockIdx.y;
    const double2* Ti = T + (((int64_t)item * n_scales + s) << fp.p);
    for (int e = threadIdx.x; e < kTile; e += 256) {
        const int k1 = e >> lcw, dc = e & (cw - 1);
        a[dc * (n1 + 1) + k1] = Ti[(int64_t)k1 * n2 + c0 + dc];
    }
    __syncthreads();
    double2* res = tile_fft<+1>(a, b, fp.p1, cw, n1 + 1, fp.tw);
    const int64_t gi = fp.item0 + item, c = gi / fp.n_chunks, q = gi % fp.n_chunks;
    const int64_t obase = c * c_stride + (int64_t)ids[s] * s_stride + q * fp.hop;
    typedef typename cplx_of<TOut>::type CO;
    for (int e = threadIdx.x; e < kTile; e += 256) {
        const int r = e >> lcw, dc = e & (cw - 1);
        const int64_t i = (int64_t)r * n2 + c0 + dc - fp.offset;  // owned sample of this chunk?
        if (i < 0 || i >= fp.hop || q * fp.hop + i >= fp.n) continue;
        const double2 v = res[dc * (n1 + 1) + r];
        if (KIND == GCWT_OUT_COMPLEX) ((CO*)out)[obase + i] = mk<TOut>((TOut)v.x, (TOut)v.y);
        else if (KIND == GCWT_OUT_AMPLITUDE) ((TOut*)out)[obase + i] = (TOut)sqrt(v.x * v.x + v.y * v.y);
        else ((TOut*)out)[obase + i] = (TOut)(v.x * v.x + v.y * v.y);
    }
}

__global__ void fs_fine_twiddle_kernel(int p, double2* __restrict__ tw_fine) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= 1024) return;
    double s, c;
    sincospi(-2.0 * (double)j / (double)(int64_t(1) << p), &s, &c);
    tw_fine[j] = make_double2(c, s);
}

}  // namespace gcwt
