// Device versions of the ghost.sigtools helpers ("next" rows):
//
//   fastconv_scipy / fastconv_fftw   ghost/sigtools/convolution.py:16-216  -> gcwt_fastconv
//   chirpz_dft                       ghost/sigtools/fourier.py:9-48        -> gcwt_dft
//   analytic_signal_scipy / _fftw    ghost/sigtools/analytic.py:15-112     -> gcwt_analytic_signal
//
// All in complex128 like the reference.  They share the batched power-of-two Stockham FFT of
// the generic CWT path; lengths that are not a power of two go through Bluestein's chirp-z
// identity (what chirpz_dft does on the CPU) with exact integer phase reduction.
#include "plan.h"
#include "common.cuh"
#include "fft_global.cuh"
#include <vector>

namespace gcwt {

typedef double2 C;

// e^{sign * i * pi * j^2 / n} with j^2 reduced mod 2n in integers
__device__ __forceinline__ C chirp(int64_t j, int64_t n, int sign) {
    const int64_t r = (j % (2 * n)) * (j % (2 * n)) % (2 * n);
    C w;
    sincospi((double)sign * (double)r / (double)n, &w.y, &w.x);
    return w;
}

__global__ void st_pad_real(const double* __restrict__ x, int64_t n, C* __restrict__ out, int64_t m) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < m) out[i] = make_double2(i < n ? x[i] : 0.0, 0.0);
}

__global__ void st_pad_complex(const C* __restrict__ x, int64_t n, C* __restrict__ out, int64_t m) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < m) out[i] = i < n ? x[i] : make_double2(0.0, 0.0);
}

__global__ void st_mul(C* __restrict__ a, const C* __restrict__ b, int64_t m, double scale) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < m) {
        const C p = cmul(a[i], b[i]);
        a[i] = make_double2(p.x * scale, p.y * scale);
    }
}

__global__ void st_scale(C* __restrict__ a, int64_t m, double scale) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < m) a[i] = make_double2(a[i].x * scale, a[i].y * scale);
}

__global__ void st_bluestein_pre(const C* __restrict__ x, int64_t n, C* __restrict__ a, int64_t m, int sign) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < m) a[i] = i < n ? cmul(x[i], chirp(i, n, sign)) : make_double2(0.0, 0.0);
}

__global__ void st_bluestein_kernel(C* __restrict__ b, int64_t n, int64_t m, int sign) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    C v = make_double2(0.0, 0.0);
    if (i < n) v = chirp(i, n, -sign);
    else if (m - i < n) v = chirp(m - i, n, -sign);
    b[i] = v;
}

__global__ void st_bluestein_post(const C* __restrict__ conv, int64_t n, C* __restrict__ out, int sign, double scale) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) {
        const C p = cmul(conv[k], chirp(k, n, sign));
        out[k] = make_double2(p.x * scale, p.y * scale);
    }
}

// scipy.signal.hilbert weights: 1 at DC (and Nyquist for even n), 2 for positive, 0 for negative frequencies
__global__ void st_hilbert_weights(C* __restrict__ X, int64_t n) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    double w;
    if (k == 0 || (2 * k == n)) w = 1.0;
    else if (2 * k < n) w = 2.0;
    else w = 0.0;
    X[k] = make_double2(X[k].x * w, X[k].y * w);
}

static inline unsigned nblk(int64_t m) { return (unsigned)((m + 255) / 256); }

// Power-of-two FFT of `m` points: result pointer returned in `res` (one of a / b).
template <int SIGN>
static void fft_pow2(C* a, C* b, int64_t m, C*& res) {
    C* pa = a; C* pb = b;
    if (m > 1) fft_batched<double, SIGN>(pa, pb, (int)m, 1, 0);
    res = pa;
}

// DFT of arbitrary length n (device in -> device out), sign = -1 forward / +1 backward (unscaled).
static int dft_any(const C* d_in, int64_t n, int sign, C* d_out) {
    if (n > (int64_t(1) << 26)) { set_error("sigtools: transform too long"); return GCWT_ERR_UNSUPPORTED; }
    const bool pow2 = (n & (n - 1)) == 0;
    if (pow2) {
        DevBuf t; int rc = t.alloc(sizeof(C) * n); if (rc) return rc;
        GCWT_CUDA_OK(cudaMemcpy(d_out, d_in, sizeof(C) * n, cudaMemcpyDeviceToDevice));
        C* res;
        if (sign < 0) fft_pow2<-1>(d_out, t.c(), n, res); else fft_pow2<+1>(d_out, t.c(), n, res);
        if (res != d_out) GCWT_CUDA_OK(cudaMemcpy(d_out, res, sizeof(C) * n, cudaMemcpyDeviceToDevice));
        return GCWT_OK;
    }
    const int64_t m = int64_t(1) << ilog2_ceil(2 * n - 1);
    DevBuf a, b, t;
    int rc = a.alloc(sizeof(C) * m); if (rc) return rc;
    rc = b.alloc(sizeof(C) * m); if (rc) return rc;
    rc = t.alloc(sizeof(C) * m); if (rc) return rc;
    st_bluestein_pre<<<nblk(m), 256>>>(d_in, n, a.c(), m, sign);
    st_bluestein_kernel<<<nblk(m), 256>>>(b.c(), n, m, sign);
    count_launch(2);
    C *fa, *fb, *fc;
    fft_pow2<-1>(a.c(), t.c(), m, fa);
    C* spare_a = (fa == a.c()) ? t.c() : a.c();
    fft_pow2<-1>(b.c(), spare_a, m, fb);
    C* spare_b = (fb == b.c()) ? spare_a : b.c();
    st_mul<<<nblk(m), 256>>>(fa, fb, m, 1.0 / (double)m);
    count_launch();
    fft_pow2<+1>(fa, spare_b, m, fc);
    st_bluestein_post<<<nblk(n), 256>>>(fc, n, d_out, sign, 1.0);
    count_launch();
    GCWT_CUDA_OK(cudaGetLastError());
    GCWT_CUDA_OK(cudaDeviceSynchronize());
    return GCWT_OK;
}

int sig_dft_host(const double* x_host, int64_t n, int sign, double* out_host) {
    DevBuf in, out;
    int rc = in.alloc(sizeof(C) * n); if (rc) return rc;
    rc = out.alloc(sizeof(C) * n); if (rc) return rc;
    GCWT_CUDA_OK(cudaMemcpy(in.p, x_host, sizeof(C) * n, cudaMemcpyHostToDevice));
    rc = dft_any(in.c(), n, sign, out.c()); if (rc) return rc;
    GCWT_CUDA_OK(cudaMemcpy(out_host, out.p, sizeof(C) * n, cudaMemcpyDeviceToHost));
    return GCWT_OK;
}

int sig_analytic_host(const double* x_host, int64_t n, double* out_host) {
    DevBuf xr, a, b;
    int rc = xr.alloc(sizeof(double) * n); if (rc) return rc;
    rc = a.alloc(sizeof(C) * n); if (rc) return rc;
    rc = b.alloc(sizeof(C) * n); if (rc) return rc;
    GCWT_CUDA_OK(cudaMemcpy(xr.p, x_host, sizeof(double) * n, cudaMemcpyHostToDevice));
    st_pad_real<<<nblk(n), 256>>>((const double*)xr.p, n, a.c(), n);
    count_launch();
    rc = dft_any(a.c(), n, -1, b.c()); if (rc) return rc;
    st_hilbert_weights<<<nblk(n), 256>>>(b.c(), n);
    count_launch();
    rc = dft_any(b.c(), n, +1, a.c()); if (rc) return rc;
    st_scale<<<nblk(n), 256>>>(a.c(), n, 1.0 / (double)n);
    count_launch();
    GCWT_CUDA_OK(cudaGetLastError());
    GCWT_CUDA_OK(cudaMemcpy(out_host, a.p, sizeof(C) * n, cudaMemcpyDeviceToHost));
    return GCWT_OK;
}

// Full linear convolution (n + m - 1 complex outputs) of complex signal and kernel.
int sig_fastconv_host(const double* sig_host, int sig_complex, int64_t n, const double* ker_host, int ker_complex,
                      int64_t m, double* out_host) {
    const int64_t tot = n + m - 1;
    if (tot > (int64_t(1) << 26)) { set_error("fastconv: n + m - 1 above 2^26, convolve in blocks"); return GCWT_ERR_UNSUPPORTED; }
    const int64_t nfft = int64_t(1) << ilog2_ceil(tot);
    DevBuf raw_s, raw_k, a, b, t;
    int rc = raw_s.alloc(sizeof(double) * n * (sig_complex ? 2 : 1)); if (rc) return rc;
    rc = raw_k.alloc(sizeof(double) * m * (ker_complex ? 2 : 1)); if (rc) return rc;
    rc = a.alloc(sizeof(C) * nfft); if (rc) return rc;
    rc = b.alloc(sizeof(C) * nfft); if (rc) return rc;
    rc = t.alloc(sizeof(C) * nfft); if (rc) return rc;
    GCWT_CUDA_OK(cudaMemcpy(raw_s.p, sig_host, sizeof(double) * n * (sig_complex ? 2 : 1), cudaMemcpyHostToDevice));
    GCWT_CUDA_OK(cudaMemcpy(raw_k.p, ker_host, sizeof(double) * m * (ker_complex ? 2 : 1), cudaMemcpyHostToDevice));
    if (sig_complex) st_pad_complex<<<nblk(nfft), 256>>>((const C*)raw_s.p, n, a.c(), nfft);
    else st_pad_real<<<nblk(nfft), 256>>>((const double*)raw_s.p, n, a.c(), nfft);
    if (ker_complex) st_pad_complex<<<nblk(nfft), 256>>>((const C*)raw_k.p, m, b.c(), nfft);
    else st_pad_real<<<nblk(nfft), 256>>>((const double*)raw_k.p, m, b.c(), nfft);
    count_launch(2);
    C *fa, *fb, *fc;
    fft_pow2<-1>(a.c(), t.c(), nfft, fa);
    C* spare_a = (fa == a.c()) ? t.c() : a.c();
    fft_pow2<-1>(b.c(), spare_a, nfft, fb);
    C* spare_b = (fb == b.c()) ? spare_a : b.c();
    st_mul<<<nblk(nfft), 256>>>(fa, fb, nfft, 1.0 / (double)nfft);
    count_launch();
    fft_pow2<+1>(fa, spare_b, nfft, fc);
    GCWT_CUDA_OK(cudaGetLastError());
    GCWT_CUDA_OK(cudaMemcpy(out_host, fc, sizeof(C) * tot, cudaMemcpyDeviceToHost));
    return GCWT_OK;
}

// ---------------------------------------------------------------------------- moments
// sum and sum of squares (of x, or of x^2 when `square`) in fp64: the global mean / std that
// plot(standardize=True) needs (ghost/wave/transforms.py:360-366) without a host pass.
template <typename T>
__global__ void moments_kernel(const T* __restrict__ x, int64_t n, int square, double* __restrict__ acc) {
    double s1 = 0.0, s2 = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double v = (double)x[i];
        if (square) v *= v;
        s1 += v;
        s2 += v * v;
    }
    __shared__ double sh1[256], sh2[256];
    sh1[threadIdx.x] = s1; sh2[threadIdx.x] = s2;
    __syncthreads();
    for (int k = 128; k > 0; k >>= 1) {
        if ((int)threadIdx.x < k) { sh1[threadIdx.x] += sh1[threadIdx.x + k]; sh2[threadIdx.x] += sh2[threadIdx.x + k]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { atomicAdd(acc, sh1[0]); atomicAdd(acc + 1, sh2[0]); }
}

int sig_moments(const void* x_dev, int type, int64_t n, int square, double* out_host, cudaStream_t st) {
    DevBuf buf;
    { int rc = buf.alloc(2 * sizeof(double)); if (rc) return rc; }
    double* d = buf.as<double>();
    GCWT_CUDA_OK(cudaMemsetAsync(d, 0, 2 * sizeof(double), st));
    const unsigned blocks = (unsigned)std::min<int64_t>(148 * 8, (n + 255) / 256);
    if (type == GCWT_F32) moments_kernel<float><<<blocks, 256, 0, st>>>((const float*)x_dev, n, square, d);
    else moments_kernel<double><<<blocks, 256, 0, st>>>((const double*)x_dev, n, square, d);
    count_launch();
    GCWT_CUDA_OK(cudaGetLastError());
    double h[2];
    GCWT_CUDA_OK(cudaMemcpyAsync(h, d, sizeof(h), cudaMemcpyDeviceToHost, st));
    GCWT_CUDA_OK(cudaStreamSynchronize(st));
    const double mean = h[0] / (double)n;
    double var = h[1] / (double)n - mean * mean;
    if (var < 0.0) var = 0.0;
    out_host[0] = mean;
    out_host[1] = sqrt(var);
    return GCWT_OK;
}

// ----------------------------------------------------------------------------- pooling for display
// plot() draws contourf over the whole (scales, samples) array (ghost/wave/transforms.py:356-367,395-396);
// a display has a few thousand columns.  One warp reduces one bin of `width` consecutive samples of one
// row to its mean or maximum (of x, or of x^2 for power from amplitude), accumulating in fp64: the
// consumer then pulls (rows x bins) values over PCIe instead of 4 bytes per coefficient.
template <typename T>
__global__ void __launch_bounds__(256)
pool_rows_kernel(const T* __restrict__ x, int64_t row_stride, int64_t n_cols, int64_t width, int mode, int square,
                 double* __restrict__ out, int64_t out_stride, int64_t n_bins) {
    const int64_t bin = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (bin >= n_bins) return;
    const int64_t lo = bin * width, hi = min(n_cols, lo + width);
    const T* row = x + (int64_t)blockIdx.y * row_stride;
    double acc = mode == 1 ? -1.0e300 : 0.0;
    for (int64_t i = lo + lane; i < hi; i += 32) {
        double v = (double)row[i];
        if (square) v *= v;
        acc = mode == 1 ? fmax(acc, v) : acc + v;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double other = __shfl_xor_sync(0xffffffffu, acc, o);
        acc = mode == 1 ? fmax(acc, other) : acc + other;
    }
    if (lane == 0) out[(int64_t)blockIdx.y * out_stride + bin] = mode == 1 ? acc : acc / (double)(hi - lo);
}

int pool_rows_launch(const void* x_dev, int type, int64_t n_rows, int64_t n_cols, int64_t row_stride, int64_t width,
                     int mode, int square, double* out_dev, int64_t out_stride, cudaStream_t st) {
    if (n_rows <= 0 || n_cols <= 0 || width <= 0 || (mode != 0 && mode != 1)) { set_error("pool_rows: bad argument"); return GCWT_ERR_ARG; }
    const int64_t n_bins = (n_cols + width - 1) / width;
    for (int64_t r0 = 0; r0 < n_rows; r0 += 65535) {                 // gridDim.y limit
        const int64_t rows = std::min<int64_t>(65535, n_rows - r0);
        dim3 grid((unsigned)((n_bins + 7) / 8), (unsigned)rows);
        if (type == GCWT_F32)
            pool_rows_kernel<float><<<grid, 256, 0, st>>>((const float*)x_dev + r0 * row_stride, row_stride, n_cols, width, mode,
                                                          square, out_dev + r0 * out_stride, out_stride, n_bins);
        else
            pool_rows_kernel<double><<<grid, 256, 0, st>>>((const double*)x_dev + r0 * row_stride, row_stride, n_cols, width,
                                                           mode, square, out_dev + r0 * out_stride, out_stride, n_bins);
        count_launch();
    }
    GCWT_CUDA_OK(cudaGetLastError());
    return GCWT_OK;
}

}  // namespace gcwt
