// n2  dl4ss_xcorr_f64 : batched short-lag cross-correlations in fp64, the data-dependent part of BSS-Eval
// (Torch_multi/bss_test.py:55 -> mir_eval.separation.bss_eval_sources: `_compute_reference_correlations`,
// `_compute_projection_filters`).  mir_eval gets the 512-lag correlations from 2^17-point FFTs; here they are
// summed directly in double (x, y are the fp32 waveforms the path produced, every product is exact in fp64):
//     out[b, i, j, k] = sum_m x[b, i, m] * y[b, j, m + lag0 + k],   k in [0, nlags),  y = 0 outside [0, N)
// One CTA owns 64 lags of one (b, i, j) and walks the whole signal in 2048-sample chunks staged in shared memory;
// a warp = 8 consecutive lags, a lane = an interleaved set of 4-sample strips: per strip 4 LDS.128 feed 32 DFMAs.
#include "common.cuh"

namespace dl4ss {

constexpr int XC_THREADS = 256;
constexpr int XC_LAGS = 64;          // lags per CTA (8 per warp)
constexpr int XC_CHUNK = 2048;       // samples of x per stage

__global__ void __launch_bounds__(XC_THREADS)
xcorr_f64_kernel(const float *__restrict__ x, const float *__restrict__ y, int Sx, int Sy, int N, int nlags,
                 int lag0, double *__restrict__ out) {
    __shared__ __align__(16) float xs[XC_CHUNK];
    __shared__ __align__(16) float ys[XC_CHUNK + XC_LAGS + 8];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int k0 = blockIdx.x * XC_LAGS;
    const int pair = blockIdx.y, i = pair / Sy, j = pair - i * Sy, b = blockIdx.z;
    const float *xp = x + ((size_t)b * Sx + i) * N;
    const float *yp = y + ((size_t)b * Sy + j) * N;
    const int kw = 8 * warp;                      // this warp's lags: k0 + kw .. +7

    double acc[8];
#pragma unroll
    for (int a = 0; a < 8; ++a) acc[a] = 0.0;

    for (int m0 = 0; m0 < N; m0 += XC_CHUNK) {
        __syncthreads();
        for (int t = tid; t < XC_CHUNK; t += XC_THREADS) {
            const int m = m0 + t;
            xs[t] = (m < N) ? xp[m] : 0.f;
        }
        for (int t = tid; t < XC_CHUNK + XC_LAGS + 8; t += XC_THREADS) {
            const int m = m0 + lag0 + k0 + t;
            ys[t] = (m >= 0 && m < N) ? yp[m] : 0.f;
        }
        __syncthreads();
        // strip s: samples 4s..4s+3 of the chunk; needs y[4s + kw .. 4s + kw + 10]
        for (int s = lane; s < XC_CHUNK / 4; s += 32) {
            const float4 xv = *reinterpret_cast<const float4 *>(xs + 4 * s);
            const float4 y0 = *reinterpret_cast<const float4 *>(ys + 4 * s + kw);
            const float4 y1 = *reinterpret_cast<const float4 *>(ys + 4 * s + kw + 4);
            const float4 y2 = *reinterpret_cast<const float4 *>(ys + 4 * s + kw + 8);
            const double xd[4] = {(double)xv.x, (double)xv.y, (double)xv.z, (double)xv.w};
            const double yd[12] = {(double)y0.x, (double)y0.y, (double)y0.z, (double)y0.w, (double)y1.x, (double)y1.y,
                                   (double)y1.z, (double)y1.w, (double)y2.x, (double)y2.y, (double)y2.z, (double)y2.w};
#pragma unroll
            for (int a = 0; a < 8; ++a)
#pragma unroll
                for (int c = 0; c < 4; ++c) acc[a] = fma(xd[c], yd[a + c], acc[a]);
        }
    }
#pragma unroll
    for (int a = 0; a < 8; ++a) {
        double v = acc[a];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        const int k = k0 + kw + a;
        if (lane == 0 && k < nlags) out[(((size_t)b * Sx + i) * Sy + j) * nlags + k] = v;
    }
}

}  // namespace dl4ss

using namespace dl4ss;

extern "C" int dl4ss_xcorr_f64(const float *x, const float *y, int B, int Sx, int Sy, int N, int nlags, int lag0,
                               double *out, void *stream) {
    if (B == 0 || Sx == 0 || Sy == 0 || nlags == 0) return DL4SS_OK;
    DL4SS_CHECK_ARG(x && y && out, "xcorr_f64: null pointer");
    DL4SS_CHECK_ARG(B > 0 && Sx > 0 && Sy > 0 && N > 0 && nlags > 0, "xcorr_f64: bad sizes B=%d Sx=%d Sy=%d N=%d nlags=%d",
                    B, Sx, Sy, N, nlags);
    DL4SS_CHECK_ARG(B <= 65535 && Sx * Sy <= 65535, "xcorr_f64: grid too large (B=%d pairs=%d)", B, Sx * Sy);
    dim3 grid(cdiv(nlags, XC_LAGS), Sx * Sy, B);
    xcorr_f64_kernel<<<grid, XC_THREADS, 0, (cudaStream_t)stream>>>(x, y, Sx, Sy, N, nlags, lag0, out);
    DL4SS_LAUNCH_CHECK("xcorr_f64_kernel");
    return DL4SS_OK;
}
