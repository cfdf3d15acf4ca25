// dl4ss_linear_fwd : C[M,N] = act(A[M,K] * W[N,K]^T + bias), fp32 on the CUDA cores.
//
// This is the exact-fp32 projection (bit-close to the reference's cuBLAS/MKL SGEMM) used for the
// small contractions (ADDJUST, align attention) and as the numerical yard-stick for the
// tensor-core path (gemm_tc.cu: tcgen05 bf16x3).  128x128x16 tiles, 8x8 register micro-tiles,
// double-buffered shared memory, 128-bit global and shared accesses.
#include "common.cuh"

namespace dl4ss {

constexpr int GBM = 128, GBN = 128, GBK = 16, GPITCH = GBM + 4;

template <int ACT>
__device__ __forceinline__ float apply_act(float x) {
    if (ACT == DL4SS_ACT_TANH) return tanh_f(x);
    if (ACT == DL4SS_ACT_SIGMOID) return sigmoid_f(x);
    return x;
}

template <bool VEC, int ACT>
__global__ void __launch_bounds__(256)
linear_kernel(const float *__restrict__ A, int lda, const float *__restrict__ W, int ldw,
              const float *__restrict__ bias, float *__restrict__ C, int ldc, int M, int N, int K) {
    __shared__ __align__(16) float As[2][GBK][GPITCH];
    __shared__ __align__(16) float Bs[2][GBK][GPITCH];

    const int tid = threadIdx.x;
    const int m0 = blockIdx.x * GBM, n0 = blockIdx.y * GBN;
    const int lrow = tid >> 2, kq = (tid & 3) * 4;
    const int ty = tid >> 4, tx = tid & 15;

    float4 ra[2], rb[2];
    auto load_tile = [&](int k0) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int r = lrow + 64 * i;
            const int gm = m0 + r, gn = n0 + r, gk = k0 + kq;
            if (VEC) {
                ra[i] = (gm < M && gk < K) ? *reinterpret_cast<const float4 *>(A + (size_t)gm * lda + gk)
                                           : make_float4(0.f, 0.f, 0.f, 0.f);
                rb[i] = (gn < N && gk < K) ? *reinterpret_cast<const float4 *>(W + (size_t)gn * ldw + gk)
                                           : make_float4(0.f, 0.f, 0.f, 0.f);
            } else {
                float t[4], u[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    t[j] = (gm < M && gk + j < K) ? A[(size_t)gm * lda + gk + j] : 0.f;
                    u[j] = (gn < N && gk + j < K) ? W[(size_t)gn * ldw + gk + j] : 0.f;
                }
                ra[i] = make_float4(t[0], t[1], t[2], t[3]);
                rb[i] = make_float4(u[0], u[1], u[2], u[3]);
            }
        }
    };
    auto store_tile = [&](int buf) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int r = lrow + 64 * i;
            As[buf][kq + 0][r] = ra[i].x; As[buf][kq + 1][r] = ra[i].y;
            As[buf][kq + 2][r] = ra[i].z; As[buf][kq + 3][r] = ra[i].w;
            Bs[buf][kq + 0][r] = rb[i].x; Bs[buf][kq + 1][r] = rb[i].y;
            Bs[buf][kq + 2][r] = rb[i].z; Bs[buf][kq + 3][r] = rb[i].w;
        }
    };

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    const int nk = (K + GBK - 1) / GBK;
    load_tile(0);
    store_tile(0);
    __syncthreads();
    int buf = 0;
    for (int kt = 0; kt < nk; ++kt) {
        if (kt + 1 < nk) load_tile((kt + 1) * GBK);
#pragma unroll
        for (int k = 0; k < GBK; ++k) {
            const float4 a0 = *reinterpret_cast<const float4 *>(&As[buf][k][ty * 4]);
            const float4 a1 = *reinterpret_cast<const float4 *>(&As[buf][k][64 + ty * 4]);
            const float4 b0 = *reinterpret_cast<const float4 *>(&Bs[buf][k][tx * 4]);
            const float4 b1 = *reinterpret_cast<const float4 *>(&Bs[buf][k][64 + tx * 4]);
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (kt + 1 < nk) store_tile(buf ^ 1);
        __syncthreads();
        buf ^= 1;
    }

    float bv[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int n = n0 + tx * 4 + (j & 3) + 64 * (j >> 2);
        bv[j] = (bias != nullptr && n < N) ? bias[n] : 0.f;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int m = m0 + ty * 4 + (i & 3) + 64 * (i >> 2);
        if (m >= M) continue;
        float *crow = C + (size_t)m * ldc;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int n = n0 + tx * 4 + 64 * h;
            float o[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) o[j] = apply_act<ACT>(acc[i][4 * h + j] + bv[4 * h + j]);
            if (VEC && n + 3 < N) {
                *reinterpret_cast<float4 *>(crow + n) = make_float4(o[0], o[1], o[2], o[3]);
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (n + j < N) crow[n + j] = o[j];
            }
        }
    }
}

template <bool VEC>
static int launch_linear(const float *A, int lda, const float *W, int ldw, const float *bias, float *C,
                         int ldc, int M, int N, int K, int act, cudaStream_t st) {
    dim3 grid(cdiv(M, GBM), cdiv(N, GBN));
    if (act == DL4SS_ACT_NONE)
        linear_kernel<VEC, DL4SS_ACT_NONE><<<grid, 256, 0, st>>>(A, lda, W, ldw, bias, C, ldc, M, N, K);
    else if (act == DL4SS_ACT_TANH)
        linear_kernel<VEC, DL4SS_ACT_TANH><<<grid, 256, 0, st>>>(A, lda, W, ldw, bias, C, ldc, M, N, K);
    else
        linear_kernel<VEC, DL4SS_ACT_SIGMOID><<<grid, 256, 0, st>>>(A, lda, W, ldw, bias, C, ldc, M, N, K);
    DL4SS_LAUNCH_CHECK("linear_kernel");
    return DL4SS_OK;
}

int linear_fwd_impl(const float *A, int lda, const float *W, int ldw, const float *bias, float *C, int ldc,
                    int M, int N, int K, int act, cudaStream_t st) {
    const bool vec = (K % 4 == 0) && (lda % 4 == 0) && (ldw % 4 == 0) && (ldc % 4 == 0) &&
                     (((uintptr_t)A | (uintptr_t)W | (uintptr_t)C) % 16 == 0);
    return vec ? launch_linear<true>(A, lda, W, ldw, bias, C, ldc, M, N, K, act, st)
               : launch_linear<false>(A, lda, W, ldw, bias, C, ldc, M, N, K, act, st);
}

}  // namespace dl4ss

extern "C" int dl4ss_linear_fwd(const float *A, int lda, const float *W, int ldw, const float *bias,
                                float *C, int ldc, int M, int N, int K, int act, void *stream) {
    if (M == 0) return DL4SS_OK;                 // empty batch (pointers may be null)
    DL4SS_CHECK_ARG(A && W && C, "linear_fwd: null operand");
    DL4SS_CHECK_ARG(M >= 0 && N >= 1 && K >= 1, "linear_fwd: bad M/N/K %d/%d/%d", M, N, K);
    DL4SS_CHECK_ARG(lda >= K && ldw >= K && ldc >= N, "linear_fwd: pitch smaller than row");
    DL4SS_CHECK_ARG(act >= DL4SS_ACT_NONE && act <= DL4SS_ACT_SIGMOID, "linear_fwd: bad act %d", act);
    if (M == 0) return DL4SS_OK;
    DL4SS_CHECK_ARG(dl4ss::cdiv(N, dl4ss::GBN) <= 65535, "linear_fwd: N too large for one launch (%d)", N);
    return dl4ss::linear_fwd_impl(A, lda, W, ldw, bias, C, ldc, M, N, K, act, (cudaStream_t)stream);
}
