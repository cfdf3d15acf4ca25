// n4  Discriminator forward (TDAA_beta/main_run_sstune_EvalVer.py:328-346): three 3x3 stride-2 convolutions with ReLU
// (1 -> 64 -> 64 -> 64 channels, no padding) over a [N, 1, T, F] spectrogram, flatten, Linear(36480 -> 1), sigmoid.
// The reference runs cuDNN convolutions + cuBLAS; here:
//   dl4ss_conv3x3s2_relu_fwd : direct fp32 convolution (exact-order-free fp32 FMA; parity 1e-5 against torch on the CPU).
//       Cin = 1  : HBM bound (writes 64 channels per pixel): one thread per output pixel, weights in shared memory, the
//                  64 channel planes written with coalesced rows.
//       Cin > 1  : CTA = (sample, 4 output rows, all columns, 64 output channels); input channels are walked in chunks
//                  of 16 whose 9-row input patch and 64 x 16 x 9 weights sit in shared memory; a thread owns 4 output
//                  channels (16 apart) x 8 output pixels (32 accumulators: 12 shared-memory loads feed 32 FMAs per tap).
//   dl4ss_rowdot_sigmoid_fwd : out[n] = sigmoid(<x[n,:], w> + b), one CTA per sample (the 36480-wide final layer).
#include "common.cuh"

namespace dl4ss {

constexpr int CV_THREADS = 256;
constexpr int CV_CO = 64;            // output channels (all layers of the reference's discriminator)
constexpr int CV_CC = 16;            // input channels per shared-memory chunk
constexpr int CV_TH = 4;             // output rows per CTA
constexpr int CV_MAXOW = 32;         // output columns the tiled kernel holds (layer 2: 31, layer 3: 15)
constexpr int CV_WP = CV_CC * 9 + 1; // weight row pitch in floats (odd: the 16 channel groups of a warp hit distinct banks)

// ---- Cin == 1
__global__ void __launch_bounds__(CV_THREADS)
conv3x3s2_c1_kernel(const float *__restrict__ x, const float *__restrict__ w, const float *__restrict__ b,
                    float *__restrict__ y, int IH, int IW, int OH, int OW, int Cout) {
    extern __shared__ float ws[];            // [Cout][9] + [Cout]
    for (int i = threadIdx.x; i < Cout * 9; i += blockDim.x) ws[i] = w[i];
    for (int i = threadIdx.x; i < Cout; i += blockDim.x) ws[Cout * 9 + i] = b ? b[i] : 0.f;
    __syncthreads();
    const int n = blockIdx.y;
    const int pix = blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= OH * OW) return;
    const int oy = pix / OW, ox = pix - oy * OW;
    const float *xp = x + ((size_t)n * IH + 2 * oy) * IW + 2 * ox;
    float v[9];
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) v[ky * 3 + kx] = __ldg(xp + ky * IW + kx);
    float *yp = y + (size_t)n * Cout * OH * OW + pix;
    for (int co = 0; co < Cout; ++co) {
        float a = ws[Cout * 9 + co];
#pragma unroll
        for (int k = 0; k < 9; ++k) a = fmaf(ws[co * 9 + k], v[k], a);
        yp[(size_t)co * OH * OW] = fmaxf(a, 0.f);
    }
}

// ---- Cin a multiple of CV_CC, Cout == CV_CO, OW <= CV_MAXOW
__global__ void __launch_bounds__(CV_THREADS)
conv3x3s2_tiled_kernel(const float *__restrict__ x, const float *__restrict__ w, const float *__restrict__ b,
                       float *__restrict__ y, int Cin, int IH, int IW, int OH, int OW) {
    extern __shared__ float sm[];
    const int PR = 2 * CV_TH + 1;                       // input rows of the patch
    const int PW = 2 * CV_MAXOW + 2;                    // patch row pitch (columns 0 .. 2*OW, zero beyond the image)
    float *xs = sm;                                     // [CV_CC][PR][PW]
    float *wsm = sm + CV_CC * PR * PW;                  // [CV_CO][CV_WP]
    const int tid = threadIdx.x;
    const int n = blockIdx.y, oy0 = blockIdx.x * CV_TH;
    const int cg = tid & 15;                            // output channels cg, cg+16, cg+32, cg+48 (odd row pitch: no bank conflicts)
    const int pg = tid >> 4;                            // pixel group: row pg / 4 of the tile, columns 8*(pg%4) .. +7
    const int prow = pg >> 2, pcol0 = (pg & 3) * 8;
    float acc[4][8];
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int q = 0; q < 8; ++q) acc[c][q] = 0.f;

    for (int c0 = 0; c0 < Cin; c0 += CV_CC) {
        __syncthreads();
        for (int i = tid; i < CV_CC * PR * PW; i += CV_THREADS) {
            const int col = i % PW, r = (i / PW) % PR, ci = i / (PW * PR);
            const int iy = 2 * oy0 + r;
            xs[i] = (iy < IH && col < IW) ? __ldg(x + (((size_t)n * Cin + c0 + ci) * IH + iy) * IW + col) : 0.f;
        }
        for (int i = tid; i < CV_CO * CV_CC * 9; i += CV_THREADS) {
            const int k = i % (CV_CC * 9), co = i / (CV_CC * 9);
            wsm[co * CV_WP + k] = __ldg(w + ((size_t)co * Cin + c0) * 9 + k);
        }
        __syncthreads();
        for (int ci = 0; ci < CV_CC; ++ci) {
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
                const float *xr = xs + (ci * PR + 2 * prow + ky) * PW + 2 * pcol0;
                float xv[17];
#pragma unroll
                for (int q = 0; q < 17; ++q) xv[q] = xr[q];
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    float wv[4];
#pragma unroll
                    for (int c = 0; c < 4; ++c) wv[c] = wsm[(cg + 16 * c) * CV_WP + ci * 9 + ky * 3 + kx];
#pragma unroll
                    for (int c = 0; c < 4; ++c)
#pragma unroll
                        for (int q = 0; q < 8; ++q) acc[c][q] = fmaf(wv[c], xv[2 * q + kx], acc[c][q]);
                }
            }
        }
    }
    const int oy = oy0 + prow;
    if (oy < OH) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int co = cg + 16 * c;
            const float bias = b ? __ldg(b + co) : 0.f;
            float *yp = y + (((size_t)n * CV_CO + co) * OH + oy) * OW;
#pragma unroll
            for (int q = 0; q < 8; ++q)
                if (pcol0 + q < OW) yp[pcol0 + q] = fmaxf(acc[c][q] + bias, 0.f);
        }
    }
}

__global__ void __launch_bounds__(256)
rowdot_sigmoid_kernel(const float *__restrict__ x, const float *__restrict__ w, const float *__restrict__ b,
                      float *__restrict__ out, int K) {
    __shared__ float part[8];
    const float *xp = x + (size_t)blockIdx.x * K;
    float a = 0.f;
    for (int k = threadIdx.x; k < K; k += 256) a = fmaf(__ldg(xp + k), __ldg(w + k), a);
    a = warp_sum(a);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = a;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = b ? b[0] : 0.f;
        for (int i = 0; i < 8; ++i) s += part[i];
        out[blockIdx.x] = 1.0f / (1.0f + expf(-s));
    }
}

}  // namespace dl4ss

using namespace dl4ss;

extern "C" int dl4ss_conv3x3s2_relu_fwd(const float *x, const float *w, const float *b, float *y, int N, int Cin, int IH,
                                        int IW, int Cout, void *stream) {
    DL4SS_CHECK_ARG(x && w && y, "conv3x3s2_relu_fwd: null operand");
    DL4SS_CHECK_ARG(N >= 0 && Cin >= 1 && Cout >= 1 && IH >= 3 && IW >= 3, "conv3x3s2_relu_fwd: bad shape N=%d Cin=%d Cout=%d %dx%d",
                    N, Cin, Cout, IH, IW);
    if (N == 0) return DL4SS_OK;
    DL4SS_CHECK_ARG(N <= 65535, "conv3x3s2_relu_fwd: at most 65535 samples per call");
    const int OH = (IH - 3) / 2 + 1, OW = (IW - 3) / 2 + 1;
    cudaStream_t st = (cudaStream_t)stream;
    if (Cin == 1) {
        dim3 grid(cdiv(OH * OW, CV_THREADS), N);
        conv3x3s2_c1_kernel<<<grid, CV_THREADS, (size_t)Cout * 10 * sizeof(float), st>>>(x, w, b, y, IH, IW, OH, OW, Cout);
        DL4SS_LAUNCH_CHECK("conv3x3s2_c1_kernel");
        return DL4SS_OK;
    }
    if (Cin % CV_CC != 0 || Cout != CV_CO || OW > CV_MAXOW) {
        set_error("conv3x3s2_relu_fwd: Cin=%d Cout=%d OW=%d unsupported (Cin = 1, or a multiple of %d with Cout = %d and at most %d "
                  "output columns)", Cin, Cout, OW, CV_CC, CV_CO, CV_MAXOW);
        return DL4SS_EUNSUPPORTED;
    }
    const size_t smem = ((size_t)CV_CC * (2 * CV_TH + 1) * (2 * CV_MAXOW + 2) + (size_t)CV_CO * CV_WP) * sizeof(float);
    DL4SS_CUDA(cudaFuncSetAttribute(conv3x3s2_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(cdiv(OH, CV_TH), N);
    conv3x3s2_tiled_kernel<<<grid, CV_THREADS, smem, st>>>(x, w, b, y, Cin, IH, IW, OH, OW);
    DL4SS_LAUNCH_CHECK("conv3x3s2_tiled_kernel");
    return DL4SS_OK;
}

extern "C" int dl4ss_rowdot_sigmoid_fwd(const float *x, const float *w, const float *b, float *out, int N, int K, void *stream) {
    DL4SS_CHECK_ARG(x && w && out, "rowdot_sigmoid_fwd: null operand");
    DL4SS_CHECK_ARG(N >= 0 && K >= 1, "rowdot_sigmoid_fwd: bad N/K %d/%d", N, K);
    if (N == 0) return DL4SS_OK;
    rowdot_sigmoid_kernel<<<N, 256, 0, (cudaStream_t)stream>>>(x, w, b, out, K);
    DL4SS_LAUNCH_CHECK("rowdot_sigmoid_kernel");
    return DL4SS_OK;
}
