// Backward-side kernels of the training step (STFT -> encoder -> masks -> MSE loss -> backward):
//   dl4ss_mask_loss_bwd   d(loss)/d(mask) of the reference objective
//                         (TDAA_beta/main_run_sstune_EvalVer.py:641,659-666 ; cRM ...cRM_EvalVer.py:566-568,741-743)
//   dl4ss_attn_dot_bwd    backward of ATTENTION 'dot' + sigmoid (| cRM) + the tanh of MIX_SPEECH.Linear:
//                         d(mask) -> d(pre-tanh embedding) and d(speaker query) in one pass over emb[.,TF,E]
//   dl4ss_rnn_bwd_step    one time step of BPTT gate arithmetic for a bidirectional LSTM / GRU layer
//                         (nn.LSTM / nn.GRU autograd in the reference, :673 `loss.backward()`).
// The dense contractions of the backward pass (dW, dx, dh, and the per-step dh_rec = dgates * W_hh) are plain
// GEMMs issued by the host code through cuBLAS this round; everything fused or element-wise is here.
#include "common.cuh"
#include <cuda_bf16.h>

namespace dl4ss {

// ----------------------------------------------------------------------------------------- loss backward
template <int MASK_KIND>
__global__ void __launch_bounds__(256)
mask_loss_bwd_kernel(const float *__restrict__ mask, const float *__restrict__ mix, const float *__restrict__ target,
                     int S, long long TF, long long total, float c0, float c1, float *__restrict__ dmask) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long b = i / TF, tf = i - b * TF;
        if (MASK_KIND == DL4SS_MASK_REAL) {
            const float x = mix[i];
            float sum = 0.f;
            for (int s = 0; s < S; ++s) sum += mask[(b * S + s) * TF + tf];
            const float g1 = c1 * (sum - 1.0f);
            for (int s = 0; s < S; ++s) {
                const long long j = (b * S + s) * TF + tf;
                dmask[j] = fmaf(c0 * (mask[j] * x - target[j]), x, g1);
            }
        } else {
            const float2 x = reinterpret_cast<const float2 *>(mix)[i];
            for (int s = 0; s < S; ++s) {
                const long long j = (b * S + s) * TF + tf;
                const float2 m = reinterpret_cast<const float2 *>(mask)[j];
                const float2 y = reinterpret_cast<const float2 *>(target)[j];
                const float dr = c0 * ((m.x * x.x - m.y * x.y) - y.x);
                const float di = c1 * ((m.x * x.y + m.y * x.x) - y.y);
                reinterpret_cast<float2 *>(dmask)[j] = make_float2(dr * x.x + di * x.y, di * x.x - dr * x.y);
            }
        }
    }
}

// ----------------------------------------------------------------------------------------- attention backward
constexpr int AB_ROWS = 128;

// PLANES: dz leaves as bf16 hi/lo planes [2][B*T][ldp] (row (b,t), column f*E+e) -- the operand form of the two GEMMs that
// consume it (dW_lin = dz^T h with MN-major operands, dh = dz W_lin) -- instead of fp32: the same 4 bytes per value, and no
// split passes over the 2 GB tensor afterwards.
template <int MODE, bool PLANES>
__global__ void __launch_bounds__(AB_ROWS)
attn_dot_bwd_kernel(const float *__restrict__ emb, const float *__restrict__ q, const float *__restrict__ mask,
                    const float *__restrict__ dmask, int S, int TF, int E, float crm_k, float crm_c,
                    float *__restrict__ dz, float *__restrict__ dq, __nv_bfloat16 *__restrict__ planes, int F, int ldp,
                    size_t plane_elems, size_t row_base) {
    extern __shared__ __align__(16) float sm[];
    const int NQ = (MODE == DL4SS_ATT_DOT_CRM) ? 2 : 1;        // energies per speaker
    const int EQ = NQ * E;
    const int EP = E | 1;                                      // odd row pitch: conflict-free row-per-thread access
    float *tile = sm;                                          // AB_ROWS * EP   emb rows -> dz rows
    float *qs = tile + AB_ROWS * EP;                           // S * EQ
    float *de = qs + S * EQ;                                   // S * NQ * AB_ROWS
    const int b = blockIdx.y;
    const int r0 = blockIdx.x * AB_ROWS;
    const int nrows = min(AB_ROWS, TF - r0);
    const int tid = threadIdx.x;

    const float *src = emb + ((size_t)b * TF + r0) * E;
    for (int i = tid; i < nrows * E; i += AB_ROWS) tile[(i / E) * EP + (i % E)] = src[i];
    for (int i = tid; i < S * EQ; i += AB_ROWS) qs[i] = q[(size_t)b * S * EQ + i];
    __syncthreads();

    float dev[8];                                              // S * NQ <= 8 energies per row
    if (tid < nrows) {
        const size_t tf = (size_t)r0 + tid;
        for (int s = 0; s < S; ++s) {
            const size_t j = ((size_t)b * S + s) * TF + tf;
            if (MODE == DL4SS_ATT_DOT) {
                const float m = mask[j];
                dev[s] = dmask[j] * m * (1.0f - m);
            } else {
                const float2 m = reinterpret_cast<const float2 *>(mask)[j];
                const float2 g = reinterpret_cast<const float2 *>(dmask)[j];
                if (crm_c > 0.f) {     // M = -1/C log((K-m)/(K+m)), m = K tanh(e)  =>  dM/de = 2/C
                    dev[2 * s] = g.x * (2.0f / crm_c);
                    dev[2 * s + 1] = g.y * (2.0f / crm_c);
                } else {               // compressed mask K tanh(e): d/de = K - m^2/K
                    dev[2 * s] = g.x * (crm_k - m.x * m.x / crm_k);
                    dev[2 * s + 1] = g.y * (crm_k - m.y * m.y / crm_k);
                }
            }
        }
    } else {
        for (int i = 0; i < S * NQ; ++i) dev[i] = 0.f;
    }
    for (int i = 0; i < S * NQ; ++i) de[i * AB_ROWS + tid] = dev[i];
    __syncthreads();
    // dq[s][e] = sum_rows de[s][row] * emb[row][e]   (emb still in the tile)
    for (int i = tid; i < S * EQ; i += AB_ROWS) {
        const int sq = i / E, e = i - sq * E;                  // sq = s*NQ + component
        const float *d = de + sq * AB_ROWS;
        float acc = 0.f;
        for (int r = 0; r < nrows; ++r) acc = fmaf(d[r], tile[r * EP + e], acc);
        // q layout [S][NQ*E]: speaker s, component c -> s*EQ + c*E + e == sq*E + e
        atomicAdd(dq + (size_t)b * S * EQ + (size_t)sq * E + e, acc);
    }
    __syncthreads();
    // dz[row][e] = (sum_s de_s * q_s[e]) * (1 - emb^2), in place in the tile
    if (tid < nrows) {
        float *row = tile + tid * EP;
        for (int e = 0; e < E; ++e) {
            float acc = 0.f;
            for (int i = 0; i < S * NQ; ++i) acc = fmaf(dev[i], qs[i * E + e], acc);
            const float x = row[e];
            row[e] = acc * (1.0f - x * x);
        }
    }
    __syncthreads();
    if constexpr (PLANES) {
        const int T = TF / F;
        if ((E & 1) == 0) {          // two values per thread: 4-byte stores (ldp and f*E + e are even)
            const int E2 = E >> 1;
            for (int i = tid; i < nrows * E2; i += AB_ROWS) {
                const int r = i / E2, e = 2 * (i - r * E2);
                const int tf = r0 + r, t = tf / F, f = tf - t * F;
                const float v0 = tile[r * EP + e], v1 = tile[r * EP + e + 1];
                const __nv_bfloat16 h0 = __float2bfloat16_rn(v0), h1 = __float2bfloat16_rn(v1);
                const size_t o = (row_base + (size_t)b * T + t) * ldp + (size_t)f * E + e;
                *reinterpret_cast<__nv_bfloat162 *>(planes + o) = __halves2bfloat162(h0, h1);
                *reinterpret_cast<__nv_bfloat162 *>(planes + plane_elems + o) =
                    __halves2bfloat162(__float2bfloat16_rn(v0 - __bfloat162float(h0)), __float2bfloat16_rn(v1 - __bfloat162float(h1)));
            }
        } else {
            for (int i = tid; i < nrows * E; i += AB_ROWS) {
                const int r = i / E, e = i - r * E;
                const int tf = r0 + r, t = tf / F, f = tf - t * F;
                const float v = tile[r * EP + e];
                const __nv_bfloat16 hi = __float2bfloat16_rn(v);
                const size_t o = (row_base + (size_t)b * T + t) * ldp + (size_t)f * E + e;
                planes[o] = hi;
                planes[plane_elems + o] = __float2bfloat16_rn(v - __bfloat162float(hi));
            }
        }
    } else {
        float *dst = dz + ((size_t)b * TF + r0) * E;
        for (int i = tid; i < nrows * E; i += AB_ROWS) dst[i] = tile[(i / E) * EP + (i % E)];
    }
}


// Flat form for the planes output with an even E (every reference config: E = 50): a thread owns one (row, column pair);
// 32 rows x E/2 pairs per pass, so consecutive threads read consecutive float2 of emb and write consecutive bf16x2 of both
// planes -- no shared-memory tile, no row-per-thread phases.  The dq partial sums stay in registers over the CTA's
// AF_PASSES passes and leave through one shared-memory reduction and S*NQ*E atomics per 32*AF_PASSES rows (the tiled kernel
// above issues them per 128 rows and measured 2.85 ms at B = 256 -- 17 % of the HBM rate its 3.1 GB need).
__device__ __forceinline__ unsigned bf2_bits(__nv_bfloat162 v) { return *reinterpret_cast<unsigned *>(&v); }
constexpr int AF_THREADS = 768;   // 24 warps x 40 registers: two CTAs per SM
constexpr int AF_RPP = 32;        // rows per pass (fewer when 32 * E/2 > AF_THREADS)
constexpr int AF_PASSES = 16;     // passes per CTA

template <int MODE, int SMAX>          // SMAX >= S speakers: sizes the register arrays
__global__ void __launch_bounds__(AF_THREADS, (SMAX * (MODE == DL4SS_ATT_DOT_CRM ? 2 : 1) <= 2) ? 2 : 1)
attn_dot_bwd_flat_kernel(const float *__restrict__ emb, const float *__restrict__ q, const float *__restrict__ mask,
                         const float *__restrict__ dmask, int S, int TF, int E2, float crm_k, float crm_c,
                         float *__restrict__ dq, __nv_bfloat16 *__restrict__ planes, int F, int ldp,
                         size_t plane_elems, size_t row_base) {
    constexpr int NQ = (MODE == DL4SS_ATT_DOT_CRM) ? 2 : 1;
    constexpr int NEMAX = SMAX * NQ;
    extern __shared__ __align__(16) float2 red[];            // [S*NQ][AF_RPP][E2]
    const int b = blockIdx.y;
    const int rpp = blockDim.x / E2;                          // rows per pass
    const int r0 = blockIdx.x * (rpp * AF_PASSES);
    const int tid = threadIdx.x;
    const int rl = tid / E2, e2 = tid - rl * E2;
    const int NE = S * NQ;                                    // energies per row
    const int E = 2 * E2, T = TF / F;

    float2 qv[NEMAX], acc[NEMAX];
#pragma unroll
    for (int i = 0; i < NEMAX; ++i) {
        acc[i] = make_float2(0.f, 0.f);
        // q layout [S][NQ*E]: energy i = s*NQ + c at i*E
        qv[i] = (i < NE) ? reinterpret_cast<const float2 *>(q + ((size_t)b * NE + i) * E)[e2] : make_float2(0.f, 0.f);
    }
    const float2 *src = reinterpret_cast<const float2 *>(emb) + (size_t)b * TF * E2;
    constexpr int U = 4;                                      // passes whose loads are issued together
    for (int k0 = 0; k0 < AF_PASSES; k0 += U) {
        float2 xs[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int tf = r0 + (k0 + u) * rpp + rl;
            xs[u] = (tf < TF) ? __ldcs(src + (size_t)tf * E2 + e2) : make_float2(0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int tf = r0 + (k0 + u) * rpp + rl;
            if (tf >= TF) continue;
            const float2 x = xs[u];
            float dzx = 0.f, dzy = 0.f;
#pragma unroll
            for (int s_ = 0; s_ < SMAX; ++s_) {
                if (s_ >= S) break;
                const size_t j = ((size_t)b * S + s_) * TF + tf;
                if (MODE == DL4SS_ATT_DOT) {
                    const float m = __ldg(mask + j);
                    const float d = __ldg(dmask + j) * m * (1.0f - m);
                    dzx = fmaf(d, qv[s_].x, dzx); dzy = fmaf(d, qv[s_].y, dzy);
                    acc[s_].x = fmaf(d, x.x, acc[s_].x); acc[s_].y = fmaf(d, x.y, acc[s_].y);
                } else {
                    const float2 m = __ldg(reinterpret_cast<const float2 *>(mask) + j);
                    const float2 g = __ldg(reinterpret_cast<const float2 *>(dmask) + j);
                    float d0, d1;
                    if (crm_c > 0.f) { d0 = g.x * (2.0f / crm_c); d1 = g.y * (2.0f / crm_c); }
                    else { d0 = g.x * (crm_k - m.x * m.x / crm_k); d1 = g.y * (crm_k - m.y * m.y / crm_k); }
                    constexpr int LAST = NEMAX - 1;
                    const int i0 = (2 * s_) & LAST, i1 = (2 * s_ + 1) & LAST;     // (& LAST: keeps the dead MODE's indices in range)
                    dzx = fmaf(d0, qv[i0].x, dzx); dzy = fmaf(d0, qv[i0].y, dzy);
                    dzx = fmaf(d1, qv[i1].x, dzx); dzy = fmaf(d1, qv[i1].y, dzy);
                    acc[i0].x = fmaf(d0, x.x, acc[i0].x); acc[i0].y = fmaf(d0, x.y, acc[i0].y);
                    acc[i1].x = fmaf(d1, x.x, acc[i1].x); acc[i1].y = fmaf(d1, x.y, acc[i1].y);
                }
            }
            const float v0 = dzx * (1.0f - x.x * x.x), v1 = dzy * (1.0f - x.y * x.y);
            const int t = tf / F, f = tf - t * F;
            const size_t o = (row_base + (size_t)b * T + t) * ldp + (size_t)f * E + 2 * e2;
            const __nv_bfloat16 h0 = __float2bfloat16_rn(v0), h1 = __float2bfloat16_rn(v1);
            __stcs(reinterpret_cast<unsigned *>(planes + o), bf2_bits(__halves2bfloat162(h0, h1)));
            __stcs(reinterpret_cast<unsigned *>(planes + plane_elems + o),
                   bf2_bits(__halves2bfloat162(__float2bfloat16_rn(v0 - __bfloat162float(h0)), __float2bfloat16_rn(v1 - __bfloat162float(h1)))));
        }
    }
    // dq[i][e] = sum over the CTA's rows: registers -> [i][row lane][pair] -> one thread per (i, pair)
#pragma unroll
    for (int i = 0; i < NEMAX; ++i)
        if (i < NE) red[(i * rpp + rl) * E2 + e2] = acc[i];
    __syncthreads();
    if (tid < NE * E2) {
        const int i = tid / E2, e = tid - i * E2;
        float2 a = make_float2(0.f, 0.f);
#pragma unroll 8
        for (int r = 0; r < rpp; ++r) {
            const float2 v = red[(i * rpp + r) * E2 + e];
            a.x += v.x; a.y += v.y;
        }
        float *d = dq + ((size_t)b * NE + i) * E + 2 * e;
        atomicAdd(d, a.x);
        atomicAdd(d + 1, a.y);
    }
}

// ----------------------------------------------------------------------------------------- BPTT step
template <int CELL>
__global__ void __launch_bounds__(256)
rnn_bwd_step_kernel(int s, const float *__restrict__ dy, const float *__restrict__ dh_rec,
                    const float *__restrict__ gates, const float *__restrict__ cells, const float *__restrict__ y,
                    float *__restrict__ carry, float *__restrict__ dgx, float *__restrict__ dgh,
                    float *__restrict__ dg_cur, int B, int T, int H) {
    constexpr int G = (CELL == DL4SS_CELL_LSTM) ? 4 : 3;
    const long long total = 2ll * B * H;
    const size_t GH = (size_t)G * H;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int u = (int)(i % H);
        const int b = (int)((i / H) % B);
        const int dir = (int)(i / ((long long)H * B));
        const int t = dir ? s : (T - 1 - s);                   // backward walks the forward order in reverse
        const int tp = dir ? t + 1 : t - 1;                    // the step the forward pass came from
        const bool has_prev = (tp >= 0 && tp < T);
        const size_t st = (size_t)dir * B * H + (size_t)b * H + u;        // [2,B,H] state index
        const size_t row = ((size_t)b * T + t) * 2 + dir;
        float dh = dy[((size_t)b * T + t) * 2 * H + (size_t)dir * H + u];
        if (s > 0) dh += dh_rec[st];
        const float *gr = gates + row * GH + u;
        float *ox = dgx + row * GH + u;
        float *oc = dg_cur + ((size_t)dir * B + b) * GH + u;
        if constexpr (CELL == DL4SS_CELL_LSTM) {
            const float ig = gr[0], fg = gr[H], gg = gr[2 * (size_t)H], og = gr[3 * (size_t)H];
            const float c = cells[row * H + u];
            const float cp = has_prev ? cells[(((size_t)b * T + tp) * 2 + dir) * H + u] : 0.f;
            const float tc = tanhf(c);
            float dc = dh * og * (1.0f - tc * tc);
            if (s > 0) dc += carry[st];
            const float dai = dc * gg * ig * (1.0f - ig);
            const float daf = dc * cp * fg * (1.0f - fg);
            const float dag = dc * ig * (1.0f - gg * gg);
            const float dao = dh * tc * og * (1.0f - og);
            carry[st] = dc * fg;
            ox[0] = dai; ox[H] = daf; ox[2 * (size_t)H] = dag; ox[3 * (size_t)H] = dao;
            oc[0] = dai; oc[H] = daf; oc[2 * (size_t)H] = dag; oc[3 * (size_t)H] = dao;
        } else {
            if (s > 0) dh += carry[st];                        // the z * h_{t-1} path
            const float rg = gr[0], zg = gr[H], ng = gr[2 * (size_t)H];
            const float hn = cells[row * H + u];               // W_hn h + b_hn saved by the forward kernel
            const float hp = has_prev ? y[((size_t)b * T + tp) * 2 * H + (size_t)dir * H + u] : 0.f;
            const float dn = dh * (1.0f - zg);
            const float dan = dn * (1.0f - ng * ng);
            const float dar = dan * hn * rg * (1.0f - rg);
            const float daz = dh * (hp - ng) * zg * (1.0f - zg);
            const float dhn = dan * rg;
            carry[st] = dh * zg;
            float *oh = dgh + row * GH + u;
            ox[0] = dar; ox[H] = daz; ox[2 * (size_t)H] = dan;
            oh[0] = dar; oh[H] = daz; oh[2 * (size_t)H] = dhn;
            oc[0] = dar; oc[H] = daz; oc[2 * (size_t)H] = dhn;
        }
    }
}

}  // namespace dl4ss

using namespace dl4ss;

extern "C" int dl4ss_mask_loss_bwd(const float *mask, int mask_kind, const float *mix, const float *target, int B,
                                   int S, int TF, float c0, float c1, float *dmask, void *stream) {
    DL4SS_CHECK_ARG(mask && mix && target && dmask, "mask_loss_bwd: null operand");
    DL4SS_CHECK_ARG(mask_kind == DL4SS_MASK_REAL || mask_kind == DL4SS_MASK_COMPLEX, "mask_loss_bwd: bad mask_kind %d", mask_kind);
    DL4SS_CHECK_ARG(B >= 0 && S >= 1 && TF >= 1, "mask_loss_bwd: bad B/S/TF");
    if (B == 0) return DL4SS_OK;
    const long long total = (long long)B * TF;
    long long blocks = cdivll(total, 256);
    if (blocks > (long long)sm_count() * 16) blocks = (long long)sm_count() * 16;
    if (mask_kind == DL4SS_MASK_REAL)
        mask_loss_bwd_kernel<DL4SS_MASK_REAL><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(mask, mix, target, S, TF, total, c0, c1, dmask);
    else
        mask_loss_bwd_kernel<DL4SS_MASK_COMPLEX><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(mask, mix, target, S, TF, total, c0, c1, dmask);
    DL4SS_LAUNCH_CHECK("mask_loss_bwd_kernel");
    return DL4SS_OK;
}

static int attn_dot_bwd_impl(const float *emb, const float *q, const float *mask, const float *dmask, int B,
                             int S, int TF, int E, int mode, float crm_k, float crm_c, float *dz, float *dq,
                             __nv_bfloat16 *planes, int F, int ldp, void *stream) {
    DL4SS_CHECK_ARG(emb && q && mask && dmask && (dz || planes) && dq, "attn_dot_bwd: null operand");
    DL4SS_CHECK_ARG(mode == DL4SS_ATT_DOT || mode == DL4SS_ATT_DOT_CRM, "attn_dot_bwd: bad mode %d", mode);
    const int NQ = (mode == DL4SS_ATT_DOT_CRM) ? 2 : 1;
    DL4SS_CHECK_ARG(B >= 0 && S >= 1 && S * NQ <= 8 && TF >= 1 && E >= 1, "attn_dot_bwd: bad shape (S*%d <= 8)", NQ);
    if (B == 0) return DL4SS_OK;
    cudaStream_t st = (cudaStream_t)stream;
    DL4SS_CUDA(cudaMemsetAsync(dq, 0, (size_t)B * S * NQ * E * sizeof(float), st));
    const size_t smem = ((size_t)AB_ROWS * (E | 1) + (size_t)S * NQ * E + (size_t)S * NQ * AB_ROWS) * sizeof(float);
    if (smem > 200 * 1024) {
        set_error("attn_dot_bwd: E=%d S=%d needs %zu B of shared memory", E, S, smem);
        return DL4SS_EUNSUPPORTED;
    }
    if (planes && (E & 1) == 0 && E <= 64) {
        const int E2 = E / 2, NE = S * NQ;
        const int rpp = (AF_RPP * E2 <= AF_THREADS) ? AF_RPP : AF_THREADS / E2;
        const size_t smem2 = (size_t)NE * rpp * E2 * sizeof(float2);
        for (int b0 = 0; b0 < B; b0 += 65535) {
            const int nb = (B - b0 < 65535) ? B - b0 : 65535;
            dim3 grid(cdiv(TF, rpp * AF_PASSES), nb);
            const size_t mo = (size_t)b0 * S * TF * NQ;
            const size_t plane_elems = (size_t)B * (TF / F) * ldp;
            const size_t row_base = (size_t)b0 * (TF / F);
#define LAUNCH_AF(MODE, SM)                                                                                      \
            do {                                                                                                 \
                DL4SS_CUDA(cudaFuncSetAttribute(attn_dot_bwd_flat_kernel<MODE, SM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2)); \
                attn_dot_bwd_flat_kernel<MODE, SM><<<grid, rpp * E2, smem2, st>>>(emb + (size_t)b0 * TF * E,  \
                    q + (size_t)b0 * NE * E, mask + mo, dmask + mo, S, TF, E2, crm_k, crm_c, dq + (size_t)b0 * NE * E, \
                    planes, F, ldp, plane_elems, row_base);                                                      \
            } while (0)
            if (mode == DL4SS_ATT_DOT) { if (S <= 2) LAUNCH_AF(DL4SS_ATT_DOT, 2); else if (S <= 4) LAUNCH_AF(DL4SS_ATT_DOT, 4); else LAUNCH_AF(DL4SS_ATT_DOT, 8); }
            else { if (S <= 2) LAUNCH_AF(DL4SS_ATT_DOT_CRM, 2); else LAUNCH_AF(DL4SS_ATT_DOT_CRM, 4); }
#undef LAUNCH_AF
            DL4SS_LAUNCH_CHECK("attn_dot_bwd_flat_kernel");
        }
        return DL4SS_OK;
    }
    for (int b0 = 0; b0 < B; b0 += 65535) {
        const int nb = (B - b0 < 65535) ? B - b0 : 65535;
        dim3 grid(cdiv(TF, AB_ROWS), nb);
        const size_t mo = (size_t)b0 * S * TF * NQ;
        const size_t plane_elems = planes ? (size_t)B * (TF / F) * ldp : 0;
        const size_t row_base = planes ? (size_t)b0 * (TF / F) : 0;
#define LAUNCH_AB(MODE, PL)                                                                                      \
        do {                                                                                                     \
            DL4SS_CUDA(cudaFuncSetAttribute(attn_dot_bwd_kernel<MODE, PL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
            attn_dot_bwd_kernel<MODE, PL><<<grid, AB_ROWS, smem, st>>>(emb + (size_t)b0 * TF * E, q + (size_t)b0 * S * NQ * E, \
                mask + mo, dmask + mo, S, TF, E, crm_k, crm_c, dz ? dz + (size_t)b0 * TF * E : nullptr,            \
                dq + (size_t)b0 * S * NQ * E, planes, F, ldp, plane_elems, row_base);                             \
        } while (0)
        if (planes) { if (mode == DL4SS_ATT_DOT) LAUNCH_AB(DL4SS_ATT_DOT, true); else LAUNCH_AB(DL4SS_ATT_DOT_CRM, true); }
        else { if (mode == DL4SS_ATT_DOT) LAUNCH_AB(DL4SS_ATT_DOT, false); else LAUNCH_AB(DL4SS_ATT_DOT_CRM, false); }
#undef LAUNCH_AB
        DL4SS_LAUNCH_CHECK("attn_dot_bwd_kernel");
    }
    return DL4SS_OK;
}

extern "C" int dl4ss_attn_dot_bwd(const float *emb, const float *q, const float *mask, const float *dmask, int B,
                                  int S, int TF, int E, int mode, float crm_k, float crm_c, float *dz, float *dq,
                                  void *stream) {
    DL4SS_CHECK_ARG(dz, "attn_dot_bwd: null operand");
    return attn_dot_bwd_impl(emb, q, mask, dmask, B, S, TF, E, mode, crm_k, crm_c, dz, dq, nullptr, 1, 0, stream);
}

extern "C" int dl4ss_attn_dot_bwd_planes(const float *emb, const float *q, const float *mask, const float *dmask, int B,
                                         int S, int T, int F, int E, int mode, float crm_k, float crm_c, void *dz_planes,
                                         int ldp, float *dq, void *stream) {
    DL4SS_CHECK_ARG(dz_planes && T >= 1 && F >= 1 && ldp >= F * E && ldp % 8 == 0 && (((uintptr_t)dz_planes) & 15) == 0,
                    "attn_dot_bwd_planes: dz_planes null / misaligned or bad T/F/ldp %d/%d/%d (ldp >= F*E, a multiple of 8)", T, F, ldp);
    return attn_dot_bwd_impl(emb, q, mask, dmask, B, S, T * F, E, mode, crm_k, crm_c, nullptr, dq,
                             (__nv_bfloat16 *)dz_planes, F, ldp, stream);
}

extern "C" int dl4ss_rnn_bwd_step(int cell, int s, const float *dy, const float *dh_rec, const float *gates_save,
                                  const float *cell_save, const float *y, float *carry, float *dgx, float *dgh,
                                  float *dg_cur, int B, int T, int H, void *stream) {
    DL4SS_CHECK_ARG(cell == DL4SS_CELL_LSTM || cell == DL4SS_CELL_GRU, "rnn_bwd_step: bad cell %d", cell);
    DL4SS_CHECK_ARG(dy && gates_save && cell_save && carry && dgx && dg_cur, "rnn_bwd_step: null operand");
    DL4SS_CHECK_ARG(cell == DL4SS_CELL_LSTM || (y && dgh), "rnn_bwd_step: GRU needs y and dgh");
    DL4SS_CHECK_ARG(s == 0 || dh_rec, "rnn_bwd_step: dh_rec is null");
    DL4SS_CHECK_ARG(B >= 0 && T >= 1 && H >= 1 && s >= 0 && s < T, "rnn_bwd_step: bad B/T/H/s");
    if (B == 0) return DL4SS_OK;
    const long long total = 2ll * B * H;
    long long blocks = cdivll(total, 256);
    if (blocks > (long long)sm_count() * 16) blocks = (long long)sm_count() * 16;
    if (cell == DL4SS_CELL_LSTM)
        rnn_bwd_step_kernel<DL4SS_CELL_LSTM><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
            s, dy, dh_rec, gates_save, cell_save, y, carry, dgx, dgh, dg_cur, B, T, H);
    else
        rnn_bwd_step_kernel<DL4SS_CELL_GRU><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
            s, dy, dh_rec, gates_save, cell_save, y, carry, dgx, dgh, dg_cur, B, T, H);
    DL4SS_LAUNCH_CHECK("rnn_bwd_step_kernel");
    return DL4SS_OK;
}

// ----------------------------------------------------------------------------------------- a1: mixture synthesis
// The per-source waveform preprocessing of the reference generators, batched on the GPU
// (TDAA_beta/predata_fromList.py:140-177 ; Torch_multi/predata_multiAims.py:144-177):
//   x = x[:n] ; x -= mean(x) ; x /= max|x| ; zero-pad to L ; x *= 10^(dB/20) ; mix = sum_s x_s
// One CTA per utterance; three passes over each source (sum, max|x-mean|, write).
namespace dl4ss {
__device__ __forceinline__ float block_reduce(float v, float *red, bool is_max) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float w = __shfl_xor_sync(0xffffffffu, v, o);
        v = is_max ? fmaxf(v, w) : v + w;
    }
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    float r = is_max ? 0.f : 0.f;
    for (int i = 0; i < nw; ++i) r = is_max ? fmaxf(r, red[i]) : r + red[i];
    return r;
}

__global__ void __launch_bounds__(512)
premix_kernel(const float *__restrict__ src, const int *__restrict__ lengths, const int *__restrict__ shifts,
              const float *__restrict__ gains_db, int S, int L, float *__restrict__ src_out, float *__restrict__ mix_out) {
    __shared__ float red[32];
    __shared__ float sc_mean[16], sc_scale[16];
    const int b = blockIdx.x, tid = threadIdx.x;
    for (int s = 0; s < S; ++s) {
        const float *x = src + ((size_t)b * S + s) * L;
        int n = lengths ? lengths[b * S + s] : L;
        n = n < 0 ? 0 : (n > L ? L : n);
        double acc = 0.0;                                   // the reference works in float64
        for (int i = tid; i < n; i += blockDim.x) acc += (double)x[i];
        // reduce the double sum through two floats (hi + lo) to reuse the float reducer
        const float hi = (float)acc, lo = (float)(acc - (double)hi);
        const float shi = block_reduce(hi, red, false), slo = block_reduce(lo, red, false);
        const float mean = n > 0 ? (float)(((double)shi + (double)slo) / (double)n) : 0.f;
        float mx = 0.f;
        for (int i = tid; i < n; i += blockDim.x) mx = fmaxf(mx, fabsf(x[i] - mean));
        mx = block_reduce(mx, red, true);
        if (tid == 0) {
            sc_mean[s] = mean;
            sc_scale[s] = (mx > 0.f ? 1.0f / mx : 0.f) * exp10f(gains_db[b * S + s] * 0.05f);
        }
    }
    __syncthreads();
    for (int i = tid; i < L; i += blockDim.x) {
        float m = 0.f;
        for (int s = 0; s < S; ++s) {
            int n = lengths ? lengths[b * S + s] : L;
            n = n < 0 ? 0 : (n > L ? L : n);
            const size_t o = ((size_t)b * S + s) * L + i;
            int j = i;
            if (shifts && i < n) {                          // AUGMENT_DATA: signal = append(signal[shift:], signal[:shift])
                j = i + shifts[b * S + s] % n;
                j = j < 0 ? j + n : (j >= n ? j - n : j);
            }
            const float v = (i < n) ? (src[o - i + j] - sc_mean[s]) * sc_scale[s] : 0.f;
            if (src_out) src_out[o] = v;
            m += v;
        }
        mix_out[(size_t)b * L + i] = m;
    }
}
}  // namespace dl4ss

extern "C" int dl4ss_premix_shift_fwd(const float *src, const int *lengths, const int *shifts, const float *gains_db,
                                      int B, int S, int L, float *src_out, float *mix_out, void *stream) {
    if (B == 0) return DL4SS_OK;
    DL4SS_CHECK_ARG(src && gains_db && mix_out, "premix_fwd: null operand");
    DL4SS_CHECK_ARG(B >= 0 && S >= 1 && S <= 16 && L >= 1, "premix_fwd: bad B/S/L %d/%d/%d (S <= 16)", B, S, L);
    DL4SS_CHECK_ARG(!(shifts && src_out == src), "premix_fwd: src_out may not alias src when shifts are given");
    premix_kernel<<<B, 512, 0, (cudaStream_t)stream>>>(src, lengths, shifts, gains_db, S, L, src_out, mix_out);
    DL4SS_LAUNCH_CHECK("premix_kernel");
    return DL4SS_OK;
}

extern "C" int dl4ss_premix_fwd(const float *src, const int *lengths, const float *gains_db, int B, int S, int L,
                                float *src_out, float *mix_out, void *stream) {
    return dl4ss_premix_shift_fwd(src, lengths, nullptr, gains_db, B, S, L, src_out, mix_out, stream);
}
