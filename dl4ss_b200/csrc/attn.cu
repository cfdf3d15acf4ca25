// Speaker attention over the T-F embedding, the (round-1) K4 pipeline stage and the mask losses.
//
//  attn_dot_kernel      ATTENTION 'dot' on a materialised embedding: one pass over emb[.,TF,E],
//                       all S speaker queries of the utterance evaluated per tile (the reference
//                       copies the tensor S times and runs a batched GEMV per copy,
//                       TDAA_beta/main_run_sstune_EvalVer.py:453-460,216-226).
//  dl4ss_emb_attn_mask_fwd   Linear+tanh -> attention -> mask without the [B,T,F,E] tensor ever
//                       existing in full: utterance chunks sized to stay L2-resident.
//  mask_loss_kernel     mask x mixture + MSE partial sums (K5).
#include "common.cuh"

namespace dl4ss {

int linear_fwd_impl(const float *A, int lda, const float *W, int ldw, const float *bias, float *C, int ldc,
                    int M, int N, int K, int act, cudaStream_t st);

constexpr int ATT_ROWS = 128;

// cRM decompression exactly as the reference evaluates it in fp32
// (TDAA_beta/main_run_sstune_cRM_EvalVer.py:512): M = -1/C * log((K - m)/(K + m)), m = K*tanh(e).
// +inf for e >~ 9.2 is reference behaviour (SURVEY 7 "cRM numerics").
__device__ __forceinline__ float crm_value(float energy, float crm_k, float crm_c) {
    float m = crm_k * tanhf(energy);
    if (crm_c <= 0.f) return m;
    return (-1.0f / crm_c) * logf((crm_k - m) / (crm_k + m));
}

template <int MODE>
__global__ void __launch_bounds__(ATT_ROWS)
attn_dot_kernel(const float *__restrict__ emb, long long emb_utt_stride, const float *__restrict__ q,
                int S, int TF, int E, float crm_k, float crm_c, float *__restrict__ out) {
    extern __shared__ __align__(16) float sm[];
    const int EQ = (MODE == DL4SS_ATT_DOT_CRM) ? 2 * E : E;
    float *tile = sm;                       // ATT_ROWS * E
    float *qs = sm + ATT_ROWS * E;          // S * EQ
    const int b = blockIdx.y;
    const int r0 = blockIdx.x * ATT_ROWS;
    const int nrows = min(ATT_ROWS, TF - r0);
    const int tid = threadIdx.x;

    const float *src = emb + (size_t)b * emb_utt_stride + (size_t)r0 * E;
    const int nel = nrows * E;
    if (((uintptr_t)src & 7) == 0 && (nel & 1) == 0) {
        const float2 *s2 = reinterpret_cast<const float2 *>(src);
        float2 *t2 = reinterpret_cast<float2 *>(tile);
        for (int i = tid; i < (nel >> 1); i += ATT_ROWS) t2[i] = s2[i];
    } else {
        for (int i = tid; i < nel; i += ATT_ROWS) tile[i] = src[i];
    }
    for (int i = tid; i < S * EQ; i += ATT_ROWS) qs[i] = q[(size_t)b * S * EQ + i];
    __syncthreads();
    if (tid >= nrows) return;
    const float *row = tile + tid * E;
    const size_t tf = (size_t)r0 + tid;
    for (int s = 0; s < S; ++s) {
        const float *qv = qs + s * EQ;
        if (MODE == DL4SS_ATT_DOT) {
            float acc = 0.f;
            for (int e = 0; e < E; ++e) acc = fmaf(row[e], qv[e], acc);
            out[((size_t)b * S + s) * TF + tf] = sigmoid_f(acc);
        } else {
            float a0 = 0.f, a1 = 0.f;
            for (int e = 0; e < E; ++e) {
                a0 = fmaf(row[e], qv[e], a0);
                a1 = fmaf(row[e], qv[E + e], a1);
            }
            float2 o = make_float2(crm_value(a0, crm_k, crm_c), crm_value(a1, crm_k, crm_c));
            reinterpret_cast<float2 *>(out)[((size_t)b * S + s) * TF + tf] = o;
        }
    }
}

static int attn_dot_impl(const float *emb, long long emb_utt_stride, const float *q, int B, int S, int TF,
                         int E, int mode, float crm_k, float crm_c, float *out, cudaStream_t st) {
    const int EQ = (mode == DL4SS_ATT_DOT_CRM) ? 2 * E : E;
    const size_t smem = ((size_t)ATT_ROWS * E + (size_t)S * EQ) * sizeof(float);
    if (smem > 200 * 1024) {
        set_error("attn_dot: E=%d S=%d needs %zu B of shared memory", E, S, smem);
        return DL4SS_EUNSUPPORTED;
    }
    // blockIdx.y carries the utterance (<= 65535 per launch)
    for (int b0 = 0; b0 < B; b0 += 65535) {
        const int nb = (B - b0 < 65535) ? B - b0 : 65535;
        dim3 grid(cdiv(TF, ATT_ROWS), nb);
        const float *e0 = emb + (size_t)b0 * emb_utt_stride;
        const float *q0 = q + (size_t)b0 * S * EQ;
        float *o0 = out + (size_t)b0 * S * TF * (mode == DL4SS_ATT_DOT_CRM ? 2 : 1);
        if (mode == DL4SS_ATT_DOT) {
            DL4SS_CUDA(cudaFuncSetAttribute(attn_dot_kernel<DL4SS_ATT_DOT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            attn_dot_kernel<DL4SS_ATT_DOT><<<grid, ATT_ROWS, smem, st>>>(e0, emb_utt_stride, q0, S, TF, E, crm_k, crm_c, o0);
        } else {
            DL4SS_CUDA(cudaFuncSetAttribute(attn_dot_kernel<DL4SS_ATT_DOT_CRM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            attn_dot_kernel<DL4SS_ATT_DOT_CRM><<<grid, ATT_ROWS, smem, st>>>(e0, emb_utt_stride, q0, S, TF, E, crm_k, crm_c, o0);
        }
        DL4SS_LAUNCH_CHECK("attn_dot_kernel");
    }
    return DL4SS_OK;
}

// --------------------------------------------------------------------------------------- K5
template <int MASK_KIND>
__global__ void __launch_bounds__(256)
mask_loss_kernel(const float *__restrict__ mask, const float *__restrict__ mix, const float *__restrict__ target,
                 int S, long long TF, long long total, double *__restrict__ loss_out) {
    double l0 = 0.0, l1 = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long b = i / TF, tf = i - b * TF;
        if (MASK_KIND == DL4SS_MASK_REAL) {
            const float x = mix[i];
            float sum = 0.f, e0 = 0.f;
            for (int s = 0; s < S; ++s) {
                const long long j = (b * S + s) * TF + tf;
                const float m = mask[j];
                const float d = m * x - target[j];
                e0 = fmaf(d, d, e0);
                sum += m;
            }
            l0 += (double)e0;
            const float d1 = sum - 1.0f;
            l1 += (double)(d1 * d1);
        } else {
            const float2 x = reinterpret_cast<const float2 *>(mix)[i];
            float e0 = 0.f, e1 = 0.f;
            for (int s = 0; s < S; ++s) {
                const long long j = (b * S + s) * TF + tf;
                const float2 m = reinterpret_cast<const float2 *>(mask)[j];
                const float2 y = reinterpret_cast<const float2 *>(target)[j];
                const float pr = m.x * x.x - m.y * x.y;
                const float pi = m.x * x.y + m.y * x.x;
                e0 = fmaf(pr - y.x, pr - y.x, e0);
                e1 = fmaf(pi - y.y, pi - y.y, e1);
            }
            l0 += (double)e0;
            l1 += (double)e1;
        }
    }
    __shared__ double red[2][8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        l0 += __shfl_xor_sync(0xffffffffu, l0, o);
        l1 += __shfl_xor_sync(0xffffffffu, l1, o);
    }
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) { red[0][w] = l0; red[1][w] = l1; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0, c = 0.0;
        for (int i = 0; i < 8; ++i) { a += red[0][i]; c += red[1][i]; }
        atomicAdd(loss_out, a);
        atomicAdd(loss_out + 1, c);
    }
}

}  // namespace dl4ss

using namespace dl4ss;

extern "C" int dl4ss_attn_dot_fwd(const float *emb, long long emb_utt_stride, const float *q, int B, int S,
                                  int TF, int E, int mode, float crm_k, float crm_c, float *mask_out,
                                  void *stream) {
    DL4SS_CHECK_ARG(emb && q && mask_out, "attn_dot_fwd: null operand");
    DL4SS_CHECK_ARG(B >= 0 && S >= 1 && TF >= 1 && E >= 1, "attn_dot_fwd: bad B/S/TF/E %d/%d/%d/%d", B, S, TF, E);
    DL4SS_CHECK_ARG(mode == DL4SS_ATT_DOT || mode == DL4SS_ATT_DOT_CRM, "attn_dot_fwd: bad mode %d", mode);
    if (B == 0) return DL4SS_OK;
    return attn_dot_impl(emb, emb_utt_stride, q, B, S, TF, E, mode, crm_k, crm_c, mask_out, (cudaStream_t)stream);
}

static int k4_chunk_utts(int B, int T, int F, int E) {
    const size_t per_utt = (size_t)T * F * E * sizeof(float);
    size_t u = ((size_t)64 << 20) / per_utt;      // ~64 MB: GEMM output is re-read from L2
    if (u < 1) u = 1;
    if (u > (size_t)B) u = (size_t)B;
    return (int)u;
}

extern "C" size_t dl4ss_emb_attn_mask_workspace_bytes(int B, int T, int F, int E) {
    if (B <= 0 || T <= 0 || F <= 0 || E <= 0) return 0;
    return (size_t)k4_chunk_utts(B, T, F, E) * T * F * E * sizeof(float);
}

extern "C" int dl4ss_emb_attn_mask_fwd(const float *h, const float *W, const float *bias, const float *q,
                                       int B, int T, int F, int E, int K, int S, int mode, float crm_k,
                                       float crm_c, float *mask_out, void *workspace, size_t workspace_bytes,
                                       void *stream) {
    DL4SS_CHECK_ARG(h && W && q && mask_out, "emb_attn_mask_fwd: null operand");
    DL4SS_CHECK_ARG(B >= 0 && T >= 1 && F >= 1 && E >= 1 && K >= 1 && S >= 1, "emb_attn_mask_fwd: bad shape");
    DL4SS_CHECK_ARG(mode == DL4SS_ATT_DOT || mode == DL4SS_ATT_DOT_CRM, "emb_attn_mask_fwd: bad mode %d", mode);
    if (B == 0) return DL4SS_OK;
    const size_t per_utt = (size_t)T * F * E * sizeof(float);
    if (!workspace || workspace_bytes < per_utt) {
        set_error("emb_attn_mask_fwd: workspace %zu B < %zu B (one utterance)", workspace_bytes, per_utt);
        return DL4SS_EWORKSPACE;
    }
    int U = (int)(workspace_bytes / per_utt);
    if (U > B) U = B;
    cudaStream_t st = (cudaStream_t)stream;
    float *buf = (float *)workspace;
    const int EQ = (mode == DL4SS_ATT_DOT_CRM) ? 2 * E : E;
    const int ocomp = (mode == DL4SS_ATT_DOT_CRM) ? 2 : 1;
    for (int u0 = 0; u0 < B; u0 += U) {
        const int nu = (B - u0 < U) ? B - u0 : U;
        int rc = linear_fwd_impl(h + (size_t)u0 * T * K, K, W, K, bias, buf, F * E, nu * T, F * E, K,
                                 DL4SS_ACT_TANH, st);
        if (rc) return rc;
        rc = attn_dot_impl(buf, (long long)T * F * E, q + (size_t)u0 * S * EQ, nu, S, T * F, E, mode, crm_k,
                           crm_c, mask_out + (size_t)u0 * S * T * F * ocomp, st);
        if (rc) return rc;
    }
    return DL4SS_OK;
}

extern "C" int dl4ss_mask_loss_fwd(const float *mask, int mask_kind, const float *mix, const float *target,
                                   int B, int S, int TF, double *loss_out, void *stream) {
    DL4SS_CHECK_ARG(mask && mix && target && loss_out, "mask_loss_fwd: null operand");
    DL4SS_CHECK_ARG(mask_kind == DL4SS_MASK_REAL || mask_kind == DL4SS_MASK_COMPLEX, "mask_loss_fwd: bad mask_kind %d", mask_kind);
    DL4SS_CHECK_ARG(B >= 0 && S >= 1 && TF >= 1, "mask_loss_fwd: bad shape");
    if (B == 0) return DL4SS_OK;
    const long long total = (long long)B * TF;
    long long blocks = cdivll(total, 256 * 4);
    const long long cap = (long long)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    cudaStream_t st = (cudaStream_t)stream;
    if (mask_kind == DL4SS_MASK_REAL)
        mask_loss_kernel<DL4SS_MASK_REAL><<<(unsigned)blocks, 256, 0, st>>>(mask, mix, target, S, TF, total, loss_out);
    else
        mask_loss_kernel<DL4SS_MASK_COMPLEX><<<(unsigned)blocks, 256, 0, st>>>(mask, mix, target, S, TF, total, loss_out);
    DL4SS_LAUNCH_CHECK("mask_loss_kernel");
    return DL4SS_OK;
}

// ------------------------------------------------------------------- permutation-invariant form of K5
// pair[b][s][s'] = sum_tf |mask[b,s]*mix[b] - target[b,s']|^2 : every prediction against every target in one pass
// over the masks, the mixture and the targets (the reference's loss pairs sources by sorted speaker index,
// TDAA_beta/main_run_sstune_EvalVer.py:632-639; the permutation search over these S x S sums is the north-star's PIT
// extension, oracle: oracle/modules_ref.py pit_mse_ref).  One CTA column per utterance, S <= 4.
namespace dl4ss {
constexpr int PIT_SMAX = 4;

template <int MASK_KIND>
__global__ void __launch_bounds__(256)
mask_pair_loss_kernel(const float *__restrict__ mask, const float *__restrict__ mix, const float *__restrict__ target,
                      int S, int TF, double *__restrict__ pair_out) {
    const int b = blockIdx.y;
    float e[PIT_SMAX][PIT_SMAX];
#pragma unroll
    for (int i = 0; i < PIT_SMAX; ++i)
#pragma unroll
        for (int j = 0; j < PIT_SMAX; ++j) e[i][j] = 0.f;
    double acc[PIT_SMAX][PIT_SMAX];
#pragma unroll
    for (int i = 0; i < PIT_SMAX; ++i)
#pragma unroll
        for (int j = 0; j < PIT_SMAX; ++j) acc[i][j] = 0.0;
    int n = 0;
    for (int tf = blockIdx.x * blockDim.x + threadIdx.x; tf < TF; tf += gridDim.x * blockDim.x) {
        if (MASK_KIND == DL4SS_MASK_REAL) {
            const float x = mix[(size_t)b * TF + tf];
            float pr[PIT_SMAX], y[PIT_SMAX];
#pragma unroll
            for (int s = 0; s < PIT_SMAX; ++s) {
                pr[s] = (s < S) ? mask[((size_t)b * S + s) * TF + tf] * x : 0.f;
                y[s] = (s < S) ? target[((size_t)b * S + s) * TF + tf] : 0.f;
            }
#pragma unroll
            for (int i = 0; i < PIT_SMAX; ++i)
#pragma unroll
                for (int j = 0; j < PIT_SMAX; ++j) { const float d = pr[i] - y[j]; e[i][j] = fmaf(d, d, e[i][j]); }
        } else {
            const float2 x = reinterpret_cast<const float2 *>(mix)[(size_t)b * TF + tf];
            float2 pr[PIT_SMAX], y[PIT_SMAX];
#pragma unroll
            for (int s = 0; s < PIT_SMAX; ++s) {
                pr[s] = y[s] = make_float2(0.f, 0.f);
                if (s < S) {
                    const float2 m = reinterpret_cast<const float2 *>(mask)[((size_t)b * S + s) * TF + tf];
                    pr[s] = make_float2(m.x * x.x - m.y * x.y, m.x * x.y + m.y * x.x);
                    y[s] = reinterpret_cast<const float2 *>(target)[((size_t)b * S + s) * TF + tf];
                }
            }
#pragma unroll
            for (int i = 0; i < PIT_SMAX; ++i)
#pragma unroll
                for (int j = 0; j < PIT_SMAX; ++j) {
                    const float dr = pr[i].x - y[j].x, di = pr[i].y - y[j].y;
                    e[i][j] = fmaf(dr, dr, fmaf(di, di, e[i][j]));
                }
        }
        if (++n == 64) {        // fold the fp32 partials into double before they grow
#pragma unroll
            for (int i = 0; i < PIT_SMAX; ++i)
#pragma unroll
                for (int j = 0; j < PIT_SMAX; ++j) { acc[i][j] += (double)e[i][j]; e[i][j] = 0.f; }
            n = 0;
        }
    }
    __shared__ double red[8];
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
#pragma unroll
    for (int i = 0; i < PIT_SMAX; ++i)
#pragma unroll
        for (int j = 0; j < PIT_SMAX; ++j) {
            if (i >= S || j >= S) continue;             // uniform
            double v = acc[i][j] + (double)e[i][j];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            __syncthreads();
            if (l == 0) red[w] = v;
            __syncthreads();
            if (threadIdx.x == 0) {
                double t = 0.0;
                for (int k = 0; k < 8; ++k) t += red[k];
                atomicAdd(pair_out + ((size_t)b * S + i) * S + j, t);
            }
        }
}
}  // namespace dl4ss

extern "C" int dl4ss_mask_pair_loss_fwd(const float *mask, int mask_kind, const float *mix, const float *target,
                                        int B, int S, int TF, double *pair_out, void *stream) {
    DL4SS_CHECK_ARG(mask && mix && target && pair_out, "mask_pair_loss_fwd: null operand");
    DL4SS_CHECK_ARG(mask_kind == DL4SS_MASK_REAL || mask_kind == DL4SS_MASK_COMPLEX, "mask_pair_loss_fwd: bad mask_kind %d", mask_kind);
    DL4SS_CHECK_ARG(B >= 0 && S >= 1 && TF >= 1, "mask_pair_loss_fwd: bad shape");
    if (S > PIT_SMAX) {
        set_error("mask_pair_loss_fwd: S=%d exceeds the %d sources the kernel keeps in registers", S, PIT_SMAX);
        return DL4SS_EUNSUPPORTED;
    }
    if (B == 0) return DL4SS_OK;
    DL4SS_CHECK_ARG(B < 65536, "mask_pair_loss_fwd: B too large for one launch");
    cudaStream_t st = (cudaStream_t)stream;
    DL4SS_CUDA(cudaMemsetAsync(pair_out, 0, (size_t)B * S * S * sizeof(double), st));
    int chunks = cdiv(TF, 256 * 8);
    const int cap = cdiv(sm_count() * 8, B);
    if (chunks > cap) chunks = cap;
    if (chunks < 1) chunks = 1;
    dim3 grid(chunks, B);
    if (mask_kind == DL4SS_MASK_REAL)
        mask_pair_loss_kernel<DL4SS_MASK_REAL><<<grid, 256, 0, st>>>(mask, mix, target, S, TF, pair_out);
    else
        mask_pair_loss_kernel<DL4SS_MASK_COMPLEX><<<grid, 256, 0, st>>>(mask, mix, target, S, TF, pair_out);
    DL4SS_LAUNCH_CHECK("mask_pair_loss_kernel");
    return DL4SS_OK;
}

// ------------------------------------------------------------------- a6 + a7: speaker queries
// e = table[idx[b,s]]  (idx == NULL: table is already e[B,S,EQ]);
// q[b,s,:] = residual*e + Wadj * [mean_t h[b,t,:] ; e]   (Wadj == NULL: q = e)
//   SPEECH_EMBEDDING.forward  TDAA_beta/main_run_sstune_EvalVer.py:357-361
//   ADDJUST.forward + residual TDAA_beta/main_run_sstune_EvalVer.py:371-377,445-446
// One CTA per utterance: the T-mean of the encoder output is a coalesced column sum.
namespace dl4ss {
__global__ void __launch_bounds__(256)
speaker_query_kernel(const float *__restrict__ h, int T, int C, const float *__restrict__ table, int EQ,
                     int num_spk, const long long *__restrict__ idx, int S, const float *__restrict__ Wadj,
                     int residual, float *__restrict__ q, float *__restrict__ hmean_out,
                     int *__restrict__ err) {
    extern __shared__ __align__(16) float sm[];
    float *hm = sm;              // C
    float *es = sm + C;          // S*EQ
    const int b = blockIdx.x, tid = threadIdx.x;
    for (int i = tid; i < S * EQ; i += 256) {
        const int s = i / EQ, e = i - s * EQ;
        if (idx == nullptr) {
            es[i] = table[(size_t)b * S * EQ + i];
        } else {
            long long id = idx[(size_t)b * S + s];
            if (id < 0 || id >= num_spk) { if (e == 0) atomicExch(err, 1); id = 0; }
            es[i] = table[(size_t)id * EQ + e];
        }
    }
    if (Wadj != nullptr) {
        const float *hb = h + (size_t)b * T * C;
        for (int c = tid; c < C; c += 256) {
            float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
            int t = 0;
            for (; t + 3 < T; t += 4) {
                a0 += hb[(size_t)t * C + c];
                a1 += hb[(size_t)(t + 1) * C + c];
                a2 += hb[(size_t)(t + 2) * C + c];
                a3 += hb[(size_t)(t + 3) * C + c];
            }
            for (; t < T; ++t) a0 += hb[(size_t)t * C + c];
            const float m = ((a0 + a1) + (a2 + a3)) / (float)T;
            hm[c] = m;
            if (hmean_out != nullptr) hmean_out[(size_t)b * C + c] = m;
        }
    }
    __syncthreads();
    const int warp = tid >> 5, lane = tid & 31;
    const int ld = C + EQ;
    for (int o = warp; o < S * EQ; o += 8) {
        const int s = o / EQ, j = o - s * EQ;
        float acc = 0.f;
        if (Wadj != nullptr) {
            const float *w = Wadj + (size_t)j * ld;
            for (int c = lane; c < C; c += 32) acc = fmaf(w[c], hm[c], acc);
            for (int e = lane; e < EQ; e += 32) acc = fmaf(w[C + e], es[s * EQ + e], acc);
            acc = warp_sum(acc);
        }
        if (lane == 0) q[((size_t)b * S + s) * EQ + j] = (Wadj == nullptr || residual) ? es[o] + acc : acc;
    }
}
}  // namespace dl4ss

extern "C" int dl4ss_speaker_query_fwd(const float *h, int B, int T, int C, const float *table, int num_spk,
                                       int EQ, const long long *idx, int S, const float *Wadj, int residual,
                                       float *q, float *hmean_out, int *err_flag, void *stream) {
    DL4SS_CHECK_ARG(table && q && (err_flag || !idx), "speaker_query_fwd: null operand");
    DL4SS_CHECK_ARG(!Wadj || h, "speaker_query_fwd: ADDJUST needs the encoder output");
    DL4SS_CHECK_ARG(B >= 0 && T >= 1 && C >= 1 && EQ >= 1 && S >= 1 && num_spk >= 1, "speaker_query_fwd: bad shape");
    if (B == 0) return DL4SS_OK;
    const size_t smem = ((size_t)C + (size_t)S * EQ) * sizeof(float);
    if (smem > 200 * 1024) { set_error("speaker_query_fwd: C=%d S=%d EQ=%d too large", C, S, EQ); return DL4SS_EUNSUPPORTED; }
    DL4SS_CUDA(cudaFuncSetAttribute(speaker_query_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    speaker_query_kernel<<<B, 256, smem, (cudaStream_t)stream>>>(h, T, C, table, EQ, num_spk, idx, S, Wadj, residual,
                                                                 q, hmean_out, err_flag);
    DL4SS_LAUNCH_CHECK("speaker_query_kernel");
    return DL4SS_OK;
}
