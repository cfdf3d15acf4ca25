// Tensor-core projections for sm_100a: tcgen05.mma (kind::f16, bf16 operands, fp32 accumulation in
// TMEM), operands fed by TMA into 128B-swizzled shared memory, warp-specialised persistent CTAs.
//
// fp32 parity on bf16 tensor cores ("bf16x3"): every fp32 operand is pre-split into two bf16 planes
// x = hi + lo (hi = bf16(x), lo = bf16(x - hi), 16 significant bits together) and each k-block
// issues three MMAs  A_hi*W_hi + A_hi*W_lo + A_lo*W_hi  into the same fp32 accumulator.  Measured
// against fp64 on the full model the masks move by ~1e-6 (single-pass bf16: 9e-4, tf32: 1.2e-4 --
// both outside the 1e-4 parity bar).
//
//   dl4ss_split_bf16              fp32 [R,K] -> bf16 planes [2][R][Kp], Kp = K rounded up to 64
//   dl4ss_linear_tc_fwd           C = A*W^T + bias                       (RNN input projections)
//   dl4ss_emb_attn_mask_tc_fwd    K4: (h*W^T + b) -> tanh -> <.,q_s> over E -> sigmoid | cRM,
//                                 evaluated in the epilogue straight out of TMEM: the [B,T,F,E]
//                                 embedding (8 MB/utterance) never exists anywhere.
//
// Tile 128 x 256 x 64, 2 smem stages x (A_hi,A_lo,B_hi,B_lo) = 192 KB, 2 TMEM accumulator stages x
// 256 columns; warp 0 = TMA producer, warp 1 = MMA issuer (+TMEM alloc), warps 2-5 = epilogue.
#include "tc_ptx.cuh"
#include <stdlib.h>
#include <mutex>

namespace dl4ss {

constexpr int TBM = 128, TBN = 256, TBK = 64, TSTAGES = 2;
constexpr int TA_BYTES = TBM * TBK * 2;                       // 16 KB
constexpr int TB_BYTES = TBN * TBK * 2;                       // 32 KB
constexpr int TSTAGE_BYTES = 2 * TA_BYTES + 2 * TB_BYTES;     // 96 KB
constexpr int TC_EPI_WARPS = 20;                              // up to 5 per TMEM sub-partition (EpiTraits<>::PARTS of them work)
constexpr int TC_THREADS = 64 + 32 * TC_EPI_WARPS;
constexpr int TC_STAGE_PITCH = 33;                            // floats per row of a warp's 32x32 store-transpose tile
constexpr int TC_RING = 4;                                    // work-item ring slots (power of two)
constexpr int TC_STORE_WARPS = 8;                             // warps of the plain epilogue (4 sub-partitions x 2 parts)
constexpr size_t TC_SMEM = (size_t)TSTAGES * TSTAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/ +
                           (size_t)TC_STORE_WARPS * 32 * TC_STAGE_PITCH * sizeof(float);

__device__ __forceinline__ bool elect_one_lane() {       // one lane of the converged warp (elect.sync: ptxas then emits straight-line uniform-datapath code)
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(pred));
    return pred != 0;
}

// ------------------------------------------------------------------------------------------ split
__global__ void __launch_bounds__(256)
split_bf16_kernel(const float *__restrict__ x, int ld, long long R, int K, int Kp, __nv_bfloat16 *__restrict__ planes) {
    const int chunks = Kp >> 3;
    const long long total = R * chunks;
    __nv_bfloat16 *hi = planes, *lo = planes + (size_t)R * Kp;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / chunks;
        const int k0 = (int)(i - r * chunks) * 8;
        const float *src = x + (size_t)r * ld + k0;
        float v[8];
        if (k0 + 8 <= K && (((uintptr_t)src) & 15) == 0) {
            const float4 a = *reinterpret_cast<const float4 *>(src);
            const float4 b = *reinterpret_cast<const float4 *>(src + 4);
            v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = (k0 + j < K) ? src[j] : 0.f;
        }
        __align__(16) __nv_bfloat16 h[8], l[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            h[j] = __float2bfloat16_rn(v[j]);
            l[j] = __float2bfloat16_rn(v[j] - __bfloat162float(h[j]));
        }
        *reinterpret_cast<uint4 *>(hi + (size_t)r * Kp + k0) = *reinterpret_cast<const uint4 *>(h);
        *reinterpret_cast<uint4 *>(lo + (size_t)r * Kp + k0) = *reinterpret_cast<const uint4 *>(l);
    }
}

// transposing split: fp32 x[R, C] -> bf16 planes [2][C][Rp] (Rp = R rounded up to 64, zero padded): the K-major
// operand form of x^T, for the contractions over the row dimension of the backward pass (dW = dY^T * X)
__global__ void __launch_bounds__(256)
split_bf16_t_kernel(const float *__restrict__ x, long long ld, int R, int C, int Rp, __nv_bfloat16 *__restrict__ planes) {
    __shared__ float tile[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;       // 32 x 8
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = r0 + ty + 8 * i, c = c0 + tx;
        tile[ty + 8 * i][tx] = (r < R && c < C) ? x[(size_t)r * ld + c] : 0.f;
    }
    __syncthreads();
    __nv_bfloat16 *hi = planes, *lo = planes + (size_t)C * Rp;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int c = c0 + ty + 8 * i, r = r0 + tx;
        if (c < C && r < Rp) {
            const float v = tile[tx][ty + 8 * i];
            const __nv_bfloat16 h = __float2bfloat16_rn(v);
            hi[(size_t)c * Rp + r] = h;
            lo[(size_t)c * Rp + r] = __float2bfloat16_rn(v - __bfloat162float(h));
        }
    }
}

// ------------------------------------------------------------------------------------------ epilogues
// ACT: DL4SS_ACT_NONE | DL4SS_ACT_TANH | DL4SS_ACT_SIGMOID; ATOMIC: split-K launch, partial tiles are summed into a zeroed C
// with red.add (ACT must be NONE).  Both compile time: the epilogue is on the critical path.
template <int ACT, bool ATOMIC = false>
struct EpiPlain {
    float *C;
    const float *bias;
    int ldc;
    int tr = 0;          // ATOMIC only: the tile is the transpose of what C holds -- element (m, n) is added to C[n * ldc + m]
};
struct EpiAttn {
    const float *bias;   // [F*E]
    const float *q;      // [B,S,EQ]
    float *out;          // [B,S,T,F] or [B,S,T,F,2]
    int T, F, S, crm;    // crm: 0 sigmoid masks, 1 cRM pairs
    float crm_k, crm_c;
};

constexpr int ATT_E = 50;                       // embedding width the fused epilogue is built for
constexpr int ATT_BINS = TBN / ATT_E;           // 5 frequency bins per N tile, densely packed: columns [50j, 50j+50)
constexpr int ATT_SMAX = 4;                     // speakers per utterance handled in registers

template <typename Epi> struct EpiTraits;
// PARTS: epilogue warps per TMEM sub-partition (each takes 1/PARTS of the tile's columns).  The plain store
// epilogue is bound by its row-strided stores and is fastest with 2, the attention epilogue is math bound: 4.
template <int ACT, bool ATOMIC> struct EpiTraits<EpiPlain<ACT, ATOMIC>> { static constexpr int NSTEP = TBN; static constexpr int PARTS = 2; };
template <> struct EpiTraits<EpiAttn> { static constexpr int NSTEP = ATT_BINS * ATT_E; static constexpr int PARTS = ATT_BINS; };

__device__ __forceinline__ float crm_value_tc(float energy, float crm_k, float crm_c) {
    float m = crm_k * tanhf(energy);
    if (crm_c <= 0.f) return m;
    return (-1.0f / crm_c) * logf((crm_k - m) / (crm_k + m));
}

// one accumulator row (this thread's TMEM lane) -> global
template <int ACT>
__device__ __forceinline__ float apply_act(float x) {
    return ACT == DL4SS_ACT_TANH ? tanh_f(x) : ACT == DL4SS_ACT_SIGMOID ? sigmoid_f(x) : x;
}

// Plain store epilogue of one warp: its 32 accumulator rows x (TBN / PARTS) columns, 32 columns at a time.
// TMEM hands every lane ONE ROW; written out like that a store instruction touches 32 different lines with
// 16 bytes each.  The 32x32 block is transposed through a per-warp smem tile instead, so that 8 lanes write
// one row's 128 contiguous bytes (4 full lines per instruction); bias and activation ride the transposed side.
template <int ACT, bool ATOMIC>
__device__ __forceinline__ void epilogue_tile(const EpiPlain<ACT, ATOMIC> &e, uint32_t taddr, int m_base, int n0, int M, int N,
                                              int part, int lane, float *stage) {
    constexpr int CH = TBN / 32 / EpiTraits<EpiPlain<ACT, ATOMIC>>::PARTS;
    const bool vec = ((e.ldc & 3) == 0) && ((((uintptr_t)e.C) & 15) == 0);
    const int rsub = lane >> 3, col = (lane & 7) * 4;
#pragma unroll 1
    for (int cc = part * CH; cc < (part + 1) * CH; ++cc) {
        float v[32];
        tmem_ld16(taddr + cc * 32, *reinterpret_cast<float(*)[16]>(&v[0]));       // warp-collective
        tmem_ld16(taddr + cc * 32 + 16, *reinterpret_cast<float(*)[16]>(&v[16]));
        const int n = n0 + cc * 32;
        if (n >= N) continue;                                                      // warp-uniform
        if (ATOMIC && e.tr) {
            // transposed output: TMEM's one-row-per-lane form is already the coalesced one (32 lanes = 32 consecutive m)
            const int m = m_base + lane;
            if (m < M) {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    if (n + j < N) atomicAdd(e.C + (size_t)(n + j) * e.ldc + m, v[j]);
            }
            continue;
        }
        __syncwarp();                                                              // the previous chunk's readers are done
#pragma unroll
        for (int j = 0; j < 32; ++j) stage[lane * TC_STAGE_PITCH + j] = v[j];
        __syncwarp();
        float bz[4] = {0.f, 0.f, 0.f, 0.f};
        if (e.bias) {
#pragma unroll
            for (int k = 0; k < 4; ++k) if (n + col + k < N) bz[k] = __ldg(e.bias + n + col + k);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int row = 4 * i + rsub;
            const int m = m_base + row;
            if (m >= M) continue;
            float o[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) o[k] = apply_act<ACT>(stage[row * TC_STAGE_PITCH + col + k] + bz[k]);
            float *dst = e.C + (size_t)m * e.ldc + n + col;
            if (ATOMIC) {
#pragma unroll
                for (int k = 0; k < 4; ++k) if (n + col + k < N) atomicAdd(dst + k, o[k]);
            } else if (vec && n + col + 3 < N) {
                *reinterpret_cast<float4 *>(dst) = make_float4(o[0], o[1], o[2], o[3]);
            } else {
#pragma unroll
                for (int k = 0; k < 4; ++k) if (n + col + k < N) dst[k] = o[k];
            }
        }
    }
}

// Fused attention epilogue of one warp: frequency bin `part` of the tile (accumulator columns [50*part, 50*part+50)),
// one accumulator row (b,t) per lane.  n0 = first W row of the tile (a multiple of 50).
__device__ __forceinline__ void epilogue_tile(const EpiAttn &e, uint32_t taddr, int m_base, int n0, int M, int N,
                                              int part, int lane, float *stage) {
    (void)N; (void)stage;
    const int EQ = e.crm ? 2 * ATT_E : ATT_E;
    const int m = m_base + lane;
    const bool valid = m < M;
    const int mm = valid ? m : 0;
    const int b = mm / e.T, t = mm - b * e.T;
    const float *qb = e.q + (size_t)b * e.S * EQ;
    const int f = n0 / ATT_E + part;
    const bool fvalid = f < e.F;                      // warp-uniform
    const float *bias = e.bias + (size_t)(fvalid ? f : 0) * ATT_E;
    float en[2 * ATT_SMAX];
#pragma unroll
    for (int s = 0; s < 2 * ATT_SMAX; ++s) en[s] = 0.f;
#pragma unroll
    for (int c = 0; c < 4; ++c) {                     // 16 + 16 + 16 + 2 columns
        float v[16];
        if (c < 3) {
            tmem_ld16(taddr + part * ATT_E + c * 16, v);
        } else {
            float v2[2];
            tmem_ld2(taddr + part * ATT_E + 48, v2);
            v[0] = v2[0]; v[1] = v2[1];
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const int ee = c * 16 + j;                // compile time
            if (ee >= ATT_E) continue;
            const float x = tanh_f(v[j] + __ldg(bias + ee));
#pragma unroll
            for (int s = 0; s < ATT_SMAX; ++s) {
                if (s < e.S) {
                    en[2 * s] = fmaf(x, __ldg(qb + s * EQ + ee), en[2 * s]);
                    if (e.crm) en[2 * s + 1] = fmaf(x, __ldg(qb + s * EQ + ATT_E + ee), en[2 * s + 1]);
                }
            }
        }
    }
    if (valid && fvalid) {
#pragma unroll
        for (int s = 0; s < ATT_SMAX; ++s) {
            if (s < e.S) {
                const size_t o = (((size_t)b * e.S + s) * e.T + t) * e.F + f;
                if (!e.crm) {
                    e.out[o] = sigmoid_f(en[2 * s]);
                } else {
                    reinterpret_cast<float2 *>(e.out)[o] =
                        make_float2(crm_value_tc(en[2 * s], e.crm_k, e.crm_c),
                                    crm_value_tc(en[2 * s + 1], e.crm_k, e.crm_c));
                }
            }
        }
    }
}

// the epilogue a K split of a tile runs: only split 0 adds the bias
template <int ACT>
__device__ __forceinline__ const EpiPlain<ACT, false> &epi_of_split(const EpiPlain<ACT, false> &e, int) { return e; }
template <int ACT>
__device__ __forceinline__ EpiPlain<ACT, true> epi_of_split(const EpiPlain<ACT, true> &e, int ks) {
    EpiPlain<ACT, true> r = e;
    if (ks != 0) r.bias = nullptr;
    return r;
}
__device__ __forceinline__ const EpiAttn &epi_of_split(const EpiAttn &e, int) { return e; }

// ------------------------------------------------------------------------------------------ kernel
// Work items are (tile, K split): nsplit > 1 cuts the k-blocks of every tile into nsplit ranges whose partial
// accumulators meet in C through fp32 atomics -- for the backward contractions over B*T rows whose M x N output is a
// handful of tiles (dW = dY^T X: 57 or 20 tiles on 148 SMs).
// MN = true: both operands are MN-major ("transposed" GEMM, C[m][n] = sum_r A[r][m] * B[r][n]: the weight gradients
// dW = dY^T X straight from the row-major bf16 planes of dY and X, no transposing split).  The contraction index r runs
// over (utterance b, frame t): a k-block is 64 frames of one utterance, fetched as 64 x 64 boxes of a 4-D tensor map
// (column, t, b, plane) -- frames past T arrive as zeros, and a per-operand frame shift (mn.shift_a / shift_b: the
// recurrent weight gradients pair dgates[t] with h[t-1] or h[t+1]) is just a coordinate offset whose out-of-range rows
// read zeros as well.  In shared memory an operand tile is [64-column atom][64 frames][128 B] (128B swizzle), which is the
// canonical MN-major UMMA layout: leading-dimension offset = one atom (8 KB), stride offset = 8 frames (1 KB).
struct MnMode {
    int tchunks;      // 64-frame chunks per utterance
    int shift_a, shift_b;
    int col0_a, col0_b;   // first column of the operand inside its plane rows (a coordinate offset: no alignment demands)
};
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(8192 >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}

template <typename Epi, bool MN = false>
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_bf16x3_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                   int M, int N, int kblocks, int m_tiles, int n_tiles, int nsplit, unsigned *sched, const Epi epi,
                   const MnMode mn) {
    constexpr int NSTEP = EpiTraits<Epi>::NSTEP;
    extern __shared__ unsigned char smem_raw[];
    unsigned char *tiles = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t *bars = reinterpret_cast<uint64_t *>(tiles + (size_t)TSTAGES * TSTAGE_BYTES);
    uint64_t *full = bars, *empty = bars + TSTAGES, *tfull = bars + 2 * TSTAGES, *tempty = bars + 2 * TSTAGES + 2;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * TSTAGES + 4);
    // work-item ring: the producer lane publishes the item ids (dynamic: taken from the launch's global counter, so CTAs
    // that become resident late -- next to another stream's recurrent launch -- simply find less work; static: blockIdx +
    // k * gridDim), the MMA lane and the epilogue warps consume them in the same order
    uint64_t *sfull = bars + 2 * TSTAGES + 5, *sempty = sfull + TC_RING;
    volatile int *sched_w = reinterpret_cast<volatile int *>(sempty + TC_RING);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int total_work = m_tiles * n_tiles * nsplit;
    const int kper = (kblocks + nsplit - 1) / nsplit;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmap_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmap_b) : "memory");
        for (int s = 0; s < TSTAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 4 * EpiTraits<Epi>::PARTS); }
        for (int r = 0; r < TC_RING; ++r) { mbar_init(&sfull[r], 1); mbar_init(&sempty[r], 1 + 4 * EpiTraits<Epi>::PARTS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            int w = sched ? (int)atomicAdd(sched, 1u) : (int)blockIdx.x;
            for (int it = 0;; ++it) {
                const int slot = it & (TC_RING - 1);
                mbar_wait(&sempty[slot], (((uint32_t)it / TC_RING) & 1u) ^ 1u);
                const bool live = w < total_work;
                sched_w[slot] = live ? w : -1;
                mbar_arrive(&sfull[slot]);
                if (!live) break;
                const int wnext = sched ? (int)atomicAdd(sched, 1u) : w + (int)gridDim.x;     // in flight while this item's k-blocks load
                const int tile = (nsplit == 1) ? w : w / nsplit, ks = w - tile * nsplit;
                const int m0 = (tile / n_tiles) * TBM, n0 = (tile % n_tiles) * NSTEP;
                const int kb0 = ks * kper, kb1 = min(kb0 + kper, kblocks);
                w = wnext;
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    unsigned char *st = tiles + (size_t)stage * TSTAGE_BYTES;
                    mbar_expect_tx(&full[stage], TSTAGE_BYTES);
                    if constexpr (MN) {
                        const int b = kb / mn.tchunks, t0 = (kb - b * mn.tchunks) * TBK;
#pragma unroll
                        for (int pl = 0; pl < 2; ++pl) {
#pragma unroll
                            for (int j = 0; j < TBM / 64; ++j)
                                tma_load_4d(st + pl * TA_BYTES + j * 8192, &tmap_a, &full[stage], mn.col0_a + m0 + 64 * j, t0 + mn.shift_a, b, pl);
#pragma unroll
                            for (int j = 0; j < TBN / 64; ++j)
                                tma_load_4d(st + 2 * TA_BYTES + pl * TB_BYTES + j * 8192, &tmap_b, &full[stage], mn.col0_b + n0 + 64 * j,
                                            t0 + mn.shift_b, b, pl);
                        }
                    } else {
                        tma_load_3d(st, &tmap_a, &full[stage], kb * TBK, m0, 0);
                        tma_load_3d(st + TA_BYTES, &tmap_a, &full[stage], kb * TBK, m0, 1);
                        tma_load_3d(st + 2 * TA_BYTES, &tmap_b, &full[stage], kb * TBK, n0, 0);
                        tma_load_3d(st + 2 * TA_BYTES + TB_BYTES, &tmap_b, &full[stage], kb * TBK, n0, 1);
                    }
                    if (++stage == TSTAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_bf16(TBM, TBN) | (MN ? ((1u << 15) | (1u << 16)) : 0u);   // bits 15 / 16: A / B MN-major
            int stage = 0; uint32_t phase = 0; int acc = 0; uint32_t acc_phase = 0;
            for (int it = 0;; ++it) {
                const int slot = it & (TC_RING - 1);
                mbar_wait(&sfull[slot], ((uint32_t)it / TC_RING) & 1u);
                const int w = sched_w[slot];
                mbar_arrive(&sempty[slot]);
                if (w < 0) break;
                const int ks = (nsplit == 1) ? 0 : w % nsplit;
                const int kb0 = ks * kper, kb1 = min(kb0 + kper, kblocks);
                mbar_wait(&tempty[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d = tmem_base + acc * TBN;
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(tiles + (size_t)stage * TSTAGE_BYTES);
                    const uint64_t a_hi = MN ? umma_desc_mn_sw128(sa) : umma_desc_sw128(sa);
                    const uint64_t a_lo = MN ? umma_desc_mn_sw128(sa + TA_BYTES) : umma_desc_sw128(sa + TA_BYTES);
                    const uint64_t b_hi = MN ? umma_desc_mn_sw128(sa + 2 * TA_BYTES) : umma_desc_sw128(sa + 2 * TA_BYTES);
                    const uint64_t b_lo = MN ? umma_desc_mn_sw128(sa + 2 * TA_BYTES + TB_BYTES) : umma_desc_sw128(sa + 2 * TA_BYTES + TB_BYTES);
                    constexpr int KSTEP = MN ? (2048 >> 4) : 2;  // per 16-element k step: K-major +32 B, MN-major +16 frames x 128 B
#pragma unroll
                    for (int k = 0; k < TBK / 16; ++k) {
                        umma_bf16(d, a_hi + KSTEP * k, b_lo + KSTEP * k, idesc, ((kb - kb0) | k) != 0);
                        umma_bf16(d, a_lo + KSTEP * k, b_hi + KSTEP * k, idesc, 1);
                        umma_bf16(d, a_hi + KSTEP * k, b_hi + KSTEP * k, idesc, 1);
                    }
                    umma_commit(&empty[stage]);
                    if (++stage == TSTAGES) { stage = 0; phase ^= 1; }
                }
                umma_commit(&tfull[acc]);
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1;
            }
        }
    } else if (((warp - 2) >> 2) < EpiTraits<Epi>::PARTS) {
        const int quarter = warp & 3;                 // TMEM lanes this warp may touch: 32*quarter ..
        const int part = (warp - 2) >> 2;             // column part of the tile
        float *stage = reinterpret_cast<float *>(tiles + (size_t)TSTAGES * TSTAGE_BYTES + 256) +
                       (size_t)((part * 4 + quarter) % TC_STORE_WARPS) * 32 * TC_STAGE_PITCH;
        int acc = 0; uint32_t acc_phase = 0;
        for (int it = 0;; ++it) {
            const int slot = it & (TC_RING - 1);
            mbar_wait(&sfull[slot], ((uint32_t)it / TC_RING) & 1u);
            const int w = sched_w[slot];
            __syncwarp();
            if (lane == 0) mbar_arrive(&sempty[slot]);
            if (w < 0) break;
            const int tile = (nsplit == 1) ? w : w / nsplit, ks = w - tile * nsplit;
            const int m0 = (tile / n_tiles) * TBM, n0 = (tile % n_tiles) * NSTEP;
            mbar_wait(&tfull[acc], acc_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * TBN;
            epilogue_tile(epi_of_split(epi, ks), taddr, m0 + quarter * 32, n0, M, N, part, lane, stage);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[acc]);
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
    if (sched != nullptr && threadIdx.x == 0) {      // the last CTA out re-arms the launch's counter pair (graph replays reuse it)
        if (atomicAdd(sched + 1, 1u) == gridDim.x - 1) {
            sched[0] = 0u;
            sched[1] = 0u;
            __threadfence();
        }
    }
}


// ------------------------------------------------------------------------------------------ 2-CTA kernel
// The plain projection is bound by the SM-to-L2 port, not by the tensor pipe: a k-block moves 96 KB of operands per SM
// (A hi/lo 32 KB + W hi/lo 64 KB) next to the 131 KB fp32 output tile.  Here two CTAs of a cluster (the two SMs of a TPC)
// work on ONE 256 x 256 tile with tcgen05.mma.cta_group::2: each CTA loads its own 128 rows of A and only HALF of the W tile
// (128 of the 256 columns), the pair's tensor cores read both halves (the peer's through distributed shared memory): 64 KB of
// operands per SM and k-block, three stages instead of two.  Roles per CTA as in the 1-CTA kernel; differences:
//   * TMA loads carry .cta_group::2 and signal the LEADER's (cluster rank 0) `full` barrier: its address with the peer bit
//     cleared; the leader expects the bytes of both CTAs;
//   * only the leader's MMA lane issues UMMAs (M = 256: 128 accumulator rows in each CTA's tensor memory); its commits are
//     multicast to the `empty` / `tfull` barriers of both CTAs;
//   * the epilogue warps of both CTAs arrive on the leader's `tempty` (remote mbarrier arrive);
//   * tensor memory is allocated with .cta_group::2 by the same warp of both CTAs; cluster barriers bracket the kernel so that
//     no CTA exits while its peer may still touch its shared memory.
constexpr int T2_STAGES = 3;
constexpr int T2_BHALF = TBN / 2;                                 // W rows (output columns) each CTA holds
constexpr int T2_B_BYTES = T2_BHALF * TBK * 2;                    // 16 KB
constexpr int T2_STAGE_BYTES = 2 * TA_BYTES + 2 * T2_B_BYTES;     // 64 KB per CTA
constexpr size_t T2_SMEM = (size_t)T2_STAGES * T2_STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/ +
                           (size_t)TC_STORE_WARPS * 32 * TC_STAGE_PITCH * sizeof(float);
constexpr uint32_t PEER_MASK = 0xFEFFFFFFu;                       // clears the CTA-rank bit of a shared::cluster address

__device__ __forceinline__ void tma_load_3d_2sm(void *smem, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(smem)), "l"((uint64_t)map), "r"(smem_u32(bar) & PEER_MASK), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void umma_bf16_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void umma_commit_2sm(uint64_t *bar) {          // arrives on `bar` of BOTH CTAs of the pair
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint64_t *bar) {       // the same barrier in the pair's rank-0 CTA
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & PEER_MASK) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

template <typename Epi>
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm2_bf16x3_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                    int M, int N, int kblocks, int m_tiles2, int n_tiles, const Epi epi) {
    constexpr int NSTEP = EpiTraits<Epi>::NSTEP;
    extern __shared__ unsigned char smem_raw[];
    unsigned char *tiles = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t *bars = reinterpret_cast<uint64_t *>(tiles + (size_t)T2_STAGES * T2_STAGE_BYTES);
    uint64_t *full = bars, *empty = bars + T2_STAGES, *tfull = bars + 2 * T2_STAGES, *tempty = bars + 2 * T2_STAGES + 2;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * T2_STAGES + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
    const int total_work = m_tiles2 * n_tiles;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmap_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmap_b) : "memory");
        for (int s = 0; s < T2_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 2 * 4 * EpiTraits<Epi>::PARTS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                       // the peer's barriers are initialised before anything is signalled on them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (elect_one_lane()) {
            int stage = 0; uint32_t phase = 0;
            for (int w = pair; w < total_work; w += npairs) {
                const int m0 = (w / n_tiles) * (2 * TBM) + (int)rank * TBM, n0 = (w % n_tiles) * NSTEP + (int)rank * T2_BHALF;
                for (int kb = 0; kb < kblocks; ++kb) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    unsigned char *st = tiles + (size_t)stage * T2_STAGE_BYTES;
                    if (rank == 0) mbar_expect_tx(&full[stage], 2 * T2_STAGE_BYTES);       // both CTAs' bytes land on this barrier
                    tma_load_3d_2sm(st, &tmap_a, &full[stage], kb * TBK, m0, 0);
                    tma_load_3d_2sm(st + TA_BYTES, &tmap_a, &full[stage], kb * TBK, m0, 1);
                    tma_load_3d_2sm(st + 2 * TA_BYTES, &tmap_b, &full[stage], kb * TBK, n0, 0);
                    tma_load_3d_2sm(st + 2 * TA_BYTES + T2_B_BYTES, &tmap_b, &full[stage], kb * TBK, n0, 1);
                    if (++stage == T2_STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (rank == 0 && elect_one_lane()) {
            constexpr uint32_t idesc = umma_idesc_bf16(2 * TBM, TBN);
            int stage = 0; uint32_t phase = 0; int acc = 0; uint32_t acc_phase = 0;
            for (int w = pair; w < total_work; w += npairs) {
                mbar_wait(&tempty[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d = tmem_base + acc * TBN;
                for (int kb = 0; kb < kblocks; ++kb) {
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(tiles + (size_t)stage * T2_STAGE_BYTES);
                    const uint64_t a_hi = umma_desc_sw128(sa), a_lo = umma_desc_sw128(sa + TA_BYTES);
                    const uint64_t b_hi = umma_desc_sw128(sa + 2 * TA_BYTES), b_lo = umma_desc_sw128(sa + 2 * TA_BYTES + T2_B_BYTES);
#pragma unroll
                    for (int k = 0; k < TBK / 16; ++k) {
                        umma_bf16_2sm(d, a_hi + 2 * k, b_lo + 2 * k, idesc, (kb | k) != 0);
                        umma_bf16_2sm(d, a_lo + 2 * k, b_hi + 2 * k, idesc, 1);
                        umma_bf16_2sm(d, a_hi + 2 * k, b_hi + 2 * k, idesc, 1);
                    }
                    umma_commit_2sm(&empty[stage]);
                    if (++stage == T2_STAGES) { stage = 0; phase ^= 1; }
                }
                umma_commit_2sm(&tfull[acc]);
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1;
            }
        }
    } else if (((warp - 2) >> 2) < EpiTraits<Epi>::PARTS) {
        const int quarter = warp & 3;
        const int part = (warp - 2) >> 2;
        float *stage = reinterpret_cast<float *>(tiles + (size_t)T2_STAGES * T2_STAGE_BYTES + 256) +
                       (size_t)((part * 4 + quarter) % TC_STORE_WARPS) * 32 * TC_STAGE_PITCH;
        int acc = 0; uint32_t acc_phase = 0;
        for (int w = pair; w < total_work; w += npairs) {
            const int m0 = (w / n_tiles) * (2 * TBM) + (int)rank * TBM, n0 = (w % n_tiles) * NSTEP;
            mbar_wait(&tfull[acc], acc_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * TBN;
            epilogue_tile(epi, taddr, m0 + quarter * 32, n0, M, N, part, lane, stage);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_leader(&tempty[acc]);
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1;
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                       // the peer is done with this CTA's shared memory, barriers and tensor memory
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

// ------------------------------------------------------------------------------------------ host
// Counter pairs {next item, CTAs finished} of the dynamically scheduled launches, zeroed once; a launch takes one pair,
// its last CTA re-zeroes it.  Eager launches cycle through a ring (far longer than any launch queue); launches recorded
// into a CUDA graph keep their pair for the graph's life (replays of a node never overlap themselves), so those come
// from a second region that is handed out once -- when it is used up, further captured launches fall back to the static
// schedule (sched == nullptr).
constexpr unsigned SCHED_EAGER = 4096, SCHED_GRAPH = 8192;
struct SchedPool {
    std::mutex mu;
    unsigned *base = nullptr;
    unsigned eager = 0, graph = 0;
};
static SchedPool g_sched[DL4SS_MAX_DEVICES];
static bool g_dynamic = [] { const char *e = getenv("DL4SS_GEMM_DYNAMIC"); return !(e && e[0] == '0'); }();

static unsigned *sched_pair(cudaStream_t st) {
    if (!g_dynamic) return nullptr;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= DL4SS_MAX_DEVICES) return nullptr;
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cs) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    SchedPool &p = g_sched[dev];
    std::lock_guard<std::mutex> lk(p.mu);
    if (p.base == nullptr) {
        if (cs != cudaStreamCaptureStatusNone) return nullptr;          // no allocation while a capture is open
        const size_t bytes = (size_t)2 * (SCHED_EAGER + SCHED_GRAPH) * sizeof(unsigned);
        unsigned *b = nullptr;
        if (cudaMalloc(&b, bytes) != cudaSuccess || cudaMemset(b, 0, bytes) != cudaSuccess) { cudaGetLastError(); return nullptr; }
        p.base = b;
    }
    if (cs != cudaStreamCaptureStatusNone) {
        if (p.graph >= SCHED_GRAPH) return nullptr;
        return p.base + 2 * (size_t)(SCHED_EAGER + p.graph++);
    }
    return p.base + 2 * (size_t)(p.eager++ % SCHED_EAGER);
}

EncodeTiledFn tensor_map_encoder() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    });
    return fn;
}

int make_bf16_map(CUtensorMap *map, const void *base, int rank, const cuuint64_t *dims, const cuuint64_t *strides,
                  const cuuint32_t *box) {
    EncodeTiledFn fn = tensor_map_encoder();
    if (!fn) { set_error("cuTensorMapEncodeTiled is unavailable (driver too old?)"); return DL4SS_ECUDA; }
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void *>(base), dims, strides,
                    box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d), rank %d", (int)r, rank); return DL4SS_ECUDA; }
    return DL4SS_OK;
}

int make_f32_map(CUtensorMap *map, const void *base, int rank, const cuuint64_t *dims, const cuuint64_t *strides,
                 const cuuint32_t *box) {
    EncodeTiledFn fn = tensor_map_encoder();
    if (!fn) { set_error("cuTensorMapEncodeTiled is unavailable (driver too old?)"); return DL4SS_ECUDA; }
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<void *>(base), dims, strides,
                    box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (f32) failed (%d), rank %d", (int)r, rank); return DL4SS_ECUDA; }
    return DL4SS_OK;
}

// planes: bf16 [2][R][Kp]; box = 64 (K) x rows x 1 plane, 128B swizzle, OOB rows zero-filled
static int make_plane_map(CUtensorMap *map, const void *planes, long long R, int Kp, int box_rows, int pitch = 0, int kvalid = 0) {
    if (pitch <= 0) pitch = Kp;              // row pitch in elements; columns >= kvalid (default: the padded Kp) read as zeros
    if (kvalid <= 0) kvalid = Kp;
    cuuint64_t dims[3] = {(cuuint64_t)kvalid, (cuuint64_t)R, 2};
    cuuint64_t strides[2] = {(cuuint64_t)pitch * 2, (cuuint64_t)R * pitch * 2};
    cuuint32_t box[3] = {(cuuint32_t)TBK, (cuuint32_t)box_rows, 1};
    return make_bf16_map(map, planes, 3, dims, strides, box);
}

// K splits that minimise (waves of work items) x (k-blocks per item) on `sms` SMs; every split keeps >= 4 k-blocks
// and no split is empty
static int pick_ksplit(long long tiles, int kblocks, int sms) {
    int best = 1;
    long long best_cost = cdivll(tiles, sms) * kblocks;
    for (int ns = 2; ns <= 32 && kblocks / ns >= 4; ++ns) {
        const int kper = cdiv(kblocks, ns);
        if ((long long)kper * (ns - 1) >= kblocks) continue;          // the last split would be empty
        const long long cost = cdivll(tiles * ns, sms) * kper;
        if (cost * 100 < best_cost * 95) { best_cost = cost; best = ns; }   // a further split must buy >= 5 %
    }
    return best;
}

static int g_max_ctas = [] { const char *e = getenv("DL4SS_GEMM_MAX_CTAS"); return e ? atoi(e) : 0; }();

template <typename Epi>
static int launch_tc(const void *a_planes, const void *w_planes, int M, int N, int K, int n_tiles, const Epi &epi,
                     cudaStream_t st, int nsplit = 1, int lda = 0) {
    const int Kp = (K + TBK - 1) / TBK * TBK;
    CUtensorMap ma, mb;
    int rc = make_plane_map(&ma, a_planes, M, Kp, TBM, lda, lda > 0 ? K : 0);
    if (rc) return rc;
    rc = make_plane_map(&mb, w_planes, N, Kp, TBN);        // rows past N (and past a tile's own rows) are ignored / zero
    if (rc) return rc;
    const int m_tiles = cdiv(M, TBM);
    const long long total = (long long)m_tiles * n_tiles * nsplit;
    int grid = sm_count();
    if (g_max_ctas > 0 && g_max_ctas < grid) grid = g_max_ctas;
    if (total < grid) grid = (int)total;
    auto kern = gemm_bf16x3_kernel<Epi>;
    DL4SS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM));
    kern<<<grid, TC_THREADS, TC_SMEM, st>>>(ma, mb, M, N, Kp / TBK, m_tiles, n_tiles, nsplit, sched_pair(st), epi, MnMode{0, 0, 0, 0, 0});
    DL4SS_LAUNCH_CHECK("gemm_bf16x3_kernel");
    return DL4SS_OK;
}

static int g_two_cta = [] { const char *e = getenv("DL4SS_GEMM_2CTA"); return e ? atoi(e) : 1; }();

// the 2-CTA (cta_group::2) form of launch_tc for the plain epilogue: used when a launch owns the GPU (no CTA cap: with two
// batches in flight the recurrent launch of the other batch leaves single SMs free, not TPC pairs) and the output is at
// least one 256-row tile pair per SM pair
template <typename Epi>
static int launch_tc2(const void *a_planes, const void *w_planes, int M, int N, int K, const Epi &epi, cudaStream_t st, int lda = 0,
                      int n_tiles_in = 0) {
    const int Kp = (K + TBK - 1) / TBK * TBK;
    CUtensorMap ma, mb;
    int rc = make_plane_map(&ma, a_planes, M, Kp, TBM, lda, lda > 0 ? K : 0);
    if (rc) return rc;
    rc = make_plane_map(&mb, w_planes, N, Kp, T2_BHALF);
    if (rc) return rc;
    const int m_tiles2 = cdiv(M, 2 * TBM), n_tiles = n_tiles_in > 0 ? n_tiles_in : cdiv(N, TBN);
    const long long total = (long long)m_tiles2 * n_tiles;
    int pairs = sm_count() / 2;
    if (g_max_ctas > 0 && g_max_ctas / 2 < pairs) pairs = g_max_ctas / 2;
    if (total < pairs) pairs = (int)total;
    auto kern = gemm2_bf16x3_kernel<Epi>;
    DL4SS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T2_SMEM));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * pairs);
    cfg.blockDim = dim3(TC_THREADS);
    cfg.dynamicSmemBytes = T2_SMEM;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    DL4SS_CUDA(cudaLaunchKernelEx(&cfg, kern, ma, mb, M, N, Kp / TBK, m_tiles2, n_tiles, epi));
    count_launch();
    return DL4SS_OK;
}
static bool use_two_cta(int M, int N) {
    return g_two_cta != 0 && (g_max_ctas == 0 || g_two_cta == 2) && (long long)cdiv(M, 2 * TBM) * cdiv(N, TBN) >= sm_count() / 2;
}

// planes: bf16 [2][B][T][ld] (row-major frames); columns [0, cols) of them; box = 64 columns x 64 frames of one utterance, one plane
static int make_mn_map(CUtensorMap *map, const void *planes, int cols, int ld, int B, int T) {
    cuuint64_t dims[4] = {(cuuint64_t)cols, (cuuint64_t)T, (cuuint64_t)B, 2};
    cuuint64_t strides[3] = {(cuuint64_t)ld * 2, (cuuint64_t)T * ld * 2, (cuuint64_t)B * T * ld * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)TBK, 1, 1};
    return make_bf16_map(map, planes, 4, dims, strides, box);
}

}  // namespace dl4ss

using namespace dl4ss;

extern "C" void dl4ss_gemm_tc_set_max_ctas(int ctas) { g_max_ctas = ctas; }
extern "C" void dl4ss_gemm_tc_set_two_cta(int mode) { g_two_cta = mode; }

extern "C" size_t dl4ss_split_bf16_bytes(long long R, int K) {
    if (R <= 0 || K <= 0) return 0;
    const size_t Kp = (size_t)(K + TBK - 1) / TBK * TBK;
    return 2 * (size_t)R * Kp * sizeof(__nv_bfloat16);
}

extern "C" int dl4ss_split_bf16(const float *x, int ld, long long R, int K, void *planes, void *stream) {
    DL4SS_CHECK_ARG(x && planes, "split_bf16: null operand");
    DL4SS_CHECK_ARG(R >= 0 && K >= 1 && ld >= K, "split_bf16: bad R/K/ld %lld/%d/%d", R, K, ld);
    DL4SS_CHECK_ARG((((uintptr_t)planes) & 15) == 0, "split_bf16: planes must be 16-byte aligned");
    if (R == 0) return DL4SS_OK;
    const int Kp = (K + TBK - 1) / TBK * TBK;
    const long long total = R * (Kp / 8);
    long long blocks = cdivll(total, 256);
    const long long cap = (long long)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    split_bf16_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, ld, R, K, Kp, (__nv_bfloat16 *)planes);
    DL4SS_LAUNCH_CHECK("split_bf16_kernel");
    return DL4SS_OK;
}

extern "C" int dl4ss_split_bf16_t(const float *x, long long ld, int R, int C, void *planes, void *stream) {
    if (R == 0 || C == 0) return DL4SS_OK;
    DL4SS_CHECK_ARG(x && planes, "split_bf16_t: null operand");
    DL4SS_CHECK_ARG(R >= 1 && C >= 1 && ld >= C, "split_bf16_t: bad R/C/ld %d/%d/%lld", R, C, ld);
    DL4SS_CHECK_ARG((((uintptr_t)planes) & 15) == 0, "split_bf16_t: planes must be 16-byte aligned");
    const int Rp = (R + TBK - 1) / TBK * TBK;
    dim3 grid(cdiv(C, 32), Rp / 32);
    DL4SS_CHECK_ARG(grid.y < 65536, "split_bf16_t: too many rows");
    split_bf16_t_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, ld, R, C, Rp, (__nv_bfloat16 *)planes);
    DL4SS_LAUNCH_CHECK("split_bf16_t_kernel");
    return DL4SS_OK;
}

extern "C" int dl4ss_linear_tc_tn_splitk_fwd(const void *a_planes, int lda, int wa, int col0_a, int shift_a,
                                             const void *b_planes, int ldb, int wb, int col0_b, int shift_b,
                                             float *C, int ldc, int M, int N, int B, int T, void *stream) {
    DL4SS_CHECK_ARG(a_planes && b_planes && C, "linear_tc_tn_splitk_fwd: null operand");
    DL4SS_CHECK_ARG(M >= 1 && N >= 1 && B >= 0 && T >= 1 && ldc >= N && col0_a >= 0 && col0_b >= 0 && col0_a + M <= wa &&
                    col0_b + N <= wb && wa <= lda && wb <= ldb,
                    "linear_tc_tn_splitk_fwd: bad M/N/B/T/lda/wa/col0_a/ldb/wb/col0_b/ldc %d/%d/%d/%d/%d/%d/%d/%d/%d/%d/%d",
                    M, N, B, T, lda, wa, col0_a, ldb, wb, col0_b, ldc);
    DL4SS_CHECK_ARG(lda % 8 == 0 && ldb % 8 == 0 && (((uintptr_t)a_planes) & 15) == 0 && (((uintptr_t)b_planes) & 15) == 0,
                    "linear_tc_tn_splitk_fwd: plane rows must be 16-byte aligned (pitch a multiple of 8 elements)");
    DL4SS_CHECK_ARG(col0_a % 8 == 0 && col0_b % 8 == 0,
                    "linear_tc_tn_splitk_fwd: col0_a / col0_b must be multiples of 8 (a TMA box starts on a 16-byte boundary)");
    cudaStream_t st = (cudaStream_t)stream;
    DL4SS_CUDA(cudaMemset2DAsync(C, (size_t)ldc * sizeof(float), 0, (size_t)N * sizeof(float), (size_t)M, st));   // the K splits meet in C through fp32 atomics
    if (B == 0) return DL4SS_OK;
    static const bool g_mn_swap = [] { const char *e = getenv("DL4SS_MN_SWAP"); return !(e && e[0] == '0'); }();
    // C = A^T B or, when that pads fewer tiles (M is cut in 128s, N in 256s: 1200 x 600 is 30 tiles one way, 25 the other;
    // 1200 x 300 is 20 against 15), C^T = B^T A with the epilogue adding the tile transposed into the same C
    const bool swap = (long long)cdiv(N, TBM) * cdiv(M, TBN) < (long long)cdiv(M, TBM) * cdiv(N, TBN) && g_mn_swap;
    if (swap) {
        std::swap(a_planes, b_planes); std::swap(lda, ldb); std::swap(wa, wb); std::swap(col0_a, col0_b);
        std::swap(shift_a, shift_b); std::swap(M, N);
    }
    CUtensorMap ma, mb;
    int rc = make_mn_map(&ma, a_planes, wa, lda, B, T);
    if (rc) return rc;
    rc = make_mn_map(&mb, b_planes, wb, ldb, B, T);
    if (rc) return rc;
    const int tchunks = cdiv(T, TBK);
    const int kblocks = B * tchunks;
    const int m_tiles = cdiv(M, TBM), n_tiles = cdiv(N, TBN);
    const int ns = pick_ksplit((long long)m_tiles * n_tiles, kblocks, sm_count());
    const long long total = (long long)m_tiles * n_tiles * ns;
    int grid = sm_count();
    if (g_max_ctas > 0 && g_max_ctas < grid) grid = g_max_ctas;
    if (total < grid) grid = (int)total;
    auto kern = gemm_bf16x3_kernel<EpiPlain<DL4SS_ACT_NONE, true>, true>;
    DL4SS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM));
    kern<<<grid, TC_THREADS, TC_SMEM, st>>>(ma, mb, M, N, kblocks, m_tiles, n_tiles, ns, sched_pair(st),
                                            EpiPlain<DL4SS_ACT_NONE, true>{C, nullptr, ldc, swap ? 1 : 0}, MnMode{tchunks, shift_a, shift_b, col0_a, col0_b});
    DL4SS_LAUNCH_CHECK("gemm_bf16x3_kernel<MN>");
    return DL4SS_OK;
}

extern "C" int dl4ss_linear_tc_fwd(const void *a_planes, const void *w_planes, const float *bias, float *C, int ldc,
                                   int M, int N, int K, int act, void *stream) {
    DL4SS_CHECK_ARG(a_planes && w_planes && C, "linear_tc_fwd: null operand");
    DL4SS_CHECK_ARG(M >= 0 && N >= 1 && K >= 1 && ldc >= N, "linear_tc_fwd: bad M/N/K/ldc %d/%d/%d/%d", M, N, K, ldc);
    if (M == 0) return DL4SS_OK;
    DL4SS_CHECK_ARG(act >= DL4SS_ACT_NONE && act <= DL4SS_ACT_SIGMOID, "linear_tc_fwd: bad act %d", act);
    if (act == DL4SS_ACT_TANH && use_two_cta(M, N))
        return launch_tc2(a_planes, w_planes, M, N, K, EpiPlain<DL4SS_ACT_TANH>{C, bias, ldc}, (cudaStream_t)stream);
    if (act == DL4SS_ACT_TANH)
        return launch_tc(a_planes, w_planes, M, N, K, cdiv(N, TBN), EpiPlain<DL4SS_ACT_TANH>{C, bias, ldc}, (cudaStream_t)stream);
    if (act == DL4SS_ACT_SIGMOID)
        return launch_tc(a_planes, w_planes, M, N, K, cdiv(N, TBN), EpiPlain<DL4SS_ACT_SIGMOID>{C, bias, ldc}, (cudaStream_t)stream);
    if (use_two_cta(M, N)) return launch_tc2(a_planes, w_planes, M, N, K, EpiPlain<DL4SS_ACT_NONE>{C, bias, ldc}, (cudaStream_t)stream);
    return launch_tc(a_planes, w_planes, M, N, K, cdiv(N, TBN), EpiPlain<DL4SS_ACT_NONE>{C, bias, ldc}, (cudaStream_t)stream);
}

extern "C" int dl4ss_linear_tc_lda_fwd(const void *a_planes, int lda, const void *w_planes, const float *bias, float *C, int ldc,
                                       int M, int N, int K, void *stream) {
    DL4SS_CHECK_ARG(a_planes && w_planes && C, "linear_tc_lda_fwd: null operand");
    DL4SS_CHECK_ARG(M >= 0 && N >= 1 && K >= 1 && ldc >= N && lda >= K && lda % 8 == 0,
                    "linear_tc_lda_fwd: bad M/N/K/ldc/lda %d/%d/%d/%d/%d (lda: a multiple of 8 elements)", M, N, K, ldc, lda);
    if (M == 0) return DL4SS_OK;
    if (use_two_cta(M, N)) return launch_tc2(a_planes, w_planes, M, N, K, EpiPlain<DL4SS_ACT_NONE>{C, bias, ldc}, (cudaStream_t)stream, lda);
    return launch_tc(a_planes, w_planes, M, N, K, cdiv(N, TBN), EpiPlain<DL4SS_ACT_NONE>{C, bias, ldc}, (cudaStream_t)stream, 1, lda);
}

extern "C" int dl4ss_linear_tc_splitk_fwd(const void *a_planes, const void *w_planes, const float *bias, float *C,
                                          int ldc, int M, int N, int K, void *stream) {
    DL4SS_CHECK_ARG(a_planes && w_planes && C, "linear_tc_splitk_fwd: null operand");
    DL4SS_CHECK_ARG(M >= 0 && N >= 1 && K >= 1 && ldc >= N, "linear_tc_splitk_fwd: bad M/N/K/ldc %d/%d/%d/%d", M, N, K, ldc);
    if (M == 0) return DL4SS_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int n_tiles = cdiv(N, TBN);
    const int ns = pick_ksplit((long long)cdiv(M, TBM) * n_tiles, cdiv(K, TBK), sm_count());
    if (ns > 1) DL4SS_CUDA(cudaMemset2DAsync(C, (size_t)ldc * sizeof(float), 0, (size_t)N * sizeof(float), (size_t)M, st));
    if (ns > 1) return launch_tc(a_planes, w_planes, M, N, K, n_tiles, EpiPlain<DL4SS_ACT_NONE, true>{C, bias, ldc}, st, ns);
    return launch_tc(a_planes, w_planes, M, N, K, n_tiles, EpiPlain<DL4SS_ACT_NONE>{C, bias, ldc}, st, 1);
}

extern "C" int dl4ss_emb_attn_mask_tc_fwd(const void *h_planes, const void *w_planes, const float *bias,
                                          const float *q, int B, int T, int F, int E, int K, int S, int mode,
                                          float crm_k, float crm_c, float *mask_out, void *stream) {
    DL4SS_CHECK_ARG(h_planes && w_planes && bias && q && mask_out, "emb_attn_mask_tc_fwd: null operand");
    DL4SS_CHECK_ARG(B >= 0 && T >= 1 && F >= 1 && K >= 1 && S >= 1, "emb_attn_mask_tc_fwd: bad shape");
    DL4SS_CHECK_ARG(mode == DL4SS_ATT_DOT || mode == DL4SS_ATT_DOT_CRM, "emb_attn_mask_tc_fwd: bad mode %d", mode);
    if (E != ATT_E || S > ATT_SMAX) {
        set_error("emb_attn_mask_tc_fwd: fused epilogue is built for E=%d, S<=%d (got E=%d S=%d)", ATT_E, ATT_SMAX, E, S);
        return DL4SS_EUNSUPPORTED;
    }
    if (B == 0) return DL4SS_OK;
    DL4SS_CHECK_ARG((long long)B * T < (1ll << 31), "emb_attn_mask_tc_fwd: B*T too large");
    EpiAttn e{bias, q, mask_out, T, F, S, mode == DL4SS_ATT_DOT_CRM ? 1 : 0, crm_k, crm_c};
    static const int attn2 = [] { const char *v = getenv("DL4SS_ATTN_2CTA"); return v ? atoi(v) : 1; }();
    if (attn2 && use_two_cta(B * T, F * E)) return launch_tc2(h_planes, w_planes, B * T, F * E, K, e, (cudaStream_t)stream, 0, cdiv(F, ATT_BINS));
    return launch_tc(h_planes, w_planes, B * T, F * E, K, cdiv(F, ATT_BINS), e, (cudaStream_t)stream);
}
