// K3 (tensor-core form)  dl4ss_rnn_layer_tc_fwd : one bidirectional LSTM / GRU layer as ONE persistent
// kernel whose recurrent product h_{t-1} * W_hh^T runs on tcgen05 (bf16x3 split, fp32 TMEM accumulators).
//
// Replaces the T sequential cuDNN steps of nn.LSTM / nn.GRU (TDAA_beta/main_run_sstune_EvalVer.py:282-293,
// ...cRM_EvalVer.py:345-356).  Decomposition (H = 300 in every reference config):
//   * a CTA owns (direction, tile of 64 utterances, slice of 20 hidden units).  Its G*20 rows of W_hh
//     (hi and lo bf16 planes, 128B-swizzled K-major) are TMA-loaded ONCE and stay in shared memory for
//     all T steps; the cell state of its (row, unit) pairs stays in registers;
//   * per step: the 15 slice-CTAs of a (direction, tile) group exchange h_{t-1} through an L2-resident
//     bf16 hi/lo ping-pong buffer: a loader thread spins on the group's release counter, then TMA-loads
//     the 64 x H tile (5 k-chunks, one mbarrier each); the MMA thread issues 3 UMMAs (M=64, N=G*20) per
//     16-wide k step as the chunks land; 8 epilogue warps pull the accumulators out of TMEM, add the
//     hoisted input projection (cp.async-prefetched one step ahead by 2 producer warps), apply the
//     gates, and publish h_t (fp32 into y, bf16 hi/lo into the exchange buffer) + release-increment;
//   * groups never wait on each other; the launch is cooperative so every CTA is resident.
// The only HBM traffic is the one read of xproj and the one write of y.
#include "tc_ptx.cuh"

namespace dl4ss {

constexpr int RT_BT = 64;                      // utterances per tile = UMMA M
constexpr int RT_HS = 20;                      // hidden units per slice
constexpr int RT_EU = RT_HS / 2;               // units per epilogue thread (two column halves)
constexpr int RT_KC = 64;                      // k per chunk (128 B of bf16: one swizzle row)
constexpr int RT_MAXKC = 5;                    // H <= 320
constexpr int RT_XP = 84;                      // xproj smem row pitch in floats (conflict-free LDS.128/64)
constexpr int RT_EPI_WARPS = 8;
constexpr int RT_PRE_THREADS = 64;             // xproj prefetch threads (warps 2,3)
constexpr int RT_THREADS = 32 * (4 + RT_EPI_WARPS);
constexpr int RT_HBLK = RT_BT * 128;           // bytes of one (k-chunk, plane) h block

struct RnnTcParams {
    const float *xproj;        // [B,T,2,G*H]
    const float *bhn;          // [2,H] (GRU) or null
    float *y;                  // [B,T,2H]
    float *gates_save;         // [B,T,2,G*H] or null
    float *cell_save;          // [B,T,2,H] or null
    __nv_bfloat16 *hbuf;       // [2 ping-pong][2 dir][2 plane][Bpad][Kp]
    unsigned *counters;        // [2][tiles_total]
    int B, T, H, Kp, nkc, Bpad;
    int tile0, tiles, tiles_total, nslices;
    long long *trace;          // optional [steps][16] clock stamps of CTA 0 (profiling hook), else null
    int trace_steps;
};

__device__ __forceinline__ void cp_async16_cg(void *smem, const void *gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem_u32(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// relaxed: the caller has just executed __threadfence() (fence + relaxed atomic = release pattern);
// red.release would pay for a second MEMBAR
__device__ __forceinline__ void red_relaxed_gpu_add(unsigned *p, unsigned v) {
    asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;\n" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void stamp(const RnnTcParams &p, int s, int slot) {
    if (p.trace != nullptr && blockIdx.x == 0 && s < p.trace_steps) p.trace[s * 16 + slot] = clock64();
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;\n" ::: "memory"); }

// 10 consecutive accumulator columns of this thread's TMEM lane (x8 + x2), then wait
__device__ __forceinline__ void tmem_ld10(uint32_t taddr, float (&v)[RT_EU]) {
    uint32_t r[10];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%10];\n\t"
        "tcgen05.ld.sync.aligned.32x32b.x2.b32 {%8, %9}, [%11];\n\t"
        "tcgen05.wait::ld.sync.aligned;\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9])
        : "r"(taddr), "r"(taddr + 8)
        : "memory");
#pragma unroll
    for (int i = 0; i < 10; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ uint32_t pack_bf16x2(__nv_bfloat16 a, __nv_bfloat16 b) {
    return (uint32_t)__bfloat16_as_ushort(a) | ((uint32_t)__bfloat16_as_ushort(b) << 16);
}

template <int CELL>
__global__ void __launch_bounds__(RT_THREADS, 1)
rnn_tc_kernel(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_h,
              const RnnTcParams p) {
    constexpr int G = (CELL == DL4SS_CELL_LSTM) ? 4 : 3;
    constexpr int NCOL = G * RT_HS;                 // accumulator columns in use: 80 / 60
    constexpr int UN = (NCOL + 7) / 8 * 8;          // UMMA N: 80 / 64
    constexpr int WBLK = UN * 128;                  // bytes of one (k-chunk, plane) W block (multiple of 1024)
    constexpr uint32_t IDESC2 = umma_idesc_bf16(RT_BT, 2 * UN);   // h_hi x [W_hi ; W_lo] in one instruction
    constexpr uint32_t IDESC1 = umma_idesc_bf16(RT_BT, UN);       // h_lo x W_hi
    static_assert(RT_EU == 10, "tmem_ld10 / store loops are written for 10 units per thread");

    extern __shared__ unsigned char smem_raw[];
    unsigned char *base = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int nkc = p.nkc;
    unsigned char *Wsm = base;                                        // [nkc][2][WBLK]
    unsigned char *Hsm = Wsm + (size_t)nkc * 2 * WBLK;                // [nkc][2][RT_HBLK]
    float *Xsm = reinterpret_cast<float *>(Hsm + (size_t)nkc * 2 * RT_HBLK);   // [2][RT_BT][RT_XP]
    uint64_t *bars = reinterpret_cast<uint64_t *>(Xsm + 2 * RT_BT * RT_XP);
    uint64_t *wfull = bars, *hfull = bars + 1, *tfull = bars + 1 + RT_MAXKC, *tempty = tfull + 1;
    uint64_t *xfull = tempty + 1, *xempty = xfull + 2;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(xempty + 2);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    int bid = blockIdx.x;
    const int slice = bid % p.nslices; bid /= p.nslices;
    const int bt = bid % p.tiles;
    const int dir = bid / p.tiles;
    const int tile = p.tile0 + bt;
    const int row0 = tile * RT_BT;                  // first utterance of the tile
    const int u0 = slice * RT_HS;
    const int T = p.T, H = p.H;
    const size_t GH = (size_t)G * H;
    unsigned *counter = p.counters + dir * p.tiles_total + tile;

    if (tid == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmap_w) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmap_h) : "memory");
        mbar_init(wfull, 1);
        for (int c = 0; c < RT_MAXKC; ++c) mbar_init(&hfull[c], 1);
        mbar_init(tfull, 1);
        mbar_init(tempty, RT_EPI_WARPS);
        for (int i = 0; i < 2; ++i) { mbar_init(&xfull[i], RT_PRE_THREADS); mbar_init(&xempty[i], RT_EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ================= loader: resident W slice once, then h_{t-1} tiles as the group publishes them
        if (lane == 0) {
            mbar_expect_tx(wfull, (uint32_t)(nkc * 2 * NCOL * 128));
            for (int c = 0; c < nkc; ++c)
                for (int pl = 0; pl < 2; ++pl)
                    tma_load_4d(Wsm + (size_t)(c * 2 + pl) * WBLK, &tmap_w, wfull, c * RT_KC, u0, dir * G, pl);
            for (int s = 1; s < T; ++s) {
                const unsigned want = (unsigned)(p.nslices * RT_EPI_WARPS) * (unsigned)s;
                stamp(p, s, 0);
                while (ld_acquire_gpu(counter) < want) { }
                stamp(p, s, 1);
                fence_proxy_async();                 // the group's generic-proxy stores -> this thread's TMA reads
                stamp(p, s, 15);
                const int pp = (s - 1) & 1;
                const int z = (pp * 2 + dir) * 2;
                for (int c = 0; c < nkc; ++c) {
                    mbar_expect_tx(&hfull[c], 2 * RT_HBLK);             // one box = both planes of the chunk
                    tma_load_3d(Hsm + (size_t)(c * 2) * RT_HBLK, &tmap_h, &hfull[c], c * RT_KC, row0, z);
                }
                stamp(p, s, 2);
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer: acc[64 x UN] = h_hi*W_lo + h_lo*W_hi + h_hi*W_hi
        if (lane == 0) {
            mbar_wait(wfull, 0);
            tc_fence_after();
            for (int s = 1; s < T; ++s) {
                if (s >= 2) { mbar_wait(tempty, (uint32_t)(s - 2) & 1u); tc_fence_after(); }
                for (int c = 0; c < nkc; ++c) {
                    mbar_wait(&hfull[c], (uint32_t)(s - 1) & 1u);
                    tc_fence_after();
                    if (c == 0) stamp(p, s, 3);
                    if (c == nkc - 1) stamp(p, s, 4);
                    const uint64_t a_hi = umma_desc_sw128(smem_u32(Hsm + (size_t)(c * 2) * RT_HBLK));
                    const uint64_t a_lo = umma_desc_sw128(smem_u32(Hsm + (size_t)(c * 2 + 1) * RT_HBLK));
                    const uint64_t b_hi = umma_desc_sw128(smem_u32(Wsm + (size_t)(c * 2) * WBLK));
                    int ksteps = (H - c * RT_KC + 15) / 16;
                    if (ksteps > RT_KC / 16) ksteps = RT_KC / 16;
                    for (int k = 0; k < ksteps; ++k) {      // +32 B per 16-element k step (>>4 = 2)
                        // an M=64 UMMA costs ~85 cycles here whatever N is, so the two products that share
                        // h_hi run as ONE instruction over the stacked [W_hi ; W_lo] rows (N = 2*UN):
                        // columns [0,UN) = hi*hi, [UN,2UN) = hi*lo, [2UN,3UN) = lo*hi
                        umma_bf16(tmem_base, a_hi + 2 * k, b_hi + 2 * k, IDESC2, (c | k) != 0);
                        umma_bf16(tmem_base + 2 * UN, a_lo + 2 * k, b_hi + 2 * k, IDESC1, (c | k) != 0);
                    }
                }
                umma_commit(tfull);
                stamp(p, s, 5);
            }
        }
    } else if (warp < 4) {
        // ================= xproj prefetch: rows of step s into Xsm[s&1], one step ahead of the epilogue
        const int pt = tid - 64;
        constexpr int V = RT_HS / 4;                 // 16-byte chunks per (row, gate)
        for (int s = 0; s < T; ++s) {
            const int buf = s & 1;
            if (s >= 2) mbar_wait(&xempty[buf], (uint32_t)((s >> 1) - 1) & 1u);
            const int t = dir ? (T - 1 - s) : s;
            float *dst = Xsm + (size_t)buf * RT_BT * RT_XP;
            for (int i = pt; i < RT_BT * G * V; i += RT_PRE_THREADS) {
                const int v = i % V, g = (i / V) % G, r = i / (V * G);
                const int b = row0 + r;
                if (b < p.B)
                    cp_async16_cg(dst + r * RT_XP + g * RT_HS + 4 * v,
                                  p.xproj + (((size_t)b * T + t) * 2 + dir) * GH + (size_t)g * H + u0 + 4 * v);
            }
            asm volatile("cp.async.commit_group;\n" ::: "memory");
            asm volatile("cp.async.wait_group 0;\n" ::: "memory");
            mbar_arrive(&xfull[buf]);
        }
    } else {
        // ================= epilogue: gates, state update, publish h_t
        const int ew = warp - 4;
        const int q = warp & 3;                      // TMEM sub-partition this warp may read
        const int ch = ew >> 2;                      // column half: units [10*ch, 10*ch+10) of the slice
        const int r = 16 * q + (lane & 15);          // M=64 accumulator row i lives in TMEM lane (i%16) + 32*(i/16)
        const int b = row0 + r;
        const bool active = (lane < 16) && (b < p.B);
        const int uc = u0 + RT_EU * ch;              // first hidden unit of this thread
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);

        float state[RT_EU];                          // LSTM: c ; GRU: h
        float bhn_r[RT_EU];
#pragma unroll
        for (int j = 0; j < RT_EU; ++j) {
            state[j] = 0.f;
            bhn_r[j] = (CELL == DL4SS_CELL_GRU) ? __ldg(p.bhn + (size_t)dir * H + uc + j) : 0.f;
        }

        for (int s = 0; s < T; ++s) {
            const int t = dir ? (T - 1 - s) : s;
            const int buf = s & 1;
            float acc[G][RT_EU];
            if (s >= 1) {
                mbar_wait(tfull, (uint32_t)(s - 1) & 1u);
                tc_fence_after();
                if (tid == 128) stamp(p, s, 6);
#pragma unroll
                for (int g = 0; g < G; ++g) {
                    float a1[RT_EU], a2[RT_EU];
                    tmem_ld10(taddr + UN + g * RT_HS + RT_EU * ch, a1);
                    tmem_ld10(taddr + 2 * UN + g * RT_HS + RT_EU * ch, a2);
                    tmem_ld10(taddr + g * RT_HS + RT_EU * ch, acc[g]);
#pragma unroll
                    for (int j = 0; j < RT_EU; ++j) acc[g][j] += a1[j] + a2[j];      // small terms first
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(tempty);
                if (tid == 128) stamp(p, s, 7);
            } else {
#pragma unroll
                for (int g = 0; g < G; ++g)
#pragma unroll
                    for (int j = 0; j < RT_EU; ++j) acc[g][j] = 0.f;
            }
            mbar_wait(&xfull[buf], (uint32_t)(s >> 1) & 1u);
            if (tid == 128) stamp(p, s, 8);
            const float *xr = Xsm + (size_t)buf * RT_BT * RT_XP + r * RT_XP + RT_EU * ch;
            float xv[G][RT_EU];
#pragma unroll
            for (int g = 0; g < G; ++g)
#pragma unroll
                for (int j = 0; j < RT_EU; j += 2) {
                    const float2 x2 = *reinterpret_cast<const float2 *>(xr + g * RT_HS + j);
                    xv[g][j] = x2.x; xv[g][j + 1] = x2.y;
                }
            __syncwarp();
            if (lane == 0) mbar_arrive(&xempty[buf]);

            float hnew[RT_EU], gv[G][RT_EU], aux[RT_EU];
#pragma unroll
            for (int j = 0; j < RT_EU; ++j) {
                if constexpr (CELL == DL4SS_CELL_LSTM) {
                    const float ig = sigmoid_f(xv[0][j] + acc[0][j]);
                    const float fg = sigmoid_f(xv[1][j] + acc[1][j]);
                    const float gg = tanh_f(xv[2][j] + acc[2][j]);
                    const float og = sigmoid_f(xv[3][j] + acc[3][j]);
                    const float c = fmaf(fg, state[j], ig * gg);
                    state[j] = c;
                    hnew[j] = og * tanh_f(c);
                    gv[0][j] = ig; gv[1][j] = fg; gv[2][j] = gg; gv[3][j] = og;
                    aux[j] = c;
                } else {
                    const float rg = sigmoid_f(xv[0][j] + acc[0][j]);
                    const float zg = sigmoid_f(xv[1][j] + acc[1][j]);
                    const float hn = acc[2][j] + bhn_r[j];
                    const float ng = tanh_f(fmaf(rg, hn, xv[2][j]));
                    hnew[j] = fmaf(zg, state[j] - ng, ng);          // (1-z)*n + z*h
                    state[j] = hnew[j];
                    gv[0][j] = rg; gv[1][j] = zg; gv[2][j] = ng;
                    aux[j] = hn;                                    // W_hn*h + b_hn, kept for backward
                }
            }
            if (tid == 128) stamp(p, s, 9);
            // publish h_t first (the group's next step hangs on it), the fp32 outputs follow off the critical path
            if (s + 1 < T) {
                if (active) {
                    const int pp = s & 1;
                    __nv_bfloat16 *hh = p.hbuf + ((size_t)((pp * 2 + dir) * 2) * p.Bpad + b) * p.Kp + uc;
                    __nv_bfloat16 *hl = hh + (size_t)p.Bpad * p.Kp;
                    uint32_t ph[RT_EU / 2], pl[RT_EU / 2];
#pragma unroll
                    for (int j = 0; j < RT_EU; j += 2) {
                        const __nv_bfloat16 h0 = __float2bfloat16_rn(hnew[j]), h1 = __float2bfloat16_rn(hnew[j + 1]);
                        const __nv_bfloat16 l0 = __float2bfloat16_rn(hnew[j] - __bfloat162float(h0));
                        const __nv_bfloat16 l1 = __float2bfloat16_rn(hnew[j + 1] - __bfloat162float(h1));
                        ph[j / 2] = pack_bf16x2(h0, h1);
                        pl[j / 2] = pack_bf16x2(l0, l1);
                    }
                    // 20 B per plane: 8+8+4 (column half 0, 8-byte aligned) or 4+8+8 (half 1): fewer L2 write
                    // transactions for the release fence to wait on
                    if (ch == 0) {
                        *reinterpret_cast<uint2 *>(hh) = make_uint2(ph[0], ph[1]);
                        *reinterpret_cast<uint2 *>(hh + 4) = make_uint2(ph[2], ph[3]);
                        *reinterpret_cast<uint32_t *>(hh + 8) = ph[4];
                        *reinterpret_cast<uint2 *>(hl) = make_uint2(pl[0], pl[1]);
                        *reinterpret_cast<uint2 *>(hl + 4) = make_uint2(pl[2], pl[3]);
                        *reinterpret_cast<uint32_t *>(hl + 8) = pl[4];
                    } else {
                        *reinterpret_cast<uint32_t *>(hh) = ph[0];
                        *reinterpret_cast<uint2 *>(hh + 2) = make_uint2(ph[1], ph[2]);
                        *reinterpret_cast<uint2 *>(hh + 6) = make_uint2(ph[3], ph[4]);
                        *reinterpret_cast<uint32_t *>(hl) = pl[0];
                        *reinterpret_cast<uint2 *>(hl + 2) = make_uint2(pl[1], pl[2]);
                        *reinterpret_cast<uint2 *>(hl + 6) = make_uint2(pl[3], pl[4]);
                    }
                }
                if (tid == 128) stamp(p, s, 10);
                __syncwarp();
                if (lane == 0) {                 // every epilogue warp releases its own rows: no CTA barrier
                    __threadfence();
                    if (tid == 128) stamp(p, s, 13);
                    red_relaxed_gpu_add(counter, 1u);
                    if (tid == 128) stamp(p, s, 14);
                }
            }
            if (active) {
                float *yo = p.y + ((size_t)b * T + t) * 2 * H + (size_t)dir * H + uc;
#pragma unroll
                for (int j = 0; j < RT_EU; j += 2) *reinterpret_cast<float2 *>(yo + j) = make_float2(hnew[j], hnew[j + 1]);
                if (p.gates_save != nullptr) {
                    float *go = p.gates_save + (((size_t)b * T + t) * 2 + dir) * GH + uc;
#pragma unroll
                    for (int g = 0; g < G; ++g)
#pragma unroll
                        for (int j = 0; j < RT_EU; j += 2)
                            *reinterpret_cast<float2 *>(go + (size_t)g * H + j) = make_float2(gv[g][j], gv[g][j + 1]);
                }
                if (p.cell_save != nullptr) {
                    float *co = p.cell_save + (((size_t)b * T + t) * 2 + dir) * H + uc;
#pragma unroll
                    for (int j = 0; j < RT_EU; j += 2) *reinterpret_cast<float2 *>(co + j) = make_float2(aux[j], aux[j + 1]);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256) : "memory");
    }
}

template <int CELL>
static int launch_rnn_tc(RnnTcParams p, const void *whh_planes, int rows_left, cudaStream_t st, int *launched_rows) {
    constexpr int G = (CELL == DL4SS_CELL_LSTM) ? 4 : 3;
    constexpr int UN = (G * RT_HS + 7) / 8 * 8;
    constexpr int WBLK = UN * 128;
    const size_t smem = 1024 + (size_t)p.nkc * 2 * WBLK + (size_t)p.nkc * 2 * RT_HBLK +
                        2ull * RT_BT * RT_XP * sizeof(float) + 256;
    auto kern = rnn_tc_kernel<CELL>;
    DL4SS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    DL4SS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, RT_THREADS, smem));
    const int max_tiles = per_sm * sm_count() / (2 * p.nslices);
    if (max_tiles < 1) {
        set_error("rnn_layer_tc_fwd: %d co-resident CTAs cannot hold one tile (%d slices x 2 directions)",
                  per_sm * sm_count(), p.nslices);
        return DL4SS_EUNSUPPORTED;
    }
    int tiles = cdiv(rows_left, RT_BT);
    if (tiles > max_tiles) tiles = max_tiles;
    p.tiles = tiles;
    *launched_rows = tiles * RT_BT;

    CUtensorMap mw, mh;
    {   // W planes bf16 [2 plane][2 dir * G][H][Kp]: box = 64 k x 20 units x G gates x 1 plane
        cuuint64_t dims[4] = {(cuuint64_t)p.Kp, (cuuint64_t)p.H, (cuuint64_t)(2 * G), 2};
        cuuint64_t strides[3] = {(cuuint64_t)p.Kp * 2, (cuuint64_t)p.H * p.Kp * 2, (cuuint64_t)2 * G * p.H * p.Kp * 2};
        cuuint32_t box[4] = {RT_KC, RT_HS, (cuuint32_t)G, 1};
        int rc = make_bf16_map(&mw, whh_planes, 4, dims, strides, box);
        if (rc) return rc;
    }
    {   // h exchange bf16 [8 = pp,dir,plane][Bpad][Kp]: box = 64 k x 64 rows
        cuuint64_t dims[3] = {(cuuint64_t)p.Kp, (cuuint64_t)p.Bpad, 8};
        cuuint64_t strides[2] = {(cuuint64_t)p.Kp * 2, (cuuint64_t)p.Bpad * p.Kp * 2};
        cuuint32_t box[3] = {RT_KC, RT_BT, 2};
        int rc = make_bf16_map(&mh, p.hbuf, 3, dims, strides, box);
        if (rc) return rc;
    }
    void *args[] = {(void *)&mw, (void *)&mh, (void *)&p};
    DL4SS_CUDA(cudaLaunchCooperativeKernel((void *)kern, dim3(2 * tiles * p.nslices), dim3(RT_THREADS), args, smem, st));
    count_launch();
    return DL4SS_OK;
}

static long long *g_trace = nullptr;
static int g_trace_steps = 0;

static bool rnn_tc_supported(int H) { return H >= RT_HS && H % RT_HS == 0 && H <= RT_MAXKC * RT_KC; }

}  // namespace dl4ss

using namespace dl4ss;

// profiling hook: device buffer of steps*16 int64 receiving CTA 0's per-phase clock64() stamps (null = off)
extern "C" void dl4ss_rnn_tc_set_trace(void *dev_buf, int steps) {
    g_trace = (long long *)dev_buf;
    g_trace_steps = dev_buf ? steps : 0;
}

extern "C" int dl4ss_rnn_tc_supported(int H, int cell) {
    return (cell == DL4SS_CELL_LSTM || cell == DL4SS_CELL_GRU) && rnn_tc_supported(H) ? 1 : 0;
}

extern "C" size_t dl4ss_rnn_tc_workspace_bytes(int B, int T, int H, int cell) {
    (void)T; (void)cell;
    if (B <= 0 || H <= 0) return 256;
    const size_t Bpad = (size_t)cdiv(B, RT_BT) * RT_BT;
    const size_t Kp = (size_t)cdiv(H, RT_KC) * RT_KC;
    const size_t ctr = ((size_t)2 * (Bpad / RT_BT) * sizeof(unsigned) + 255) / 256 * 256;
    return ctr + 8 * Bpad * Kp * sizeof(__nv_bfloat16);
}

extern "C" int dl4ss_rnn_layer_tc_fwd(int cell, const float *xproj, const void *whh_planes, const float *bhn,
                                      float *y, int B, int T, int H, float *gates_save, float *cell_save,
                                      void *workspace, size_t workspace_bytes, void *stream) {
    DL4SS_CHECK_ARG(cell == DL4SS_CELL_LSTM || cell == DL4SS_CELL_GRU, "rnn_layer_tc_fwd: bad cell %d", cell);
    DL4SS_CHECK_ARG(xproj && whh_planes && y, "rnn_layer_tc_fwd: null operand");
    DL4SS_CHECK_ARG(cell == DL4SS_CELL_LSTM || bhn, "rnn_layer_tc_fwd: GRU needs bhn");
    DL4SS_CHECK_ARG(B >= 0 && T >= 1 && H >= 1, "rnn_layer_tc_fwd: bad B/T/H %d/%d/%d", B, T, H);
    if (!rnn_tc_supported(H)) {
        set_error("rnn_layer_tc_fwd: H=%d unsupported (needs a multiple of %d, <= %d); use dl4ss_rnn_layer_fwd",
                  H, RT_HS, RT_MAXKC * RT_KC);
        return DL4SS_EUNSUPPORTED;
    }
    if (B == 0) return DL4SS_OK;
    const size_t need = dl4ss_rnn_tc_workspace_bytes(B, T, H, cell);
    if (!workspace || workspace_bytes < need) {
        set_error("rnn_layer_tc_fwd: workspace %zu B < %zu B", workspace_bytes, need);
        return DL4SS_EWORKSPACE;
    }
    DL4SS_CHECK_ARG((((uintptr_t)workspace) & 255) == 0, "rnn_layer_tc_fwd: workspace must be 256-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    DL4SS_CUDA(cudaMemsetAsync(workspace, 0, need, st));     // counters, and the zero k-padding / row padding of h
    RnnTcParams p;
    p.xproj = xproj; p.bhn = bhn; p.y = y; p.gates_save = gates_save; p.cell_save = cell_save;
    p.B = B; p.T = T; p.H = H;
    p.nkc = cdiv(H, RT_KC);
    p.Kp = p.nkc * RT_KC;
    p.Bpad = cdiv(B, RT_BT) * RT_BT;
    p.tiles_total = p.Bpad / RT_BT;
    p.nslices = H / RT_HS;
    p.counters = (unsigned *)workspace;
    const size_t ctr = ((size_t)2 * p.tiles_total * sizeof(unsigned) + 255) / 256 * 256;
    p.hbuf = (__nv_bfloat16 *)((unsigned char *)workspace + ctr);
    p.tiles = 0;
    p.trace = g_trace; p.trace_steps = g_trace_steps;
    int b0 = 0;
    while (b0 < B) {
        p.tile0 = b0 / RT_BT;
        int done = 0;
        int rc = (cell == DL4SS_CELL_LSTM) ? launch_rnn_tc<DL4SS_CELL_LSTM>(p, whh_planes, B - b0, st, &done)
                                           : launch_rnn_tc<DL4SS_CELL_GRU>(p, whh_planes, B - b0, st, &done);
        if (rc) return rc;
        b0 += done;
    }
    return DL4SS_OK;
}
