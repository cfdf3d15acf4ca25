// dl4ss_rnn_layer_bwd : the T-step BPTT chain of one bidirectional LSTM / GRU layer as ONE persistent kernel.
//
// Replaces what autograd does for nn.LSTM / nn.GRU under `loss.backward()` in the reference's training loop
// (TDAA_beta/main_run_sstune_EvalVer.py:673 ; cRM: TDAA_beta/main_run_sstune_cRM_EvalVer.py:751): per time step,
// walking each direction's forward order in reverse,
//     dh_t   = dy_t + dgates_{next} * W_hh            ([B,G*H] x [G*H,H], the only dense product on the chain)
//     dgates = gate arithmetic(dh_t, saved gates / cells, carried dc or dh*z)
// The weight / input gradients (dW_ih, dW_hh, dx) are large GEMMs over the saved dgates and run after the chain
// on the tensor-core projection kernel; only the chain itself is latency bound, so it lives here:
//   * a CTA owns (direction, tile of BW_BT utterances, slice of BW_HS hidden units).  The slice's columns of W_hh
//     (transposed: one row of G*H floats per unit) stay in shared memory for all T steps; the carried state
//     (LSTM dc*f, GRU dh*z) of its cells stays in registers of the cell's owner thread;
//   * per step the CTA reads the previous step's recurrent-side gate gradients of its tile -- the [B,T,2,G*H]
//     gradient array the weight GEMMs need anyway is the exchange buffer, freshly written by the sibling slices and
//     still in L2 -- with cp.async.cg, accumulates dh for its BW_BT x BW_HS cells on the fp32 pipe, applies the
//     gate arithmetic and writes this step's gate gradients;
//   * the product is register tiled 4 utterances x 5 units per lane (9 conflict-free 128-bit shared-memory loads feed
//     80 FMAs), the 16 tiles of the cell block sit in every warp, K = G*H is split over the warps and two half-warps,
//     partial sums meet through one shuffle and a shared-memory reduction by the owner threads;
//   * siblings of a (direction, tile) group synchronise through one L2 counter (release / acquire); groups never wait
//     on each other; the launch is cooperative so the waits cannot deadlock.
#include "tc_ptx.cuh"
#include <stdlib.h>

namespace dl4ss {

constexpr int BW_BT = 16;                 // utterances per tile
constexpr int BW_HS = 20;                 // hidden units per slice
constexpr int BW_CELLS = BW_BT * BW_HS;   // 320 cells per CTA, one owner thread each
constexpr int BW_NW = 20;                 // warps
constexpr int BW_THREADS = 32 * BW_NW;
constexpr int BW_RB = 4, BW_RU = 5;       // register tile: rows {rb + 4i}, units {ub + 4j}
constexpr int BW_CTR_STRIDE = 64;         // one 256-byte line per group counter

struct RnnBwdParams {
    const float *dy;        // [B,T,2H]
    const float *whh;       // [2,G*H,H]
    const float *gates;     // [B,T,2,G*H] saved activations
    const float *cells;     // [B,T,2,H]  LSTM c_t | GRU W_hn h + b_hn
    const float *y;         // [B,T,2H] layer output (GRU: h_{t-1})
    float *dgx;             // [B,T,2,G*H] d/d(xproj)
    float *dgh;             // GRU: d/d(W_hh h + b_hh); LSTM: null (= dgx)
    unsigned *counters;     // [2 * tiles] * BW_CTR_STRIDE
    int B, T, H, GHP;
    int b_begin, batch_tiles, nslices;
    long long *trace;       // optional [steps][8] clock stamps of CTA 0 (profiling hook), else null
    int trace_steps;
};

__device__ __forceinline__ void bw_stamp(const RnnBwdParams &p, int s, int slot) {
    if (p.trace != nullptr && blockIdx.x == 0 && threadIdx.x == 0 && s < p.trace_steps) p.trace[s * 8 + slot] = clock64();
}

__device__ __forceinline__ void bw_cp_async16(void *smem, const void *gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ unsigned bw_ld_acquire(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
    return v;
}

template <int CELL>
__global__ void __launch_bounds__(BW_THREADS, 1)
rnn_bwd_kernel(const RnnBwdParams p) {
    constexpr int G = (CELL == DL4SS_CELL_LSTM) ? 4 : 3;
    extern __shared__ __align__(16) float smem[];
    const int H = p.H, T = p.T, GHP = p.GHP;
    const int GH = G * H;
    const int nquads = GH >> 2;
    float *Wt = smem;                                   // [BW_HS][GHP]   Wt[j][k] = W_hh[k][u0 + j]
    float *dg = Wt + (size_t)BW_HS * GHP;               // [BW_BT][GHP]   previous step's recurrent-side gate grads
    float *part = dg + (size_t)BW_BT * GHP;             // [BW_NW][BW_CELLS] partial sums per warp

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    int bid = blockIdx.x;
    const int slice = bid % p.nslices; bid /= p.nslices;
    const int bt = bid % p.batch_tiles;
    const int dir = bid / p.batch_tiles;
    const int row0 = p.b_begin + bt * BW_BT;
    const int u0 = slice * BW_HS;
    unsigned *counter = p.counters + (size_t)(dir * p.batch_tiles + bt) * BW_CTR_STRIDE;
    const float *dgr = (CELL == DL4SS_CELL_GRU) ? p.dgh : p.dgx;   // what the recurrent product consumes

    // ---- resident transposed W_hh slice; zeroed exchange tile (rows beyond B and the pad columns stay zero)
    {
        const float *wsrc = p.whh + (size_t)dir * GH * H + u0;
        for (int i = tid; i < GH * BW_HS; i += BW_THREADS) {
            const int k = i / BW_HS, j = i - k * BW_HS;
            Wt[(size_t)j * GHP + k] = __ldg(wsrc + (size_t)k * H + j);
        }
        for (int i = tid; i < BW_HS * (GHP - GH); i += BW_THREADS) {
            const int j = i / (GHP - GH), k = GH + i - j * (GHP - GH);
            Wt[(size_t)j * GHP + k] = 0.f;
        }
        for (int i = tid; i < BW_BT * GHP; i += BW_THREADS) dg[i] = 0.f;
    }
    __syncthreads();

    // product mapping: lane = (ksub, rb, ub); tile rows rb + 4i (i < 4), units ub + 4j (j < 5)
    const int ub = lane & 3, rb = (lane >> 2) & 3, ksub = lane >> 4;
    const float *dgl = dg + (size_t)rb * GHP;
    const float *wl = Wt + (size_t)ub * GHP;
    // owner mapping: thread -> cell (r, j)
    const bool owner = tid < BW_CELLS;
    const int orow = tid / BW_HS, oj = tid - orow * BW_HS;
    const int ob = row0 + orow;
    const bool oval = owner && ob < p.B;
    const int ou = u0 + oj;
    float carry = 0.f;

    for (int s = 0; s < T; ++s) {
        const int t = dir ? s : (T - 1 - s);               // backward walks the forward order in reverse
        const int tp = dir ? t + 1 : t - 1;                // the step the forward pass came from
        const int tl = dir ? t - 1 : t + 1;                // the step this chain processed last
        const bool has_prev = (tp >= 0 && tp < T);

        // operands of the gate arithmetic do not depend on the chain: fetch them before waiting on it
        float dyv = 0.f, gv[G], cv = 0.f, pv = 0.f;
#pragma unroll
        for (int g = 0; g < G; ++g) gv[g] = 0.f;
        const size_t row = ((size_t)ob * T + t) * 2 + dir;
        if (oval) {
            dyv = __ldg(p.dy + ((size_t)ob * T + t) * 2 * H + (size_t)dir * H + ou);
#pragma unroll
            for (int g = 0; g < G; ++g) gv[g] = __ldg(p.gates + row * GH + (size_t)g * H + ou);
            cv = __ldg(p.cells + row * H + ou);
            if (has_prev) {
                if (CELL == DL4SS_CELL_LSTM) pv = __ldg(p.cells + (((size_t)ob * T + tp) * 2 + dir) * H + ou);
                else pv = __ldg(p.y + ((size_t)ob * T + tp) * 2 * H + (size_t)dir * H + ou);
            }
        }

        float dh = dyv;
        if (s > 0) {
            bw_stamp(p, s, 0);
            if (tid == 0) {
                const unsigned want = (unsigned)p.nslices * (unsigned)s;
                while (bw_ld_acquire(counter) < want) { __nanosleep(20); }
            }
            bw_stamp(p, s, 1);
            __syncthreads();
            for (int i = tid; i < BW_BT * nquads; i += BW_THREADS) {
                const int r = i / nquads, q = i - r * nquads;
                const int b = row0 + r;
                if (b < p.B)
                    bw_cp_async16(dg + (size_t)r * GHP + 4 * q, dgr + (((size_t)b * T + tl) * 2 + dir) * GH + 4 * q);
            }
            asm volatile("cp.async.commit_group;\n" ::: "memory");
            asm volatile("cp.async.wait_group 0;\n" ::: "memory");
            __syncthreads();
            bw_stamp(p, s, 2);

            float acc[BW_RB][BW_RU];
#pragma unroll
            for (int i = 0; i < BW_RB; ++i)
#pragma unroll
                for (int j = 0; j < BW_RU; ++j) acc[i][j] = 0.f;
#pragma unroll 2
            for (int q = 2 * warp + ksub; q < nquads; q += 2 * BW_NW) {
                float4 a[BW_RB], w[BW_RU];
#pragma unroll
                for (int i = 0; i < BW_RB; ++i) a[i] = *reinterpret_cast<const float4 *>(dgl + (size_t)(4 * i) * GHP + 4 * q);
#pragma unroll
                for (int j = 0; j < BW_RU; ++j) w[j] = *reinterpret_cast<const float4 *>(wl + (size_t)(4 * j) * GHP + 4 * q);
#pragma unroll
                for (int i = 0; i < BW_RB; ++i)
#pragma unroll
                    for (int j = 0; j < BW_RU; ++j) {
                        float v = acc[i][j];
                        v = fmaf(a[i].x, w[j].x, v);
                        v = fmaf(a[i].y, w[j].y, v);
                        v = fmaf(a[i].z, w[j].z, v);
                        v = fmaf(a[i].w, w[j].w, v);
                        acc[i][j] = v;
                    }
            }
            // the two half-warps meet, then each half writes half of the tile's partial sums
            float *pw = part + (size_t)warp * BW_CELLS;
#pragma unroll
            for (int i = 0; i < BW_RB; ++i)
#pragma unroll
                for (int j = 0; j < BW_RU; ++j) {
                    const float v = acc[i][j] + __shfl_xor_sync(0xffffffffu, acc[i][j], 16);
                    if (((i * BW_RU + j) & 1) == ksub) pw[(rb + 4 * i) * BW_HS + ub + 4 * j] = v;
                }
            bw_stamp(p, s, 3);
            __syncthreads();
            bw_stamp(p, s, 4);
            if (owner) {
                float v = 0.f;
#pragma unroll
                for (int w = 0; w < BW_NW; ++w) v += part[w * BW_CELLS + tid];
                dh += v;
            }
        }

        if (oval) {
            float *ox = p.dgx + row * GH + ou;
            if constexpr (CELL == DL4SS_CELL_LSTM) {
                const float ig = gv[0], fg = gv[1], gg = gv[2], og = gv[3];
                const float tc = tanhf(cv);
                const float dc = fmaf(dh * og, 1.0f - tc * tc, carry);
                const float dai = dc * gg * ig * (1.0f - ig);
                const float daf = dc * pv * fg * (1.0f - fg);
                const float dag = dc * ig * (1.0f - gg * gg);
                const float dao = dh * tc * og * (1.0f - og);
                carry = dc * fg;
                __stcg(ox, dai); __stcg(ox + H, daf); __stcg(ox + 2 * (size_t)H, dag); __stcg(ox + 3 * (size_t)H, dao);
            } else {
                dh += carry;                                   // the z * h_{t-1} path
                const float rg = gv[0], zg = gv[1], ng = gv[2];
                const float hn = cv;                           // W_hn h + b_hn saved by the forward kernel
                const float dn = dh * (1.0f - zg);
                const float dan = dn * (1.0f - ng * ng);
                const float dar = dan * hn * rg * (1.0f - rg);
                const float daz = dh * (pv - ng) * zg * (1.0f - zg);
                const float dhn = dan * rg;
                carry = dh * zg;
                float *oh = p.dgh + row * GH + ou;
                __stcg(ox, dar); __stcg(ox + H, daz); __stcg(ox + 2 * (size_t)H, dan);
                __stcg(oh, dar); __stcg(oh + H, daz); __stcg(oh + 2 * (size_t)H, dhn);
            }
        }
        if (s + 1 < T) {
            bw_stamp(p, s, 5);
            __syncthreads();
            bw_stamp(p, s, 6);
            if (tid == 0) {
                // release = MEMBAR.GPU + RED, cumulative over the barrier above: a separate __threadfence() paid a second membar
                asm volatile("red.release.gpu.global.add.u32 [%0], %1;\n" ::"l"(counter), "r"(1u) : "memory");
            }
            bw_stamp(p, s, 7);
        }
    }
}


// ------------------------------------------------------------------------------------------------------------------
// Tensor-core form of the same chain (bf16x3: operands split into bf16 hi/lo planes, three products per fp32 product,
// fp32 accumulation).  The per-step product is [16 utterances x G*H] x [G*H x 20 units]: K = 1200 against a 16 x 20
// output.  tcgen05 is the wrong tool for this shape -- a UMMA has K = 16 and costs ~75 cycles however small it is
// (DESIGN.md 4), so 75 dependent k steps would take ~5.6 k cycles on the single issuing thread -- whereas warp-level
// mma.sync.m16n8k16 spreads the k steps over 16 warps (5 k steps x 9 MMAs each) and leaves the accumulators in
// registers: the product drops from ~6.7 k cycles on the fp32 pipe to a few hundred.  What the kernel changes around it:
//   * W_hh^T slice resident in shared memory as bf16 hi/lo planes [2][24][pitch] (B fragments are plain 32-bit loads);
//   * the gate gradients are exchanged as bf16 hi/lo planes (`xplanes` [2][B*T][2][GHg], the same 4 bytes per value
//     as fp32), written by the owner threads next to the fp32 dgx / dgh the weight GEMMs read, so that A fragments
//     need no conversion; a tile's 16 rows x 2 planes arrive as 32 bulk copies (cp.async.bulk) completing on an
//     mbarrier, two issued by lane 0 of every warp once thread 0 has seen the group counter (the lanes of one warp would
//     issue them one after the other through the uniform datapath): no per-thread cp.async, no wait_group + CTA barrier;
//   * row pitch == 4 (mod 32) 32-bit words: the 8 rows x 4 k-pairs of a fragment load hit 32 different banks.
//   * up to BM_MAXT utterance tiles per CTA (round 2): 64 utterances fill the 120 co-resident CTAs with one tile each; a larger
//     batch used to run in chunks of 64, one launch after the other, each a pure latency chain.  Now a CTA walks its tiles
//     in turn inside every step, with the same resident W_hh^T slice and the same tile buffer: while tile i's gradients
//     travel to its siblings (release -> counter visible -> bulk load: ~4 k of the 7.5 k cycles of a step) the CTA works
//     on tile i+1.  The six warps that own no cells look at the NEXT tile's counter during the product and, when its
//     siblings have published (normal from 2 tiles per CTA on), request its rows right after the product, under this
//     tile's gate arithmetic; the release (MEMBAR.GPU + RED, ~1.3 k cycles) is issued by one of those warps, so it stalls
//     no owner.  256 utterances: 3.30 ms per layer in one launch against 4 x 1.31 ms (5.2 k instead of 7.45 k cycles per
//     tile and step).
constexpr int BM_NW = 16;
constexpr int BM_THREADS = 32 * BM_NW;
constexpr int BM_NP = 24;                 // units padded to 3 n-tiles of 8
constexpr int BM_FREE0 = 10;              // warps [10, 16) hold no cell owners (BW_CELLS = 320 = 10 warps)
constexpr int BM_MAXT = 4;                // utterance tiles per CTA: while one tile's gate gradients travel to its siblings
                                          // (release -> counter visible -> bulk load: ~4 k of the 7.5 k cycles of a step) the CTA
                                          // works on the next tile with the same resident W_hh^T slice and the same tile buffer

struct RnnBwdTcParams {
    const float *dy, *whh, *gates, *cells, *y;
    float *dgx, *dgh;
    __nv_bfloat16 *xplanes; // [2][B*T][2][GHg] recurrent-side gate gradients as bf16 hi/lo planes (pad columns zero)
    unsigned *counters;
    int B, T, H, GHg, pitch, nks;
    int b_begin, batch_tiles, nslices;
    int tpc, ngroups;       // utterance tiles a CTA walks per step (<= BM_MAXT), groups of tpc tiles in this launch
    long long *trace;
    int trace_steps;
};

__device__ __forceinline__ void bm_stamp(const RnnBwdTcParams &p, int s, int slot) {
    if (p.trace != nullptr && blockIdx.x == 0 && threadIdx.x == 0 && s < p.trace_steps) p.trace[s * 8 + slot] = clock64();
}
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void bulk_g2s(void *smem, const void *gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                 ::"r"(smem_u32(smem)), "l"(gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

template <int CELL>
__global__ void __launch_bounds__(BM_THREADS, 1)
rnn_bwd_tc_kernel(const RnnBwdTcParams p) {
    constexpr int G = (CELL == DL4SS_CELL_LSTM) ? 4 : 3;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int H = p.H, T = p.T, pitch = p.pitch, GHg = p.GHg;
    const int GH = G * H;
    __nv_bfloat16 *Wt = reinterpret_cast<__nv_bfloat16 *>(smem_raw);          // [2][BM_NP][pitch]
    __nv_bfloat16 *dgA = Wt + (size_t)2 * BM_NP * pitch;                       // [2][BW_BT][pitch]
    float *part = reinterpret_cast<float *>(dgA + (size_t)2 * BW_BT * pitch);  // [BM_NW][BW_BT * BM_NP]
    uint64_t *bar = reinterpret_cast<uint64_t *>(part + BM_NW * BW_BT * BM_NP);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    int bid = blockIdx.x;
    const int slice = bid % p.nslices; bid /= p.nslices;
    const int grp = bid % p.ngroups;
    const int dir = bid / p.ngroups;
    const int bt0 = grp * p.tpc;
    const int nt = min(p.tpc, p.batch_tiles - bt0);          // tiles this CTA walks
    const int u0 = slice * BW_HS;
    const size_t plane_elems = (size_t)p.B * T * 2 * GHg;

    {   // zero both operand arrays (pad rows / columns / absent utterances stay zero), then the resident W planes
        uint4 *z = reinterpret_cast<uint4 *>(smem_raw);
        const int n16 = (2 * (BM_NP + BW_BT) * pitch * 2) / 16;
        for (int i = tid; i < n16; i += BM_THREADS) z[i] = make_uint4(0u, 0u, 0u, 0u);
        __syncthreads();
        const float *wsrc = p.whh + (size_t)dir * GH * H + u0;
        for (int i = tid; i < GH * BW_HS; i += BM_THREADS) {
            const int k = i / BW_HS, j = i - k * BW_HS;
            const float v = __ldg(wsrc + (size_t)k * H + j);
            const __nv_bfloat16 hi = __float2bfloat16_rn(v);
            Wt[(size_t)j * pitch + k] = hi;
            Wt[(size_t)(BM_NP + j) * pitch + k] = __float2bfloat16_rn(v - __bfloat162float(hi));
        }
        if (tid == 0) {
            mbar_init(&bar[0], 1);
            mbar_init(&bar[1], 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic zero fill before the bulk copies' writes
    }
    __syncthreads();

    // fragment coordinates of mma.m16n8k16: lane = 4*g + t
    const int fg = lane >> 2, ft = lane & 3;
    const uint32_t *A32 = reinterpret_cast<const uint32_t *>(dgA);
    const uint32_t *W32 = reinterpret_cast<const uint32_t *>(Wt);
    const int pw = pitch >> 1;                                   // row pitch in 32-bit words
    // owner mapping: thread -> cell (r, j) of every tile
    const bool owner = tid < BW_CELLS;
    const int orow = tid / BW_HS, oj = tid - orow * BW_HS;
    const int ou = u0 + oj;
    float carry[BM_MAXT];
#pragma unroll
    for (int i = 0; i < BM_MAXT; ++i) carry[i] = 0.f;
    uint32_t uses0 = 0, uses1 = 0;                               // completed uses of the two mbarriers (phase parity)
    bool early = false;                                          // this tile's rows were requested during the previous tile
    volatile int *early_flag = reinterpret_cast<volatile int *>(bar + 2);

    for (int s = 0; s < T; ++s) {
        const int t = dir ? s : (T - 1 - s);
        const int tp = dir ? t + 1 : t - 1;
        const int tl = dir ? t - 1 : t + 1;
        const bool has_prev = (tp >= 0 && tp < T);

#pragma unroll
        for (int ti = 0; ti < BM_MAXT; ++ti) {
            if (ti < nt) {
                const int row0 = p.b_begin + (bt0 + ti) * BW_BT;
                const int nrows = min(BW_BT, p.B - row0);
                unsigned *counter = p.counters + (size_t)(dir * p.batch_tiles + bt0 + ti) * BW_CTR_STRIDE;
                const int ob = row0 + orow;
                const bool oval = owner && ob < p.B;
                const bool tr0 = (ti == 0);

                float dyv = 0.f, gv[G], cv = 0.f, pv = 0.f;
#pragma unroll
                for (int g = 0; g < G; ++g) gv[g] = 0.f;
                const size_t row = ((size_t)ob * T + t) * 2 + dir;
                if (oval) {
                    dyv = __ldg(p.dy + ((size_t)ob * T + t) * 2 * H + (size_t)dir * H + ou);
#pragma unroll
                    for (int g = 0; g < G; ++g) gv[g] = __ldg(p.gates + row * GH + (size_t)g * H + ou);
                    cv = __ldg(p.cells + row * H + ou);
                    if (has_prev) {
                        if (CELL == DL4SS_CELL_LSTM) pv = __ldg(p.cells + (((size_t)ob * T + tp) * 2 + dir) * H + ou);
                        else pv = __ldg(p.y + ((size_t)ob * T + tp) * 2 * H + (size_t)dir * H + ou);
                    }
                }

                // the (tile, step) that follows this one in the CTA's walk, and whether it reads exchanged rows
                const bool last_ti = (ti + 1 >= nt);
                const int s_n = last_ti ? s + 1 : s;
                const int ti_n = last_ti ? 0 : ti + 1;
                const bool have_next = (s_n >= 1 && s_n < T);
                const int row0_n = p.b_begin + (bt0 + ti_n) * BW_BT;
                const int nrows_n = min(BW_BT, p.B - row0_n);
                const int t_n = dir ? s_n : (T - 1 - s_n);
                const int tl_n = dir ? t_n - 1 : t_n + 1;
                const unsigned *counter_n = p.counters + (size_t)(dir * p.batch_tiles + bt0 + ti_n) * BW_CTR_STRIDE;

                float dh = dyv;
                unsigned seen_n = 0;
                if (s > 0) {
                    if (!early) {
                        // thread 0 watches the tile's group counter; the 32 row copies are then issued by lane 0 of every warp (two
                        // each): 32 copies from the lanes of ONE warp leave through the uniform datapath one after the other (~1 k cycles)
                        if (tid == 0) {
                            if (tr0) bm_stamp(p, s, 0);
                            const unsigned want = (unsigned)p.nslices * (unsigned)s;
                            while (bw_ld_acquire(counter) < want) { __nanosleep(20); }
                            mbar_expect_tx(&bar[0], (uint32_t)(nrows * 2 * GHg * 2));
                            mbar_arrive(&bar[1]);
                            if (tr0) bm_stamp(p, s, 1);
                        }
                        mbar_wait(&bar[1], uses1 & 1u);
                        ++uses1;
                        if (lane == 0) {
                            asm volatile("fence.proxy.async.global;\n" ::: "memory");   // the siblings' generic stores -> bulk-copy reads
#pragma unroll
                            for (int i = 0; i < 32 / BM_NW; ++i) {
                                const int idx = warp * (32 / BM_NW) + i;
                                const int r = idx >> 1, pl = idx & 1;
                                if (r < nrows)
                                    bulk_g2s(dgA + (size_t)(pl * BW_BT + r) * pitch,
                                             p.xplanes + (size_t)pl * plane_elems + (((size_t)(row0 + r) * T + tl) * 2 + dir) * GHg,
                                             (uint32_t)(GHg * 2), &bar[0]);
                            }
                        }
                    }
                    mbar_wait(&bar[0], uses0 & 1u);
                    ++uses0;
                    if (tr0) bm_stamp(p, s, 2);
                    // one look at the NEXT tile's counter, in flight during the product (warp 15 holds no cell owners)
                    if (warp == BM_NW - 1 && lane == 0 && have_next) seen_n = bw_ld_acquire(counter_n);

                    float acc[3][4];
#pragma unroll
                    for (int n = 0; n < 3; ++n)
#pragma unroll
                        for (int i = 0; i < 4; ++i) acc[n][i] = 0.f;
                    for (int ks = warp; ks < p.nks; ks += BM_NW) {
                        const int kw = ks * 8 + ft;                          // word index of k = 16*ks + 2*ft
                        uint32_t ah[4], al[4];
                        const uint32_t *a0 = A32 + (size_t)fg * pw + kw;
                        const uint32_t *a1 = a0 + (size_t)BW_BT * pw;        // lo plane
                        ah[0] = a0[0]; ah[1] = a0[8 * pw]; ah[2] = a0[4]; ah[3] = a0[8 * pw + 4];
                        al[0] = a1[0]; al[1] = a1[8 * pw]; al[2] = a1[4]; al[3] = a1[8 * pw + 4];
#pragma unroll
                        for (int n = 0; n < 3; ++n) {
                            const uint32_t *b0 = W32 + (size_t)(8 * n + fg) * pw + kw;
                            const uint32_t *b1 = b0 + (size_t)BM_NP * pw;    // lo plane
                            const uint32_t bh0 = b0[0], bh1 = b0[4], bl0 = b1[0], bl1 = b1[4];
                            mma_bf16_16816(acc[n], ah, bl0, bl1);
                            mma_bf16_16816(acc[n], al, bh0, bh1);
                            mma_bf16_16816(acc[n], ah, bh0, bh1);
                        }
                    }
                    float *pwarp = part + (size_t)warp * (BW_BT * BM_NP);
#pragma unroll
                    for (int n = 0; n < 3; ++n) {
                        *reinterpret_cast<float2 *>(pwarp + fg * BM_NP + 8 * n + 2 * ft) = make_float2(acc[n][0], acc[n][1]);
                        *reinterpret_cast<float2 *>(pwarp + (fg + 8) * BM_NP + 8 * n + 2 * ft) = make_float2(acc[n][2], acc[n][3]);
                    }
                    if (tr0) bm_stamp(p, s, 3);
                    __syncthreads();
                    if (tr0) bm_stamp(p, s, 4);
                    // the tile buffer is free: if the next tile's siblings have published, request its rows now, under the gate
                    // arithmetic of this tile (with >= 2 tiles per CTA they normally have: their release is a whole tile old)
                    // (the six warps without cell owners share the 32 copies: one warp's lanes would issue them one after the other)
                    if (warp >= BM_FREE0) {
                        if (warp == BM_NW - 1 && lane == 0) {
                            const bool rdy = have_next && seen_n >= (unsigned)p.nslices * (unsigned)s_n;
                            if (rdy) mbar_expect_tx(&bar[0], (uint32_t)(nrows_n * 2 * GHg * 2));
                            *early_flag = rdy ? 1 : 0;
                        }
                        asm volatile("bar.sync 1, %0;\n" ::"n"(32 * (BM_NW - BM_FREE0)) : "memory");
                        if (*early_flag != 0 && lane < 6) {
                            const int idx = (warp - BM_FREE0) * 6 + lane;            // 36 slots for 32 copies
                            const int r = idx >> 1, pl = idx & 1;
                            if (idx < 32 && r < nrows_n) {
                                asm volatile("fence.proxy.async.global;\n" ::: "memory");
                                bulk_g2s(dgA + (size_t)(pl * BW_BT + r) * pitch,
                                         p.xplanes + (size_t)pl * plane_elems + (((size_t)(row0_n + r) * T + tl_n) * 2 + dir) * GHg,
                                         (uint32_t)(GHg * 2), &bar[0]);
                            }
                        }
                    }
                    if (owner) {
                        float v = 0.f;
#pragma unroll
                        for (int w = 0; w < BM_NW; ++w) v += part[w * (BW_BT * BM_NP) + orow * BM_NP + oj];
                        dh += v;
                    }
                }

                // gate arithmetic; the bf16 planes the siblings wait for are published first, the fp32 arrays the weight GEMMs
                // read follow after the release, off the chain
                float rec[G], dxn = 0.f;
#pragma unroll
                for (int g = 0; g < G; ++g) rec[g] = 0.f;
                if (oval) {
                    if constexpr (CELL == DL4SS_CELL_LSTM) {
                        const float ig = gv[0], fgt = gv[1], gg = gv[2], og = gv[3];
                        const float tc = tanhf(cv);
                        const float dc = fmaf(dh * og, 1.0f - tc * tc, carry[ti]);
                        rec[0] = dc * gg * ig * (1.0f - ig);
                        rec[1] = dc * pv * fgt * (1.0f - fgt);
                        rec[2] = dc * ig * (1.0f - gg * gg);
                        rec[3] = dh * tc * og * (1.0f - og);
                        carry[ti] = dc * fgt;
                    } else {
                        dh += carry[ti];
                        const float rg = gv[0], zg = gv[1], ng = gv[2];
                        const float hn = cv;
                        const float dn = dh * (1.0f - zg);
                        dxn = dn * (1.0f - ng * ng);
                        rec[0] = dxn * hn * rg * (1.0f - rg);
                        rec[1] = dh * (pv - ng) * zg * (1.0f - zg);
                        rec[2] = dxn * rg;
                        carry[ti] = dh * zg;
                    }
                    __nv_bfloat16 *xh = p.xplanes + row * GHg + ou;
#pragma unroll
                    for (int g = 0; g < G; ++g) {
                        const __nv_bfloat16 hi = __float2bfloat16_rn(rec[g]);
                        xh[(size_t)g * H] = hi;
                        xh[plane_elems + (size_t)g * H] = __float2bfloat16_rn(rec[g] - __bfloat162float(hi));
                    }
                }
                if (s == 0 && tid == 0) *early_flag = 0;
                if (tr0) bm_stamp(p, s, 5);
                // every thread is past the reduction (partial sums) and has published its cells
                __syncthreads();
                early = (*early_flag != 0);
                if (tr0) bm_stamp(p, s, 6);
                if (s + 1 < T && tid == 32 * (BM_NW - 2)) {
                    // release = MEMBAR.GPU + RED, cumulative over the barrier above (a separate __threadfence() paid a second membar);
                    // issued from warp 14, which owns no cells: the ~1.3 k cycles of the membar stall nobody's stores
                    asm volatile("red.release.gpu.global.add.u32 [%0], %1;\n" ::"l"(counter), "r"(1u) : "memory");
                }
                if (tr0) bm_stamp(p, s, 7);
                if (oval && p.dgx != nullptr) {
                    float *ox = p.dgx + row * GH + ou;
                    if constexpr (CELL == DL4SS_CELL_LSTM) {
#pragma unroll
                        for (int g = 0; g < G; ++g) __stcs(ox + (size_t)g * H, rec[g]);
                    } else {
                        float *oh = p.dgh + row * GH + ou;
                        __stcs(ox, rec[0]); __stcs(ox + H, rec[1]); __stcs(ox + 2 * (size_t)H, dxn);
#pragma unroll
                        for (int g = 0; g < G; ++g) __stcs(oh + (size_t)g * H, rec[g]);
                    }
                }
            }
        }
    }
}

static long long *g_bwd_trace = nullptr;
static int g_bwd_trace_steps = 0;
static int g_bwd_tpc = [] { const char *e = getenv("DL4SS_BWD_TILES_PER_CTA"); return e ? atoi(e) : 0; }();   // A/B: force >= n tiles per CTA

static int bwd_ghp(int GH) { return GH + ((8 - GH % 32 + 32) % 32); }   // row pitch == 8 (mod 32) floats: the 8 distinct
                                                                         // 16-byte chunks of a warp load tile the 32 banks

template <int CELL>
static int launch_rnn_bwd(RnnBwdParams p, int rows_left, cudaStream_t st, int *launched_rows) {
    constexpr int G = (CELL == DL4SS_CELL_LSTM) ? 4 : 3;
    p.GHP = bwd_ghp(G * p.H);
    p.nslices = p.H / BW_HS;
    const size_t smem = ((size_t)(BW_HS + BW_BT) * p.GHP + (size_t)BW_NW * BW_CELLS) * sizeof(float);
    if (smem > 227 * 1024) {
        set_error("rnn_layer_bwd: H=%d needs %zu B of shared memory per CTA", p.H, smem);
        return DL4SS_EUNSUPPORTED;
    }
    auto kern = rnn_bwd_kernel<CELL>;
    DL4SS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    DL4SS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, BW_THREADS, smem));
    const int max_tiles = per_sm * sm_count() / (2 * p.nslices);
    if (max_tiles < 1) {
        set_error("rnn_layer_bwd: %d co-resident CTAs cannot hold one tile (%d slices x 2 directions)",
                  per_sm * sm_count(), p.nslices);
        return DL4SS_EUNSUPPORTED;
    }
    int tiles = cdiv(rows_left, BW_BT);
    if (tiles > max_tiles) tiles = max_tiles;
    p.batch_tiles = tiles;
    *launched_rows = tiles * BW_BT;
    void *args[] = {(void *)&p};
    DL4SS_CUDA(cudaLaunchCooperativeKernel((void *)kern, dim3(2 * tiles * p.nslices), dim3(BW_THREADS), args, smem, st));
    count_launch();
    return DL4SS_OK;
}

static int bwd_tc_ghg(int GH) { return (GH + 7) / 8 * 8; }                // 16-byte rows for the bulk copies
static int bwd_tc_pitch(int GH) {                                          // bf16 elements; == 4 (mod 32) words
    int w = cdiv(GH, 16) * 8;
    w += (4 - w % 32 + 32) % 32;
    return 2 * w;
}
static size_t bwd_tc_smem(int GH) {
    return (size_t)2 * (BM_NP + BW_BT) * bwd_tc_pitch(GH) * sizeof(__nv_bfloat16) +
           (size_t)BM_NW * BW_BT * BM_NP * sizeof(float) + 32;
}

template <int CELL>
static int launch_rnn_bwd_tc(RnnBwdTcParams p, int rows_left, cudaStream_t st, int *launched_rows) {
    constexpr int G = (CELL == DL4SS_CELL_LSTM) ? 4 : 3;
    const int GH = G * p.H;
    p.GHg = bwd_tc_ghg(GH);
    p.pitch = bwd_tc_pitch(GH);
    p.nks = cdiv(GH, 16);
    p.nslices = p.H / BW_HS;
    const size_t smem = bwd_tc_smem(GH);
    auto kern = rnn_bwd_tc_kernel<CELL>;
    DL4SS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    DL4SS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, BM_THREADS, smem));
    const int max_tiles = per_sm * sm_count() / (2 * p.nslices);
    if (max_tiles < 1) {
        set_error("rnn_layer_bwd_tc: %d co-resident CTAs cannot hold one tile (%d slices x 2 directions)",
                  per_sm * sm_count(), p.nslices);
        return DL4SS_EUNSUPPORTED;
    }
    // as few tiles per CTA as the co-resident CTAs allow; up to BM_MAXT * max_tiles tiles per launch
    int tiles = cdiv(rows_left, BW_BT);
    if (tiles > max_tiles * BM_MAXT) tiles = max_tiles * BM_MAXT;
    int tpc = cdiv(tiles, max_tiles);
    if (g_bwd_tpc > tpc) tpc = g_bwd_tpc < BM_MAXT ? g_bwd_tpc : BM_MAXT;
    p.batch_tiles = tiles;
    p.tpc = tpc;
    p.ngroups = cdiv(tiles, tpc);
    *launched_rows = tiles * BW_BT;
    void *args[] = {(void *)&p};
    DL4SS_CUDA(cudaLaunchCooperativeKernel((void *)kern, dim3(2 * p.ngroups * p.nslices), dim3(BM_THREADS), args, smem, st));
    count_launch();
    return DL4SS_OK;
}

static bool rnn_bwd_tc_supported(int H, int cell) {
    if (!(cell == DL4SS_CELL_LSTM || cell == DL4SS_CELL_GRU)) return false;
    if (H < BW_HS || H % BW_HS != 0) return false;
    const int G = (cell == DL4SS_CELL_LSTM) ? 4 : 3;
    return bwd_tc_smem(G * H) <= 227 * 1024;
}

static bool rnn_bwd_supported(int H, int cell) {
    if (!(cell == DL4SS_CELL_LSTM || cell == DL4SS_CELL_GRU)) return false;
    if (H < BW_HS || H % BW_HS != 0) return false;
    const int G = (cell == DL4SS_CELL_LSTM) ? 4 : 3;
    const size_t smem = ((size_t)(BW_HS + BW_BT) * bwd_ghp(G * H) + (size_t)BW_NW * BW_CELLS) * sizeof(float);
    return smem <= 227 * 1024;
}

}  // namespace dl4ss

using namespace dl4ss;

// profiling hook: device buffer of steps*8 int64 receiving CTA 0's per-phase clock64() stamps (null = off)
extern "C" void dl4ss_rnn_bwd_set_trace(void *dev_buf, int steps) {
    g_bwd_trace = (long long *)dev_buf;
    g_bwd_trace_steps = dev_buf ? steps : 0;
}

extern "C" int dl4ss_rnn_bwd_supported(int H, int cell) { return rnn_bwd_supported(H, cell) ? 1 : 0; }

extern "C" size_t dl4ss_rnn_bwd_workspace_bytes(int B, int T, int H, int cell) {
    (void)T; (void)H; (void)cell;
    if (B <= 0) return 256;
    return (size_t)2 * cdiv(B, BW_BT) * BW_CTR_STRIDE * sizeof(unsigned);
}

extern "C" int dl4ss_rnn_layer_bwd(int cell, const float *dy, const float *whh, const float *gates_save,
                                   const float *cell_save, const float *y, float *dgx, float *dgh, int B, int T, int H,
                                   void *workspace, size_t workspace_bytes, void *stream) {
    DL4SS_CHECK_ARG(cell == DL4SS_CELL_LSTM || cell == DL4SS_CELL_GRU, "rnn_layer_bwd: bad cell %d", cell);
    DL4SS_CHECK_ARG(dy && whh && gates_save && cell_save && dgx, "rnn_layer_bwd: null operand");
    DL4SS_CHECK_ARG(cell == DL4SS_CELL_LSTM || (y && dgh), "rnn_layer_bwd: GRU needs y and dgh");
    DL4SS_CHECK_ARG(B >= 0 && T >= 1 && H >= 1, "rnn_layer_bwd: bad B/T/H %d/%d/%d", B, T, H);
    if (!rnn_bwd_supported(H, cell)) {
        set_error("rnn_layer_bwd: H=%d unsupported (needs a multiple of %d whose W_hh slice fits shared memory); "
                  "use dl4ss_rnn_bwd_step", H, BW_HS);
        return DL4SS_EUNSUPPORTED;
    }
    if (B == 0) return DL4SS_OK;
    const size_t need = dl4ss_rnn_bwd_workspace_bytes(B, T, H, cell);
    if (!workspace || workspace_bytes < need) {
        set_error("rnn_layer_bwd: workspace %zu B < %zu B", workspace_bytes, need);
        return DL4SS_EWORKSPACE;
    }
    DL4SS_CHECK_ARG((((uintptr_t)workspace) & 255) == 0, "rnn_layer_bwd: workspace must be 256-byte aligned");
    DL4SS_CHECK_ARG((((uintptr_t)dgx) & 15) == 0 && (!dgh || (((uintptr_t)dgh) & 15) == 0),
                    "rnn_layer_bwd: dgx / dgh must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    DL4SS_CUDA(cudaMemsetAsync(workspace, 0, need, st));
    RnnBwdParams p;
    p.dy = dy; p.whh = whh; p.gates = gates_save; p.cells = cell_save; p.y = y; p.dgx = dgx; p.dgh = dgh;
    p.B = B; p.T = T; p.H = H; p.GHP = 0; p.batch_tiles = 0; p.nslices = 0;
    p.trace = g_bwd_trace; p.trace_steps = g_bwd_trace_steps;
    unsigned *ctr = (unsigned *)workspace;
    int b0 = 0;
    while (b0 < B) {
        p.b_begin = b0;
        p.counters = ctr;
        int done = 0;
        int rc = (cell == DL4SS_CELL_LSTM) ? launch_rnn_bwd<DL4SS_CELL_LSTM>(p, B - b0, st, &done)
                                           : launch_rnn_bwd<DL4SS_CELL_GRU>(p, B - b0, st, &done);
        if (rc) return rc;
        ctr += (size_t)2 * (done / BW_BT) * BW_CTR_STRIDE;
        b0 += done;
    }
    return DL4SS_OK;
}

extern "C" int dl4ss_rnn_bwd_tc_supported(int H, int cell) { return rnn_bwd_tc_supported(H, cell) ? 1 : 0; }

extern "C" size_t dl4ss_rnn_bwd_tc_xplanes_bytes(int B, int T, int H, int cell) {
    if (B <= 0 || T <= 0 || H <= 0) return 0;
    const int G = (cell == DL4SS_CELL_LSTM) ? 4 : 3;
    return (size_t)2 * B * T * 2 * bwd_tc_ghg(G * H) * sizeof(__nv_bfloat16);
}

extern "C" int dl4ss_rnn_layer_bwd_tc(int cell, const float *dy, const float *whh, const float *gates_save,
                                      const float *cell_save, const float *y, float *dgx, float *dgh, void *xplanes,
                                      int B, int T, int H, void *workspace, size_t workspace_bytes, void *stream) {
    DL4SS_CHECK_ARG(cell == DL4SS_CELL_LSTM || cell == DL4SS_CELL_GRU, "rnn_layer_bwd_tc: bad cell %d", cell);
    // dgx (the fp32 copy of the gate gradients) may be NULL for the LSTM: the weight / input / bias gradients can all be
    // taken from the bf16 planes (xplanes), and the chain then issues half as many scattered stores per step
    DL4SS_CHECK_ARG(dy && whh && gates_save && cell_save && xplanes && (dgx || cell == DL4SS_CELL_LSTM), "rnn_layer_bwd_tc: null operand");
    DL4SS_CHECK_ARG(cell == DL4SS_CELL_LSTM || (y && dgh), "rnn_layer_bwd_tc: GRU needs y and dgh");
    DL4SS_CHECK_ARG(B >= 0 && T >= 1 && H >= 1, "rnn_layer_bwd_tc: bad B/T/H %d/%d/%d", B, T, H);
    if (!rnn_bwd_tc_supported(H, cell)) {
        set_error("rnn_layer_bwd_tc: H=%d unsupported (needs a multiple of %d whose W_hh slice fits shared memory); "
                  "use dl4ss_rnn_bwd_step", H, BW_HS);
        return DL4SS_EUNSUPPORTED;
    }
    if (B == 0) return DL4SS_OK;
    const size_t need = dl4ss_rnn_bwd_workspace_bytes(B, T, H, cell);
    if (!workspace || workspace_bytes < need) {
        set_error("rnn_layer_bwd_tc: workspace %zu B < %zu B", workspace_bytes, need);
        return DL4SS_EWORKSPACE;
    }
    DL4SS_CHECK_ARG((((uintptr_t)workspace) & 255) == 0, "rnn_layer_bwd_tc: workspace must be 256-byte aligned");
    DL4SS_CHECK_ARG((((uintptr_t)xplanes) & 15) == 0, "rnn_layer_bwd_tc: xplanes must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    DL4SS_CUDA(cudaMemsetAsync(workspace, 0, need, st));
    RnnBwdTcParams p;
    p.dy = dy; p.whh = whh; p.gates = gates_save; p.cells = cell_save; p.y = y; p.dgx = dgx; p.dgh = dgh;
    p.xplanes = (__nv_bfloat16 *)xplanes;
    p.B = B; p.T = T; p.H = H; p.GHg = p.pitch = p.nks = 0; p.batch_tiles = 0; p.nslices = 0; p.tpc = 1; p.ngroups = 0;
    p.trace = g_bwd_trace; p.trace_steps = g_bwd_trace_steps;
    unsigned *ctr = (unsigned *)workspace;
    int b0 = 0;
    while (b0 < B) {
        p.b_begin = b0;
        p.counters = ctr;
        int done = 0;
        int rc = (cell == DL4SS_CELL_LSTM) ? launch_rnn_bwd_tc<DL4SS_CELL_LSTM>(p, B - b0, st, &done)
                                           : launch_rnn_bwd_tc<DL4SS_CELL_GRU>(p, B - b0, st, &done);
        if (rc) return rc;
        ctr += (size_t)2 * (done / BW_BT) * BW_CTR_STRIDE;
        b0 += done;
    }
    return DL4SS_OK;
}
