// K3 (tensor-core form)  dl4ss_rnn_layer_tc_fwd : one bidirectional LSTM / GRU layer as ONE persistent
// kernel whose recurrent product W_hh * h_{t-1} runs on tcgen05 (bf16x3 split, fp32 TMEM accumulators).
//
// Replaces the T sequential cuDNN steps of nn.LSTM / nn.GRU (TDAA_beta/main_run_sstune_EvalVer.py:282-293,
// ...cRM_EvalVer.py:345-356).  H = 300 in every reference config; the step is a latency chain
// (184 MFLOP per step and direction at B=256), so the design shortens the chain and overlaps two of them:
//   * a CTA owns (direction, slice of 20 hidden units, up to TWO tiles of 32 utterances).  The slice's
//     80 rows of W_hh (4 gate rows per unit, unit-major; hi and lo bf16 planes) are written ONCE into
//     TENSOR MEMORY (tcgen05.st, 320 of the 512 columns) and stay there for all T steps as the M=128
//     A operand of the UMMA: an smem A operand costs ~64 cycles of operand fetch per instruction, which
//     dominates UMMAs this small (measured 82 cycles each), a TMEM A operand does not; the 32-utterance
//     h tile is the N operand, so the small dimension sits where the tensor core does not mind;
//   * per step and tile the 15 slice-CTAs of a (direction, tile) group exchange h through an L2-resident
//     bf16 hi/lo ping-pong buffer: epilogue warps store their piece, fence, relaxed red.add on the
//     group counter; a loader thread spins (ld.acquire), fence.proxy.async, TMA-loads the 32 x H tile
//     (5 k-chunks, one mbarrier each); the MMA thread issues per 16-wide k step
//         W_hi x [h_hi ; h_lo] (N=64, the planes are adjacent in smem)  and  W_lo x h_hi (N=32)
//     into three accumulator column blocks the epilogue adds;
//   * 6 epilogue warps per tile: a thread owns accumulator row (unit u, gate g), pulls 16 batch columns
//     out of TMEM, a 4-lane shuffle transpose hands every thread the 4 gates of 4 (unit, utterance)
//     cells, it adds the hoisted input projection (cp.async-prefetched one step ahead by 2 producer
//     warps, the kernel's only HBM read), applies the gates, publishes h_t, then writes y;
//   * the two tiles of a CTA are independent chains served in alternation: one tile's exchange latency
//     hides behind the other tile's MMA + gate math.  Groups never wait on each other; the launch is
//     cooperative so every CTA is resident.
#include "tc_ptx.cuh"
#include <stdlib.h>

namespace dl4ss {

constexpr int RT_NT = 32;                      // utterances per tile = UMMA N
constexpr int RT_HS = 32;                      // hidden units per slice: 4 gate rows x 32 units fill the 128 TMEM lanes
constexpr int RT_ROWS = 4 * RT_HS;             // W rows per slice (GRU: 4th row of a unit is zero; units >= H: zero)
constexpr int RT_KC = 64;                      // k per chunk (128 B of bf16: one swizzle row)
constexpr int RT_MAXKC = 5;                    // H <= 320
constexpr int RT_TILES = 3;                    // tiles interleaved per CTA
constexpr int RT_TCOLS = 2 * RT_NT;            // TMEM accumulator columns per tile: hi*hi | hi*lo + lo*hi
constexpr int RT_WCOL = RT_TILES * RT_TCOLS;   // first TMEM column of the resident W (A operand)
constexpr int RT_WPLANE = RT_MAXKC * RT_KC / 2; // TMEM columns of one W plane (2 bf16 per column)
constexpr int RT_EPI_PER_TILE = 8;             // epilogue warps per tile: 4 sub-partitions x 2 column halves
constexpr int RT_EPI_WARPS = RT_TILES * RT_EPI_PER_TILE;     // warps [0, 16): epilogue (warp % 4 = TMEM sub-partition)
constexpr int RT_W_MGR = RT_EPI_WARPS;         // warps 24..26: tile managers (h tile load + UMMA issue); warp 24 owns the TMEM allocation
constexpr int RT_THREADS = 32 * (RT_EPI_WARPS + RT_TILES);
constexpr int RT_WBLK = RT_ROWS * 128;         // bytes of one (k-chunk, plane) W block
constexpr int RT_HBLK = RT_NT * 128;           // bytes of one (k-chunk, plane) h block
constexpr int RT_XTILE = RT_NT * 4 * RT_HS;    // floats of one xproj buffer: [utterance][gate][32 units], 128-byte rows, TMA 128B swizzle
constexpr int RT_CTR_STRIDE = 64;              // uints between group counters: one 256-byte line each (atomics and the
                                               // pollers of different groups must not share an L2 line)

struct RnnTcParams {
    const float *xproj;        // [B,T,2,G*H]
    const float *bhn;          // [2,H] (GRU) or null
    float *y;                  // [B,T,2H]
    float *gates_save;         // [B,T,2,G*H] or null
    float *cell_save;          // [B,T,2,H] or null
    __nv_bfloat16 *y_planes;   // optional [2 (hi,lo)][B*T][Kpy]: y pre-split for the next tensor-core projection
    float *hmean_out;          // optional [B,2H]: mean over T of y (ADDJUST input)
    int Kpy;
    __nv_bfloat16 *hbuf;       // [2 ping-pong][2 dir][2 plane][Bpad][Kp]
    unsigned *counters;        // [2][tiles_total]
    const __nv_bfloat16 *wplanes;   // packed W_hh planes [2][2 dir * 4H][Kp]
    int B, T, H, Kp, nkc, Bpad;
    int tile0, ntiles, tpg, ngroups, tiles_total, nslices;    // this launch: tiles [tile0, tile0+ntiles), tpg per CTA
    long long *trace;          // optional [steps][16] clock stamps of CTA 0 (profiling hook), else null
    int trace_steps;
    // yx: the group exchanges h through y_planes itself (row (b, t-1) of the layer output IS h_{t-1}) instead of a separate
    // ping-pong buffer: one set of publish stores per step instead of two.  A TMA box has to start on a 16-byte boundary, so
    // the reverse direction's window starts hshift = H % 8 columns early (column H - hshift) and its resident W_hh rows are
    // shifted by as many k positions (zeros in front); the <= 12 foreign columns either window touches meet zero weights.
    int yx, hshift;
};

// xproj is read exactly once: evict-first in L2, so that the 769 MB stream does not displace the layer's own output
// (y and its bf16 planes, which the next projection reads)
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;\n" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void cp_async16_cg(void *smem, const void *gmem, uint64_t pol) {
    asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;\n" ::"r"(smem_u32(smem)), "l"(gmem), "l"(pol) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void stamp(const RnnTcParams &p, int s, int slot) {
    if (p.trace != nullptr && blockIdx.x == 0 && s < p.trace_steps) p.trace[s * 16 + slot] = clock64();
}
__device__ __forceinline__ bool elect_one() {       // one lane of the (converged) warp
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.global;\n" ::: "memory"); }

// branch-free select (selp): a chain of ?: on a lane-dependent index compiles to divergent branches
__device__ __forceinline__ float selp_f(float a, float b, int pick_a) {
    float r;
    asm("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %3, 0;\n\tselp.f32 %0, %1, %2, p;\n\t}\n" : "=f"(r) : "f"(a), "f"(b), "r"(pick_a));
    return r;
}
// a[k] for k = 2*k1 + k0 given the lane's precomputed bits
__device__ __forceinline__ float sel4(int k0, int k1, float a0, float a1, float a2, float a3) {
    const float lo = selp_f(a1, a0, k0);
    const float hi = selp_f(a3, a2, k0);
    return selp_f(hi, lo, k1);
}

// 4 x 4 transpose between a thread's four values and the four lanes that differ in lane bits (LO, LO+1):
// afterwards x[r] is what lane (bits = r) held in x[my bits].  Two butterfly stages (partner 2 << LO, then 1 << LO),
// each 2 shuffles + 6 selects; b1 / b0 are this lane's two bits.
template <int LO>
__device__ __forceinline__ void transpose4(float (&x)[4], int b0, int b1) {
    {
        const float sa = selp_f(x[0], x[2], b1), sb = selp_f(x[1], x[3], b1);
        const float ra = __shfl_xor_sync(0xffffffffu, sa, 2 << LO), rb = __shfl_xor_sync(0xffffffffu, sb, 2 << LO);
        x[0] = selp_f(ra, x[0], b1); x[1] = selp_f(rb, x[1], b1);
        x[2] = selp_f(x[2], ra, b1); x[3] = selp_f(x[3], rb, b1);
    }
    {
        const float sa = selp_f(x[0], x[1], b0), sb = selp_f(x[2], x[3], b0);
        const float ra = __shfl_xor_sync(0xffffffffu, sa, 1 << LO), rb = __shfl_xor_sync(0xffffffffu, sb, 1 << LO);
        x[0] = selp_f(ra, x[0], b0); x[2] = selp_f(rb, x[2], b0);
        x[1] = selp_f(x[1], ra, b0); x[3] = selp_f(x[3], rb, b0);
    }
}
// The gate arithmetic is MUFU bound (EX2 + RCP per sigmoid / tanh: 10 per LSTM cell, 640 cycles per tile and step on the
// SM's 16 lanes), so two activations share one reciprocal: 1/a and 1/b from r = 1/(a*b).  The exponents are clamped at
// +-40 (sigmoid(-40) = 4e-18, 1 - tanh(20) = 8e-18) so that the product of two denominators stays finite.
__device__ __forceinline__ void sigmoid2_f(float xa, float xb, float &sa, float &sb) {
    const float da = 1.0f + __expf(-fmaxf(xa, -40.0f)), db = 1.0f + __expf(-fmaxf(xb, -40.0f));
    const float r = __fdividef(1.0f, da * db);
    sa = db * r;
    sb = da * r;
}
__device__ __forceinline__ void sigmoid_tanh_f(float xs, float xt, float &sg, float &th) {
    const float ds = 1.0f + __expf(-fmaxf(xs, -40.0f)), dt = 1.0f + __expf(fminf(2.0f * xt, 40.0f));
    const float r = __fdividef(1.0f, ds * dt);
    sg = dt * r;
    th = fmaf(-2.0f * ds, r, 1.0f);          // 1 - 2 / (exp(2x) + 1)
}

__device__ __forceinline__ unsigned pack_bf16x2(__nv_bfloat16 a, __nv_bfloat16 b) {
    return (unsigned)__bfloat16_as_ushort(a) | ((unsigned)__bfloat16_as_ushort(b) << 16);
}

template <int CELL>
__global__ void __launch_bounds__(RT_THREADS, 1)
rnn_tc_kernel(const __grid_constant__ CUtensorMap tmap_h, const __grid_constant__ CUtensorMap tmap_x, const RnnTcParams p) {
    constexpr int G = (CELL == DL4SS_CELL_LSTM) ? 4 : 3;
    constexpr uint32_t IDESC_HH = umma_idesc_bf16(128, 2 * RT_NT);   // W_hi x [h_hi ; h_lo]
    constexpr uint32_t IDESC_LH = umma_idesc_bf16(128, RT_NT);       // W_lo x h_hi

    extern __shared__ unsigned char smem_raw[];
    unsigned char *base = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int nkc = p.nkc;
    unsigned char *Hsm = base;                                                   // [RT_TILES][nkc][2][RT_HBLK]
    float *Xsm = reinterpret_cast<float *>(Hsm + (size_t)RT_TILES * nkc * 2 * RT_HBLK);   // [RT_TILES][2][RT_NT][RT_XP]
    uint64_t *bars = reinterpret_cast<uint64_t *>(Xsm + RT_TILES * 2 * RT_XTILE);
    uint64_t *hfull = bars;                                  // [RT_TILES][RT_MAXKC]
    uint64_t *tfull = hfull + RT_TILES * RT_MAXKC;           // [RT_TILES]
    uint64_t *tempty = tfull + RT_TILES;                     // [RT_TILES]
    uint64_t *xfull = tempty + RT_TILES;                     // [RT_TILES][2]
    uint64_t *xempty = xfull + RT_TILES * 2;                 // [RT_TILES][2]
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(xempty + RT_TILES * 2);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    int bid = blockIdx.x;
    const int slice = bid % p.nslices; bid /= p.nslices;
    const int grp = bid % p.ngroups;
    const int dir = bid / p.ngroups;
    const int tile_first = p.tile0 + grp * p.tpg;
    int nt = p.tile0 + p.ntiles - tile_first;       // tiles this CTA serves (1..RT_TILES)
    if (nt > p.tpg) nt = p.tpg;
    const int u0 = slice * RT_HS;
    const int T = p.T, H = p.H;
    const size_t GH = (size_t)G * H;
    unsigned *counters = p.counters + (size_t)(dir * p.tiles_total + tile_first) * RT_CTR_STRIDE;

    if (tid == 0) stamp(p, 0, 2);        // trace: kernel entry
    if (tid == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmap_h) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmap_x) : "memory");
        for (int i = 0; i < RT_TILES * RT_MAXKC; ++i) mbar_init(&hfull[i], 1);
        for (int i = 0; i < RT_TILES; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], RT_EPI_PER_TILE); }
        for (int i = 0; i < RT_TILES * 2; ++i) { mbar_init(&xfull[i], 1); mbar_init(&xempty[i], RT_EPI_PER_TILE); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == RT_W_MGR) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < 4) {
        // resident A operand: TMEM lane = W row 32*warp + lane (row 4*u + g of the slice; units >= H zero),
        // column RT_WCOL + plane*RT_WPLANE + k/2 holds the bf16 pair (k, k+1)
        const int r = 32 * warp + lane;
        const bool rvalid = u0 + (r >> 2) < H;
        const uint32_t ta = tmem_base + ((uint32_t)(32 * warp) << 16) + RT_WCOL;
        const bool shifted = p.yx && dir == 1 && p.hshift == 4;       // k' = k + 4 (hshift is 0 or 4: H % 4 == 0)
        for (int pl = 0; pl < 2; ++pl) {
            const __nv_bfloat16 *row = p.wplanes + ((size_t)pl * 8 * H + (size_t)dir * 4 * H + 4 * u0 + (rvalid ? r : 0)) * p.Kp;
            const uint4 *src = reinterpret_cast<const uint4 *>(row);
            const uint2 *src2 = reinterpret_cast<const uint2 *>(row);       // 4 k positions each
            // four k steps' loads in flight before their stores (Kp is a multiple of 64): the row-per-lane reads are latency bound
            for (int kk0 = 0; kk0 < p.Kp / 16; kk0 += 4) {
                uint4 a[4], b[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int kk = kk0 + j;
                    a[j] = make_uint4(0u, 0u, 0u, 0u); b[j] = a[j];
                    if (rvalid) {
                        if (!shifted) { a[j] = __ldg(src + 2 * kk); b[j] = __ldg(src + 2 * kk + 1); }
                        else {
                            const uint2 z = make_uint2(0u, 0u);
                            const uint2 q0 = kk ? __ldg(src2 + 4 * kk - 1) : z, q1 = __ldg(src2 + 4 * kk);
                            const uint2 q2 = __ldg(src2 + 4 * kk + 1), q3 = __ldg(src2 + 4 * kk + 2);
                            a[j] = make_uint4(q0.x, q0.y, q1.x, q1.y); b[j] = make_uint4(q2.x, q2.y, q3.x, q3.y);
                        }
                    }
                }
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    tmem_st8(ta + pl * RT_WPLANE + (kk0 + j) * 8, a[j].x, a[j].y, a[j].z, a[j].w, b[j].x, b[j].y, b[j].z, b[j].w);
            }
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (tid == 0) stamp(p, 0, 3);        // trace: W_hh resident in tensor memory

    if (warp >= RT_W_MGR && warp < RT_W_MGR + RT_TILES) {
        // ================= tile manager (one warp per tile, one elected lane working): waits for the group's h_{t-1},
        // TMA-loads the 32 x H tile (both planes, 5 k-chunks), issues the step's UMMAs, commits.
        // acc[128 gate rows x 32 utterances] lives in two TMEM column blocks per tile.  UMMAs this small (128x64x16,
        // 128x32x16) execute in ~70 cycles each whatever feeds them (measured: smem or TMEM A operand, split
        // accumulators), so the count is kept at two per k step.  The issue path matters as much: the manager
        // shares its scheduler with six epilogue warps, so (i) the working lane is chosen with elect.sync
        // -- a `lane == 0` branch makes ptxas wrap every UTCHMMA in an elect/branch loop over the
        // possibly-active lanes -- and (ii) every tile has its own manager on its own scheduler, so the three
        // tiles' instruction streams are issued in parallel (one manager for all tiles issued a UMMA every ~115
        // cycles instead of the tensor pipe's ~75).
        const int tl = warp - RT_W_MGR;
        if (tl < nt && elect_one()) {           // the other 31 lanes park at the final barrier
            const uint64_t hdesc0 = umma_desc_sw128(smem_u32(Hsm + (size_t)tl * nkc * 2 * RT_HBLK));
            unsigned char *hs = Hsm + (size_t)tl * nkc * 2 * RT_HBLK;
            const uint32_t a_hi = tmem_base + RT_WCOL, a_lo = a_hi + RT_WPLANE;
            const uint32_t d = tmem_base + tl * RT_TCOLS;
            const int hsh = (p.yx && dir == 1) ? p.hshift : 0;
            const int last_ksteps = (H + hsh - (nkc - 1) * RT_KC + 15) / 16;
            const int ycol0 = dir * H - hsh;                          // yx: first column of this direction's window in y_planes
            const unsigned per_step = (unsigned)p.nslices;            // one release per slice CTA and step
            const unsigned *ctr = counters + tl * RT_CTR_STRIDE;
            // the hoisted input projection of step s: ONE 5-D TMA box [32 utterances][G gates][32 units] (fp32, evict-first:
            // xproj is read exactly once) into Xsm[tile][s & 1]; columns >= H and utterances >= B arrive as zeros
            const uint64_t xpol = l2_evict_first_policy();
            auto load_x = [&](int s) {
                const int buf = s & 1;
                if (s >= 2) mbar_wait(&xempty[tl * 2 + buf], (uint32_t)((s >> 1) - 1) & 1u);
                mbar_expect_tx(&xfull[tl * 2 + buf], RT_NT * G * RT_HS * 4);
                tma_load_5d_hint(Xsm + (size_t)(tl * 2 + buf) * RT_XTILE, &tmap_x, &xfull[tl * 2 + buf], u0, 0, dir,
                                 dir ? (T - 1 - s) : s, (tile_first + tl) * RT_NT, xpol);
            };
            load_x(0);
            if (T > 1) load_x(1);
            for (int s = 1; s < T; ++s) {
                const int z = (((s - 1) & 1) * 2 + dir) * 2;
                const unsigned want = per_step * (unsigned)s;
                if (tl == 0) stamp(p, s, 0);
                // armed before the wait: the previous phase of every chunk barrier was consumed by the last step's UMMAs
                for (int c = 0; c < nkc; ++c) mbar_expect_tx(&hfull[tl * RT_MAXKC + c], 2 * RT_HBLK);     // one box = both planes of the chunk
                while (ld_acquire_gpu(ctr) < want) { __nanosleep(40); }
                if (tl == 0) stamp(p, s, 1);
                fence_proxy_async();             // the group's generic-proxy stores -> this thread's TMA reads
                for (int c = 0; c < nkc; ++c) {
                    if (p.yx)       // rows (b, t of step s-1) of the layer output: [64 columns][1 frame][32 utterances][2 planes]
                        tma_load_4d(hs + (size_t)(c * 2) * RT_HBLK, &tmap_h, &hfull[tl * RT_MAXKC + c], ycol0 + c * RT_KC,
                                    dir ? (T - s) : (s - 1), (tile_first + tl) * RT_NT, 0);
                    else
                        tma_load_3d(hs + (size_t)(c * 2) * RT_HBLK, &tmap_h, &hfull[tl * RT_MAXKC + c], c * RT_KC,
                                    (tile_first + tl) * RT_NT, z);
                }
                if (tl == 0) stamp(p, s, 2);
                if (s >= 2) { mbar_wait(&tempty[tl], (uint32_t)(s - 2) & 1u); tc_fence_after(); }
#pragma unroll
                for (int c = 0; c < RT_MAXKC; ++c) {
                    if (c < nkc) {
                        mbar_wait(&hfull[tl * RT_MAXKC + c], (uint32_t)(s - 1) & 1u);
                        tc_fence_after();
                        if (tl == 0 && c == 0) stamp(p, s, 3);
                        if (tl == 0 && c == nkc - 1) stamp(p, s, 4);
                        const uint64_t h_hi = hdesc0 + (uint64_t)((c * 2 * RT_HBLK) >> 4);
                        const int ksteps = (c == nkc - 1) ? last_ksteps : RT_KC / 16;
#pragma unroll
                        for (int k = 0; k < RT_KC / 16; ++k) {      // B: +32 B per k step (>>4 = 2); A: +8 columns
                            if (k < ksteps) {
                                const int kk = c * (RT_KC / 16) + k;
                                // W_hi x [h_hi ; h_lo] -> columns [0,64) ; then W_lo x h_hi onto [32,64)
                                if (c == 0 && k == 0) umma_bf16_ts<false>(d, a_hi, h_hi, IDESC_HH);
                                else umma_bf16_ts<true>(d, a_hi + kk * 8, h_hi + 2 * k, IDESC_HH);
                                umma_bf16_ts<true>(d + RT_NT, a_lo + kk * 8, h_hi + 2 * k, IDESC_LH);
                            }
                        }
                    }
                }
                umma_commit(&tfull[tl]);
                if (tl == 0) stamp(p, s, 5);
                if (s + 1 < T) load_x(s + 1);       // its buffer was released two steps ago: never waits
            }
        }
    } else {
        // ================= epilogue: gates, state update, publish h_t
        const int sp = warp & 3;                     // TMEM sub-partition (lanes 32*sp..) this warp may read
        const int j4 = warp >> 2;                    // 0..RT_TILES*2-1
        const int tl = j4 >> 1;                      // tile served by this warp
        const int chalf = j4 & 1;                    // batch columns [16*chalf, 16*chalf+16) of the tile
        if (tl < nt) {
            const int g = lane & 3;                  // this lane's accumulator row is gate g of unit ul
            const int g0 = g & 1, g1 = g & 2;
            const int ul = 8 * sp + (lane >> 2);     // unit within the slice (row = 4*ul + g = 32*sp + lane)
            const bool uvalid = u0 + ul < H;
            const int u = u0 + ul;
            const int row0 = (tile_first + tl) * RT_NT;
            const uint32_t taddr = tmem_base + ((uint32_t)(sp * 32) << 16) + tl * RT_TCOLS + chalf * 16;
            unsigned *counter = counters + tl * RT_CTR_STRIDE;
            const bool tr = (blockIdx.x == 0 && warp == 0);      // traced warp
            // output role (after a second 4 x 4 transpose, between the cell index i and lane bits 2-3): this thread holds
            // utterance column oc, units [uq, uq + 4) -> one 16-byte store per fp32 row, 8 bytes per bf16 plane row
            const int a0 = (lane >> 2) & 1, a1b = (lane >> 3) & 1;
            const int oc = chalf * 16 + 4 * ((lane >> 2) & 3) + g;
            const int uq = u0 + 8 * sp + 4 * (lane >> 4);
            const bool ovalid = uq < H && row0 + oc < p.B;
            const bool releaser = (sp == 3 && chalf == 1);       // the warp that publishes the tile's step for this CTA
            int bcol[4];                             // my 4 cells: utterance columns 16*chalf + 4*i + g
#pragma unroll
            for (int i = 0; i < 4; ++i) bcol[i] = chalf * 16 + 4 * i + g;
            float state[4] = {0.f, 0.f, 0.f, 0.f};   // LSTM: c ; GRU: h
            float hsum[4] = {0.f, 0.f, 0.f, 0.f};    // sum over t of the thread's output cells (ADDJUST mean)
            const float bhn = (CELL == DL4SS_CELL_GRU && uvalid) ? __ldg(p.bhn + (size_t)dir * H + u) : 0.f;

            for (int s = 0; s < T; ++s) {
                const int t = dir ? (T - 1 - s) : s;
                const int buf = s & 1;
                // the step's input projections first: their tile landed two steps ago, and the warp would otherwise idle on the
                // accumulator barrier (the 16 shared-memory reads + swizzle arithmetic were ~0.4 k cycles of the chain)
                mbar_wait(&xfull[tl * 2 + buf], (uint32_t)(s >> 1) & 1u);
                if (tr && lane == 0) stamp(p, s, 11);
                const float *xs = Xsm + (size_t)(tl * 2 + buf) * RT_XTILE;
                float xv[4][G];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int gg = 0; gg < G; ++gg) {
                        const int row = bcol[i] * G + gg;                 // 128-byte row of the box; 16-byte chunks XOR (row & 7)
                        xv[i][gg] = xs[row * RT_HS + ((((ul >> 2) ^ (row & 7)) << 2) | (ul & 3))];
                    }
                __syncwarp();
                if (lane == 0) mbar_arrive(&xempty[tl * 2 + buf]);
                if (tr && lane == 0) stamp(p, s, 15);
                float G4[4][4];                      // [cell i][gate] recurrent pre-activations
                if (s >= 1) {
                    mbar_wait(&tfull[tl], (uint32_t)(s - 1) & 1u);
                    tc_fence_after();
                    if (tr && lane == 0) stamp(p, s, 6);
                    float v[16], a1[16];
                    tmem_ld16(taddr + RT_NT, a1);            // W_hi * h_lo + W_lo * h_hi
                    tmem_ld16(taddr, v);                     // W_hi * h_hi
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] += a1[j];
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tempty[tl]);
                    if (tr && lane == 0) stamp(p, s, 7);
                    // 4-lane transpose: lane g holds gate g for columns 4i..4i+3, wants gates 0..3 of column 4i+g
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        float q[4] = {v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]};
                        transpose4<0>(q, g0, g1);
#pragma unroll
                        for (int gg = 0; gg < 4; ++gg) G4[i][gg] = q[gg];
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int gg = 0; gg < 4; ++gg) G4[i][gg] = 0.f;
                }
                if (tr && lane == 0) stamp(p, s, 8);

                float hnew[4], gv[4][G], aux[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    if constexpr (CELL == DL4SS_CELL_LSTM) {
                        float ig, fg, gg, og;
                        sigmoid2_f(xv[i][0] + G4[i][0], xv[i][1] + G4[i][1], ig, fg);
                        sigmoid_tanh_f(xv[i][3] + G4[i][3], xv[i][2] + G4[i][2], og, gg);
                        const float c = fmaf(fg, state[i], ig * gg);
                        state[i] = c;
                        hnew[i] = og * tanh_f(c);
                        gv[i][0] = ig; gv[i][1] = fg; gv[i][2] = gg; gv[i][3] = og;
                        aux[i] = c;
                    } else {
                        float rg, zg;
                        sigmoid2_f(xv[i][0] + G4[i][0], xv[i][1] + G4[i][1], rg, zg);
                        const float hn = G4[i][2] + bhn;
                        const float ng = tanh_f(fmaf(rg, hn, xv[i][2]));
                        hnew[i] = fmaf(zg, state[i] - ng, ng);          // (1-z)*n + z*h
                        state[i] = hnew[i];
                        gv[i][0] = rg; gv[i][1] = zg; gv[i][2] = ng;
                        aux[i] = hn;                                    // W_hn*h + b_hn, kept for backward
                    }
                }
                if (tr && lane == 0) stamp(p, s, 9);
                // publish h_t first (the group's next step hangs on it); the fp32 outputs follow off the critical path
                // second transpose: 4 utterances x 1 unit -> 1 utterance x 4 consecutive units, so that h / y leave in
                // 8- and 16-byte pieces (5 stores per thread instead of 20: fewer L2 write transactions under the release)
                float ho[4] = {hnew[0], hnew[1], hnew[2], hnew[3]};
                transpose4<2>(ho, a0, a1b);
                __nv_bfloat16 hhi[4], hlo[4];
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    hhi[r] = __float2bfloat16_rn(ho[r]);
                    hlo[r] = __float2bfloat16_rn(ho[r] - __bfloat162float(hhi[r]));
                    hsum[r] += ho[r];
                }
                const uint2 phi = make_uint2(pack_bf16x2(hhi[0], hhi[1]), pack_bf16x2(hhi[2], hhi[3]));
                const uint2 plo = make_uint2(pack_bf16x2(hlo[0], hlo[1]), pack_bf16x2(hlo[2], hlo[3]));
                if (p.yx && ovalid) {       // the layer output row is the publish
                    const size_t o = ((size_t)(row0 + oc) * T + t) * p.Kpy + (size_t)dir * H + uq;
                    *reinterpret_cast<uint2 *>(p.y_planes + o) = phi;
                    *reinterpret_cast<uint2 *>(p.y_planes + (size_t)p.B * T * p.Kpy + o) = plo;
                }
                if (s + 1 < T) {
                    if (ovalid && !p.yx) {
                        const int pp = s & 1;
                        __nv_bfloat16 *hh = p.hbuf + ((size_t)((pp * 2 + dir) * 2) * p.Bpad + row0 + oc) * p.Kp + uq;
                        *reinterpret_cast<uint2 *>(hh) = phi;
                        *reinterpret_cast<uint2 *>(hh + (size_t)p.Bpad * p.Kp) = plo;
                    }
                    if (tr && lane == 0) stamp(p, s, 10);
                    // one release per (CTA, tile, step): the tile's other warps only arrive on the named barrier, the releasing
                    // warp waits for them; its MEMBAR.GPU + RED then covers their stores (cumulativity through the barrier).
                    // 8 x fewer atomics on the group's counter line, which serialise at ~27 cycles each in L2
                    if (releaser) {
                        const bool trr = (blockIdx.x == 0 && tl == 0);
                        if (trr && lane == 0) stamp(p, s, 12);
                        asm volatile("bar.sync %0, %1;\n" ::"r"(1 + tl), "n"(RT_EPI_PER_TILE * 32) : "memory");
                        if (lane == 0) {
                            const bool tr = trr;
                            if (tr) stamp(p, s, 13);
                            asm volatile("red.release.gpu.global.add.u32 [%0], %1;\n" ::"l"(counter), "r"(1u) : "memory");
                            if (tr) stamp(p, s, 14);
                        }
                    } else {
                        asm volatile("bar.arrive %0, %1;\n" ::"r"(1 + tl), "n"(RT_EPI_PER_TILE * 32) : "memory");
                    }
                }
                if (ovalid) {
                    const int b = row0 + oc;
                    if (p.y != nullptr)      // inference with planes: nobody reads the fp32 copy (half of the layer's output bytes)
                        *reinterpret_cast<float4 *>(p.y + ((size_t)b * T + t) * 2 * H + (size_t)dir * H + uq) = make_float4(ho[0], ho[1], ho[2], ho[3]);
                    if (p.y_planes != nullptr && !p.yx) {
                        const size_t o = ((size_t)b * T + t) * p.Kpy + (size_t)dir * H + uq;
                        *reinterpret_cast<uint2 *>(p.y_planes + o) = phi;
                        *reinterpret_cast<uint2 *>(p.y_planes + (size_t)p.B * T * p.Kpy + o) = plo;
                    }
                }
                if (uvalid && (p.gates_save != nullptr || p.cell_save != nullptr)) {       // training only: kept per cell
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int b = row0 + bcol[i];
                        if (b < p.B) {
                            if (p.gates_save != nullptr) {
                                float *go = p.gates_save + (((size_t)b * T + t) * 2 + dir) * GH + u;
#pragma unroll
                                for (int gg = 0; gg < G; ++gg) go[(size_t)gg * H] = gv[i][gg];
                            }
                            if (p.cell_save != nullptr)
                                p.cell_save[(((size_t)b * T + t) * 2 + dir) * H + u] = aux[i];
                        }
                    }
                }
            }
            if (p.hmean_out != nullptr && ovalid)
                *reinterpret_cast<float4 *>(p.hmean_out + (size_t)(row0 + oc) * 2 * H + (size_t)dir * H + uq) =
                    make_float4(hsum[0] / (float)T, hsum[1] / (float)T, hsum[2] / (float)T, hsum[3] / (float)T);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (tid == 0) stamp(p, 0, 4);        // trace: all steps done
    if (warp == RT_W_MGR) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

// whh fp32 [2][G*H][H] -> bf16 planes [2 (hi,lo)][2 dir * 4H rows][Kp]; row dir*4H + 4*u + g holds gate g of unit u
// (g >= G: zeros), k zero padded to Kp
__global__ void __launch_bounds__(256)
pack_whh_kernel(const float *__restrict__ whh, int G, int H, int Kp, __nv_bfloat16 *__restrict__ planes) {
    const long long rows = 2ll * 4 * H;
    const long long total = rows * Kp;
    __nv_bfloat16 *hi = planes, *lo = planes + total;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / Kp;
        const int k = (int)(i - r * Kp);
        const int dir = (int)(r / (4 * H));
        const int q = (int)(r - (long long)dir * 4 * H);
        const int u = q >> 2, g = q & 3;
        float v = 0.f;
        if (g < G && k < H) v = whh[((size_t)dir * G * H + (size_t)g * H + u) * H + k];
        const __nv_bfloat16 h = __float2bfloat16_rn(v);
        hi[i] = h;
        lo[i] = __float2bfloat16_rn(v - __bfloat162float(h));
    }
}

// zero columns [c0, c1) and [c2, Kpy) of every row of the planes (all multiples of 8 columns: whole 16-byte pieces)
__global__ void __launch_bounds__(256)
zero_plane_cols_kernel(__nv_bfloat16 *planes, long long rows, int Kpy, int c0, int c1, int c2) {
    const int n1 = (c1 - c0) >> 3, n2 = (Kpy - c2) >> 3, per = n1 + n2;
    const long long total = rows * per;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / per;
        const int q = (int)(i - r * per);
        const int col = q < n1 ? c0 + 8 * q : c2 + 8 * (q - n1);
        *reinterpret_cast<uint4 *>(planes + r * Kpy + col) = make_uint4(0u, 0u, 0u, 0u);
    }
}

static long long *g_trace = nullptr;
static int g_trace_steps = 0;
static int g_cluster_pairs = [] { const char *e = getenv("DL4SS_RNN_CLUSTER_PAIRS"); return e ? atoi(e) : 0; }();
static int g_yx = [] { const char *e = getenv("DL4SS_RNN_YX"); return e ? atoi(e) : 1; }();
static int g_tiles_per_cta = [] { const char *e = getenv("DL4SS_RNN_TILES_PER_CTA"); return e ? atoi(e) : 0; }();

template <int CELL>
static int launch_rnn_tc(RnnTcParams p, const void *whh_planes, int tiles_left, cudaStream_t st, int *launched_tiles) {
    const size_t smem = 1024 + (size_t)RT_TILES * p.nkc * 2 * RT_HBLK +
                        (size_t)RT_TILES * 2 * RT_XTILE * sizeof(float) + 512;
    auto kern = rnn_tc_kernel<CELL>;
    DL4SS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    DL4SS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, RT_THREADS, smem));
    const int max_groups = per_sm * sm_count() / (2 * p.nslices);
    if (max_groups < 1) {
        set_error("rnn_layer_tc_fwd: %d co-resident CTAs cannot hold one group (%d slices x 2 directions)",
                  per_sm * sm_count(), p.nslices);
        return DL4SS_EUNSUPPORTED;
    }
    // tiles per CTA: as few as the co-resident CTAs allow (shortest step), unless dl4ss_rnn_tc_set_tiles_per_cta asks
    // for denser CTAs (fewer SMs per launch: two launches on different streams then run side by side)
    int ntiles = tiles_left;
    if (ntiles > max_groups * RT_TILES) ntiles = max_groups * RT_TILES;
    int tpg = cdiv(ntiles, max_groups);
    if (g_tiles_per_cta > tpg) tpg = g_tiles_per_cta < RT_TILES ? g_tiles_per_cta : RT_TILES;
    const int ngroups = cdiv(ntiles, tpg);
    p.ntiles = ntiles; p.tpg = tpg; p.ngroups = ngroups;
    *launched_tiles = ntiles;

    CUtensorMap mh;
    p.wplanes = (const __nv_bfloat16 *)whh_planes;
    if (p.yx) {   // the layer output planes bf16 [2][B][T][Kpy]: box = 64 columns x 1 frame x 32 utterances x both planes
        cuuint64_t dims[4] = {(cuuint64_t)p.Kpy, (cuuint64_t)p.T, (cuuint64_t)p.B, 2};
        cuuint64_t strides[3] = {(cuuint64_t)p.Kpy * 2, (cuuint64_t)p.T * p.Kpy * 2, (cuuint64_t)p.B * p.T * p.Kpy * 2};
        cuuint32_t box[4] = {RT_KC, 1, RT_NT, 2};
        int rc = make_bf16_map(&mh, p.y_planes, 4, dims, strides, box);
        if (rc) return rc;
    } else {   // h exchange bf16 [8 = pp,dir,plane][Bpad][Kp]: box = 64 k x 32 rows x both planes
        cuuint64_t dims[3] = {(cuuint64_t)p.Kp, (cuuint64_t)p.Bpad, 8};
        cuuint64_t strides[2] = {(cuuint64_t)p.Kp * 2, (cuuint64_t)p.Bpad * p.Kp * 2};
        cuuint32_t box[3] = {RT_KC, RT_NT, 2};
        int rc = make_bf16_map(&mh, p.hbuf, 3, dims, strides, box);
        if (rc) return rc;
    }
    CUtensorMap mx;
    {   // xproj fp32 [B][T][2][G][H]: box = 32 units x G gates x 1 direction x 1 frame x 32 utterances
        const cuuint64_t G = (CELL == DL4SS_CELL_LSTM) ? 4 : 3;
        cuuint64_t dims[5] = {(cuuint64_t)p.H, G, 2, (cuuint64_t)p.T, (cuuint64_t)p.B};
        cuuint64_t strides[4] = {(cuuint64_t)p.H * 4, G * p.H * 4, 2 * G * p.H * 4, (cuuint64_t)p.T * 2 * G * p.H * 4};
        cuuint32_t box[5] = {RT_HS, (cuuint32_t)G, 1, 1, RT_NT};
        int rc = make_f32_map(&mx, p.xproj, 5, dims, strides, box);
        if (rc) return rc;
    }
    void *args[] = {(void *)&mh, (void *)&mx, (void *)&p};
    if (g_cluster_pairs) {
        // placement only: as 2-CTA clusters the launch packs TPCs (60 CTAs on 30 TPCs instead of one SM of 60 TPCs), which leaves
        // whole TPCs to a 2-CTA projection launch of the other in-flight batch
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(2 * ngroups * p.nslices);
        cfg.blockDim = dim3(RT_THREADS);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = st;
        cudaLaunchAttribute attr[2];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        attr[1].id = cudaLaunchAttributeCooperative;
        attr[1].val.cooperative = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 2;
        DL4SS_CUDA(cudaLaunchKernelExC(&cfg, (const void *)kern, args));
    } else {
        DL4SS_CUDA(cudaLaunchCooperativeKernel((void *)kern, dim3(2 * ngroups * p.nslices), dim3(RT_THREADS), args, smem, st));
    }
    count_launch();
    return DL4SS_OK;
}

static bool rnn_tc_supported(int H) { return H >= 4 && H % 4 == 0 && H <= RT_MAXKC * RT_KC; }

}  // namespace dl4ss

using namespace dl4ss;

// profiling hook: device buffer of steps*16 int64 receiving CTA 0's per-phase clock64() stamps (null = off)
extern "C" void dl4ss_rnn_tc_set_trace(void *dev_buf, int steps) {
    g_trace = (long long *)dev_buf;
    g_trace_steps = dev_buf ? steps : 0;
}

extern "C" void dl4ss_rnn_tc_set_tiles_per_cta(int tiles) { g_tiles_per_cta = tiles; }
extern "C" void dl4ss_rnn_tc_set_cluster_pairs(int on) { g_cluster_pairs = on; }

extern "C" int dl4ss_rnn_tc_supported(int H, int cell) {
    return (cell == DL4SS_CELL_LSTM || cell == DL4SS_CELL_GRU) && rnn_tc_supported(H) ? 1 : 0;
}

extern "C" size_t dl4ss_rnn_tc_whh_bytes(int H) {
    if (H <= 0) return 0;
    const size_t Kp = (size_t)cdiv(H, RT_KC) * RT_KC;
    return 2 * (size_t)8 * H * Kp * sizeof(__nv_bfloat16);
}

extern "C" int dl4ss_rnn_tc_pack_whh(int cell, const float *whh, int H, void *planes, void *stream) {
    DL4SS_CHECK_ARG(cell == DL4SS_CELL_LSTM || cell == DL4SS_CELL_GRU, "rnn_tc_pack_whh: bad cell %d", cell);
    DL4SS_CHECK_ARG(whh && planes && H >= 1, "rnn_tc_pack_whh: null operand / bad H");
    DL4SS_CHECK_ARG((((uintptr_t)planes) & 15) == 0, "rnn_tc_pack_whh: planes must be 16-byte aligned");
    const int Kp = cdiv(H, RT_KC) * RT_KC;
    const long long total = 8ll * H * Kp;
    long long blocks = cdivll(total, 256);
    if (blocks > (long long)sm_count() * 8) blocks = (long long)sm_count() * 8;
    pack_whh_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(whh, cell == DL4SS_CELL_LSTM ? 4 : 3, H, Kp,
                                                                         (__nv_bfloat16 *)planes);
    DL4SS_LAUNCH_CHECK("pack_whh_kernel");
    return DL4SS_OK;
}

extern "C" size_t dl4ss_rnn_tc_workspace_bytes(int B, int T, int H, int cell) {
    (void)T; (void)cell;
    if (B <= 0 || H <= 0) return 256;
    const size_t Bpad = (size_t)cdiv(B, RT_NT) * RT_NT;
    const size_t Kp = (size_t)cdiv(H, RT_KC) * RT_KC;
    const size_t ctr = (size_t)2 * (Bpad / RT_NT) * RT_CTR_STRIDE * sizeof(unsigned);
    return ctr + 8 * Bpad * Kp * sizeof(__nv_bfloat16);
}

extern "C" int dl4ss_rnn_layer_tc_fwd(int cell, const float *xproj, const void *whh_planes, const float *bhn,
                                      float *y, int B, int T, int H, float *gates_save, float *cell_save,
                                      void *y_planes, float *hmean_out,
                                      void *workspace, size_t workspace_bytes, void *stream) {
    DL4SS_CHECK_ARG(cell == DL4SS_CELL_LSTM || cell == DL4SS_CELL_GRU, "rnn_layer_tc_fwd: bad cell %d", cell);
    DL4SS_CHECK_ARG(xproj && whh_planes && (y || y_planes), "rnn_layer_tc_fwd: null operand (y may be NULL only when y_planes is given)");
    DL4SS_CHECK_ARG((((uintptr_t)xproj) & 15) == 0, "rnn_layer_tc_fwd: xproj must be 16-byte aligned");
    DL4SS_CHECK_ARG(cell == DL4SS_CELL_LSTM || bhn, "rnn_layer_tc_fwd: GRU needs bhn");
    DL4SS_CHECK_ARG(B >= 0 && T >= 1 && H >= 1, "rnn_layer_tc_fwd: bad B/T/H %d/%d/%d", B, T, H);
    if (!rnn_tc_supported(H)) {
        set_error("rnn_layer_tc_fwd: H=%d unsupported (needs a multiple of 4, <= %d); use dl4ss_rnn_layer_fwd",
                  H, RT_MAXKC * RT_KC);
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
    p.y_planes = (__nv_bfloat16 *)y_planes; p.hmean_out = hmean_out;
    p.Kpy = cdiv(2 * H, RT_KC) * RT_KC;
    p.B = B; p.T = T; p.H = H;
    p.nkc = cdiv(H, RT_KC);
    p.Kp = p.nkc * RT_KC;
    p.Bpad = cdiv(B, RT_NT) * RT_NT;
    p.tiles_total = p.Bpad / RT_NT;
    p.nslices = cdiv(H, RT_HS);
    p.counters = (unsigned *)workspace;
    const size_t ctr = (size_t)2 * p.tiles_total * RT_CTR_STRIDE * sizeof(unsigned);
    p.hbuf = (__nv_bfloat16 *)((unsigned char *)workspace + ctr);
    p.ntiles = p.tpg = p.ngroups = 0;
    p.trace = g_trace; p.trace_steps = g_trace_steps;
    p.hshift = H % 8;
    p.yx = (y_planes != nullptr && g_yx && (((uintptr_t)y_planes) & 15) == 0) ? 1 : 0;
    if (p.yx) {
        // the columns a direction's window shares with the other direction, the 16-wide k step's overhang and the row padding
        // are read before anybody has written them: they meet zero weights, so they only have to be finite -- zeroed here
        const int c0 = H - p.hshift, c1 = (H + 15) / 16 * 16 < p.Kpy ? (H + 15) / 16 * 16 : p.Kpy;
        const long long rows = 2ll * B * T, total = rows * (((c1 - c0) >> 3) + ((p.Kpy - 2 * H) >> 3));
        long long blocks = cdivll(total, 256);
        if (blocks > (long long)sm_count() * 8) blocks = (long long)sm_count() * 8;
        if (total > 0) {            // H a multiple of 64: the windows are exact, nothing to zero
            zero_plane_cols_kernel<<<(unsigned)blocks, 256, 0, st>>>((__nv_bfloat16 *)y_planes, rows, p.Kpy, c0, c1, 2 * H);
            DL4SS_LAUNCH_CHECK("zero_plane_cols_kernel");
        }
    }
    int t0 = 0;
    while (t0 < p.tiles_total) {
        p.tile0 = t0;
        int done = 0;
        int rc = (cell == DL4SS_CELL_LSTM) ? launch_rnn_tc<DL4SS_CELL_LSTM>(p, whh_planes, p.tiles_total - t0, st, &done)
                                           : launch_rnn_tc<DL4SS_CELL_GRU>(p, whh_planes, p.tiles_total - t0, st, &done);
        if (rc) return rc;
        t0 += done;
    }
    return DL4SS_OK;
}
