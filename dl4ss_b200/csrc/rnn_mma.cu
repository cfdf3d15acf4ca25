// dl4ss_rnn_layer_mma_fwd : one bidirectional LSTM / GRU layer as ONE persistent kernel for hidden sizes the
// TMEM-resident tcgen05 form (rnn_tc.cu, H <= 320) cannot hold -- the speaker classifier's BLSTM 3x600
// (MIX_SPEECH_classifier, TDAA_beta/main_run_sstune_EvalVer.py:305-326; nn.LSTM call :310-316).
//
// With H = 600 the two directions' W_hh are 11.5 MB as bf16 hi/lo planes: a third of all shared memory on the chip,
// so W can be resident only ONCE.  A CTA therefore owns (direction, slice of 10 hidden units) -- 2 x 60 = 120 CTAs --
// and walks EVERY utterance tile of the batch each time step:
//   * the slice's 40 gate rows of W_hh (unit-major: column 4*j + gate) stay in shared memory as bf16 hi/lo planes
//     for all T steps (104 KB at H = 600);
//   * per (step, tile of 16 utterances) the product [16 x H] x [H x 40] runs on warp-level mma.sync.m16n8k16
//     (bf16x3: hi*hi + hi*lo + lo*hi, fp32 accumulation).  K is split over the five MMA warps of a group, each warp
//     against all five 8-column n-tiles, so one A fragment pair feeds 15 MMAs (fragments come in with ldmatrix: 8
//     instructions, 28 shared-memory wavefronts per k step; the
//     first form, one n-tile per warp over the whole K, needed 60 and ran at 84 % of the shared-memory pipe); the
//     partial tiles meet in shared memory (two named barriers of the group's 160 threads), where every lane picks up
//     the four gate pre-activations of ONE cell (16 rows x 10 units = 160 cells per group), adds the hoisted input
//     projection (prefetched before the wait), applies the gates, keeps c / h of its cell in shared memory;
//   * the 60 slices of a (direction, tile) exchange h_t through an L2-resident bf16 hi/lo ping-pong buffer (release
//     counter per tile, as K3); a tile's 16 rows x 2 planes arrive as bulk copies on mbarriers, in two K halves with
//     their own full / empty barriers: while the MMA warps work on the second half the loader already refills the
//     first half with the group's next tile;
//   * two warp groups serve alternate tiles with their own h buffer, loader warp and RELEASE warp: the ~2 k cycle
//     membar + red of a publish runs on the release warp while the group's MMA warps are already on their next tile,
//     and one group's exchange latency hides behind the other group's product.  With many tiles per CTA (B = 256:
//     16 per direction) the kernel is throughput bound on the product instead of latency bound on the exchange.
// tcgen05 would need the 128-row M for 40 gate rows and a TMEM-resident W that does not fit (640 of 512 columns).
#include "tc_ptx.cuh"

namespace dl4ss {

constexpr int MM_BT = 16;                       // utterances per tile (one m16 tile)
constexpr int MM_HS = 10;                       // hidden units per slice
constexpr int MM_N = 4 * MM_HS;                 // gate columns per slice, unit-major (GRU: 4th column of a unit is zero)
constexpr int MM_NT = MM_N / 8;                 // n-tiles = MMA warps per group
constexpr int MM_GROUPS = 2;
constexpr int MM_WPG = MM_NT + 2;               // warps per group: MMA warps, loader, releaser
constexpr int MM_THREADS = 32 * MM_GROUPS * MM_WPG;
constexpr int MM_MAX_TILES = 16;                // tiles per launch (cell state of every tile lives in shared memory)
constexpr int MM_CTR_STRIDE = 64;

struct RnnMmaParams {
    const float *xproj;        // [B,T,2,G*H]
    const float *whh;          // [2,G*H,H]
    const float *bhn;          // [2,H] (GRU) or null
    float *y;                  // [B,T,2H]
    float *gates_save;         // [B,T,2,G*H] or null
    float *cell_save;          // [B,T,2,H] or null
    __nv_bfloat16 *hx;         // [2 ping-pong][2 dir][2 plane][B][Kg]
    unsigned *counters;        // [2][ntiles] * MM_CTR_STRIDE (this launch)
    int B, T, H, Kg, pitch, nks;
    int b_begin, ntiles, nslices;
};

__device__ __forceinline__ unsigned mm_ld_acquire(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void mm_mma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// four / two 8x8 b16 matrices: lane i supplies the 16-byte row address of row (i % 8) of matrix (i / 8)
__device__ __forceinline__ void mm_ldsm4(uint32_t (&r)[4], uint32_t saddr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(saddr));
}
__device__ __forceinline__ void mm_ldsm2(uint32_t &r0, uint32_t &r1, uint32_t saddr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];\n" : "=r"(r0), "=r"(r1) : "r"(saddr));
}
__device__ __forceinline__ void mm_bulk_g2s(void *smem, const void *gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                 ::"r"(smem_u32(smem)), "l"(gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

template <int CELL>
__global__ void __launch_bounds__(MM_THREADS, 1)
rnn_mma_kernel(const RnnMmaParams p) {
    constexpr int G = (CELL == DL4SS_CELL_LSTM) ? 4 : 3;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int H = p.H, T = p.T, pitch = p.pitch, Kg = p.Kg;
    const size_t GH = (size_t)G * H;
    __nv_bfloat16 *Wt = reinterpret_cast<__nv_bfloat16 *>(smem_raw);                  // [2][MM_N][pitch]
    __nv_bfloat16 *hbuf = Wt + (size_t)2 * MM_N * pitch;                              // [MM_GROUPS][2][MM_BT][pitch]
    float *cstate = reinterpret_cast<float *>(hbuf + (size_t)MM_GROUPS * 2 * MM_BT * pitch);   // [MM_MAX_TILES][MM_BT][MM_HS]
    float *part = cstate + MM_MAX_TILES * MM_BT * MM_HS;                              // [MM_GROUPS][MM_NT][MM_BT][MM_N] K-split partial sums
    uint64_t *bars = reinterpret_cast<uint64_t *>(part + MM_GROUPS * MM_NT * MM_BT * MM_N);
    // the h buffer of a group is handed over in two K halves, each with its own full / empty barrier: while the MMA warps
    // work on the second half the loader already refills the first half with the group's next tile
    uint64_t *hfull = bars, *hempty = bars + 2 * MM_GROUPS, *rel = bars + 4 * MM_GROUPS, *reldone = bars + 5 * MM_GROUPS;
    const int ks_half = (p.nks + 1) >> 1;                     // k steps of the first half
    const int k_half = min(ks_half * 16, Kg);                 // elements of a row in the first half (16-byte multiple)

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int grp = warp / MM_WPG, wg = warp - grp * MM_WPG;
    const int slice = blockIdx.x % p.nslices;
    const int dir = blockIdx.x / p.nslices;
    const int u0 = slice * MM_HS;
    const size_t hx_plane = (size_t)p.B * Kg;                 // elements of one (pp, dir, plane) block
    unsigned *counters = p.counters + (size_t)dir * p.ntiles * MM_CTR_STRIDE;

    {   // zero the operand arrays and the cell state, then the resident W planes
        uint4 *z = reinterpret_cast<uint4 *>(smem_raw);
        const int n16 = (int)((reinterpret_cast<unsigned char *>(bars) - smem_raw) / 16);
        for (int i = tid; i < n16; i += MM_THREADS) z[i] = make_uint4(0u, 0u, 0u, 0u);
        __syncthreads();
        const float *wsrc = p.whh + (size_t)dir * GH * H;
        for (int i = tid; i < G * MM_HS * H; i += MM_THREADS) {
            const int k = i % H, r = i / H;                   // r = g * MM_HS + j
            const int g = r / MM_HS, j = r - g * MM_HS;
            const float v = __ldg(wsrc + ((size_t)g * H + u0 + j) * H + k);
            const __nv_bfloat16 hi = __float2bfloat16_rn(v);
            Wt[(size_t)(4 * j + g) * pitch + k] = hi;
            Wt[(size_t)(MM_N + 4 * j + g) * pitch + k] = __float2bfloat16_rn(v - __bfloat162float(hi));
        }
        if (tid == 0) {
            for (int g = 0; g < MM_GROUPS; ++g) {
                for (int h = 0; h < 2; ++h) {
                    mbar_init(&hfull[2 * g + h], 1);
                    mbar_init(&hempty[2 * g + h], MM_NT);
                }
                mbar_init(&rel[g], MM_NT);
                mbar_init(&reldone[g], 1);
            }
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();

    __nv_bfloat16 *hb = hbuf + (size_t)grp * 2 * MM_BT * pitch;

    if (wg == MM_NT) {
        // ================= loader: h_{t-1} of the group's next tile as soon as its 60 slices have published it
        uint32_t ph_e = 0;
        int nload = 0;
        for (int s = 1; s < T; ++s) {
            const int pp = (s - 1) & 1;
            for (int tau = grp; tau < p.ntiles; tau += MM_GROUPS) {
                const int row0 = p.b_begin + tau * MM_BT;
                const int nrows = min(MM_BT, p.B - row0);
                const int r = lane >> 1, pl = lane & 1;
                const __nv_bfloat16 *src = p.hx + ((size_t)((pp * 2 + dir) * 2 + pl)) * hx_plane + (size_t)(row0 + r) * Kg;
                __nv_bfloat16 *dst = hb + (size_t)(pl * MM_BT + r) * pitch;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int e0 = h ? k_half : 0, ne = h ? Kg - k_half : k_half;     // elements of this half
                    if (nload > 0) mbar_wait(&hempty[2 * grp + h], ph_e);
                    if (lane == 0) {
                        if (h == 0) {
                            const unsigned want = (unsigned)p.nslices * (unsigned)s;
                            while (mm_ld_acquire(counters + (size_t)tau * MM_CTR_STRIDE) < want) { __nanosleep(20); }
                        }
                        if (ne > 0) mbar_expect_tx(&hfull[2 * grp + h], (uint32_t)(nrows * 2 * ne * 2));
                        else mbar_arrive(&hfull[2 * grp + h]);
                    }
                    __syncwarp();
                    if (h == 0) asm volatile("fence.proxy.async.global;\n" ::: "memory");
                    if (r < nrows && ne > 0) mm_bulk_g2s(dst + e0, src + e0, (uint32_t)(ne * 2), &hfull[2 * grp + h]);
                }
                if (nload > 0) ph_e ^= 1;
                ++nload;
            }
        }
    } else if (wg == MM_NT + 1) {
        // ================= releaser: publishes a tile's h_t once the group's MMA warps have stored it
        uint32_t ph = 0;
        for (int s = 0; s + 1 < T; ++s) {
            for (int tau = grp; tau < p.ntiles; tau += MM_GROUPS) {
                mbar_wait(&rel[grp], ph); ph ^= 1;
                if (lane == 0) {
                    mbar_arrive(&reldone[grp]);
                    asm volatile("red.release.gpu.global.add.u32 [%0], %1;\n"
                                 ::"l"(counters + (size_t)tau * MM_CTR_STRIDE), "r"(1u) : "memory");
                }
            }
        }
    } else {
        // ================= MMA warps: k steps wg, wg + 5, ... against all n-tiles; then the cells of units 2*wg, 2*wg+1
        const int fg = lane >> 2, ft = lane & 3;
        const uint32_t *A32 = reinterpret_cast<const uint32_t *>(hb);
        const uint32_t *W32 = reinterpret_cast<const uint32_t *>(Wt);
        const int pw = pitch >> 1;
        // ldmatrix row addresses (bytes, shared window).  A (16 x 16 of h): matrices (rows 0-7, k 0-7), (rows 8-15, k 0-7),
        // (rows 0-7, k 8-15), (rows 8-15, k 8-15) = fragment registers a0..a3.  B (W, row = gate column n, k contiguous):
        // matrices (n-tile j, k 0-7), (n-tile j, k 8-15), (n-tile j+1, k 0-7), (n-tile j+1, k 8-15) = b0, b1 of two n-tiles.
        const int lrow = lane & 7, lmat = lane >> 3;
        const uint32_t a_hi_addr = smem_u32(hb) + (uint32_t)(((lmat & 1) * 8 + lrow) * pitch + (lmat >> 1) * 8) * 2;
        const uint32_t a_lo_addr = a_hi_addr + (uint32_t)(MM_BT * pitch) * 2;
        const uint32_t b_hi_addr = smem_u32(Wt) + (uint32_t)(((lmat >> 1) * 8 + lrow) * pitch + (lmat & 1) * 8) * 2;
        const uint32_t b_lo_addr = b_hi_addr + (uint32_t)(MM_N * pitch) * 2;
        (void)A32; (void)W32; (void)pw;
        float *pmine = part + (size_t)(grp * MM_NT + wg) * (MM_BT * MM_N);
        const float *pgrp = part + (size_t)grp * MM_NT * (MM_BT * MM_N);
        // the cell this lane finishes: even ft -> row fg, odd ft -> row fg + 8; unit 2*wg + (ft >> 1)
        const int crow = (ft & 1) ? fg + 8 : fg;
        const int cj = 2 * wg + (ft >> 1);
        const int u = u0 + cj;
        const float bhn = (CELL == DL4SS_CELL_GRU) ? __ldg(p.bhn + (size_t)dir * H + u) : 0.f;
        uint32_t ph_f = 0, ph_rd = 0;
        int nrel = 0;

        for (int s = 0; s < T; ++s) {
            const int t = dir ? (T - 1 - s) : s;
            for (int tau = grp; tau < p.ntiles; tau += MM_GROUPS) {
                const int b = p.b_begin + tau * MM_BT + crow;
                const bool valid = b < p.B;
                const size_t xrow = (((size_t)(valid ? b : 0) * T + t) * 2 + dir);
                float xv[G];
#pragma unroll
                for (int g = 0; g < G; ++g) xv[g] = valid ? __ldg(p.xproj + xrow * GH + (size_t)g * H + u) : 0.f;

                float gt[4] = {0.f, 0.f, 0.f, 0.f};
                if (s > 0) {
                    // K is split over the group's five warps (k step = warp, warp + 5, ...), each warp against all five
                    // n-tiles: one A fragment pair feeds 15 MMAs (28 shared-memory loads per k step instead of 60)
                    float acc[MM_NT][4];
#pragma unroll
                    for (int n = 0; n < MM_NT; ++n)
#pragma unroll
                        for (int i = 0; i < 4; ++i) acc[n][i] = 0.f;
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        mbar_wait(&hfull[2 * grp + h], ph_f);
                        const int ks1 = h ? p.nks : ks_half;
                        for (int ks = (h ? ks_half : 0) + wg; ks < ks1; ks += MM_NT) {
                            const uint32_t kb = (uint32_t)ks * 32;                 // bytes of 16 k
                            uint32_t ah[4], al[4];
                            mm_ldsm4(ah, a_hi_addr + kb);
                            mm_ldsm4(al, a_lo_addr + kb);
#pragma unroll
                            for (int n = 0; n < MM_NT; n += 2) {
                                uint32_t bhv[4], blv[4];
                                const uint32_t off = (uint32_t)(8 * n * pitch) * 2 + kb;
                                if (n + 1 < MM_NT) {
                                    mm_ldsm4(bhv, b_hi_addr + off);
                                    mm_ldsm4(blv, b_lo_addr + off);
                                } else {            // last, unpaired n-tile: lanes 16-31 supply (ignored) in-range addresses
                                    mm_ldsm2(bhv[0], bhv[1], b_hi_addr + off - (uint32_t)((lmat >> 1) * 8 * pitch) * 2);
                                    mm_ldsm2(blv[0], blv[1], b_lo_addr + off - (uint32_t)((lmat >> 1) * 8 * pitch) * 2);
                                    bhv[2] = bhv[3] = blv[2] = blv[3] = 0u;
                                }
                                mm_mma(acc[n], ah, blv[0], blv[1]);
                                mm_mma(acc[n], al, bhv[0], bhv[1]);
                                mm_mma(acc[n], ah, bhv[0], bhv[1]);
                                if (n + 1 < MM_NT) {
                                    mm_mma(acc[n + 1], ah, blv[2], blv[3]);
                                    mm_mma(acc[n + 1], al, bhv[2], bhv[3]);
                                    mm_mma(acc[n + 1], ah, bhv[2], bhv[3]);
                                }
                            }
                        }
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&hempty[2 * grp + h]);
                    }
                    ph_f ^= 1;
#pragma unroll
                    for (int n = 0; n < MM_NT; ++n) {
                        *reinterpret_cast<float2 *>(pmine + fg * MM_N + 8 * n + 2 * ft) = make_float2(acc[n][0], acc[n][1]);
                        *reinterpret_cast<float2 *>(pmine + (fg + 8) * MM_N + 8 * n + 2 * ft) = make_float2(acc[n][2], acc[n][3]);
                    }
                    asm volatile("bar.sync %0, %1;" ::"r"(1 + 2 * grp), "r"(32 * MM_NT) : "memory");
                    float4 sum = *reinterpret_cast<const float4 *>(pgrp + crow * MM_N + 4 * cj);
#pragma unroll
                    for (int w = 1; w < MM_NT; ++w) {
                        const float4 v = *reinterpret_cast<const float4 *>(pgrp + w * (MM_BT * MM_N) + crow * MM_N + 4 * cj);
                        sum.x += v.x; sum.y += v.y; sum.z += v.z; sum.w += v.w;
                    }
                    gt[0] = sum.x; gt[1] = sum.y; gt[2] = sum.z; gt[3] = sum.w;
                    asm volatile("bar.sync %0, %1;" ::"r"(2 + 2 * grp), "r"(32 * MM_NT) : "memory");
                }

                float *st = cstate + ((size_t)tau * MM_BT + crow) * MM_HS + cj;
                float hnew, gv[G], aux;
                if constexpr (CELL == DL4SS_CELL_LSTM) {
                    const float ig = sigmoid_f(xv[0] + gt[0]);
                    const float fgt = sigmoid_f(xv[1] + gt[1]);
                    const float gg = tanh_f(xv[2] + gt[2]);
                    const float og = sigmoid_f(xv[3] + gt[3]);
                    const float cc = fmaf(fgt, *st, ig * gg);
                    *st = cc;
                    hnew = og * tanh_f(cc);
                    gv[0] = ig; gv[1] = fgt; gv[2] = gg; gv[3] = og;
                    aux = cc;
                } else {
                    const float rg = sigmoid_f(xv[0] + gt[0]);
                    const float zg = sigmoid_f(xv[1] + gt[1]);
                    const float hn = gt[2] + bhn;
                    const float ng = tanh_f(fmaf(rg, hn, xv[2]));
                    hnew = fmaf(zg, *st - ng, ng);
                    *st = hnew;
                    gv[0] = rg; gv[1] = zg; gv[2] = ng;
                    aux = hn;
                }
                if (s + 1 < T) {
                    if (valid) {
                        const __nv_bfloat16 hi = __float2bfloat16_rn(hnew);
                        __nv_bfloat16 *dst = p.hx + ((size_t)(((s & 1) * 2 + dir) * 2)) * hx_plane + (size_t)b * Kg + u;
                        dst[0] = hi;
                        dst[hx_plane] = __float2bfloat16_rn(hnew - __bfloat162float(hi));
                    }
                    __syncwarp();
                    if (lane == 0) {
                        if (nrel > 0) { mbar_wait(&reldone[grp], ph_rd); ph_rd ^= 1; }
                        mbar_arrive(&rel[grp]);
                    }
                    ++nrel;
                }
                if (valid) {
                    p.y[((size_t)b * T + t) * 2 * H + (size_t)dir * H + u] = hnew;
                    if (p.gates_save != nullptr) {
#pragma unroll
                        for (int g = 0; g < G; ++g) p.gates_save[xrow * GH + (size_t)g * H + u] = gv[g];
                    }
                    if (p.cell_save != nullptr) p.cell_save[xrow * H + u] = aux;
                }
            }
        }
    }
}

static int mma_kg(int H) { return (H + 7) / 8 * 8; }
static int mma_pitch(int H) {                       // bf16 elements; row pitch == 4 (mod 32) 32-bit words
    int w = cdiv(H, 16) * 8;
    w += (4 - w % 32 + 32) % 32;
    return 2 * w;
}
static size_t mma_smem(int H) {
    return (size_t)(2 * MM_N + MM_GROUPS * 2 * MM_BT) * mma_pitch(H) * sizeof(__nv_bfloat16) +
           (size_t)(MM_MAX_TILES * MM_BT * MM_HS + MM_GROUPS * MM_NT * MM_BT * MM_N) * sizeof(float) + 6 * MM_GROUPS * sizeof(uint64_t);
}
static bool rnn_mma_supported(int H, int cell) {
    if (!(cell == DL4SS_CELL_LSTM || cell == DL4SS_CELL_GRU)) return false;
    if (H < MM_HS || H % MM_HS != 0) return false;
    if (2 * (H / MM_HS) > sm_count()) return false;          // every slice of both directions must be co-resident
    return mma_smem(H) <= 227 * 1024;
}

}  // namespace dl4ss

using namespace dl4ss;

extern "C" int dl4ss_rnn_mma_supported(int H, int cell) { return rnn_mma_supported(H, cell) ? 1 : 0; }

extern "C" size_t dl4ss_rnn_mma_workspace_bytes(int B, int T, int H, int cell) {
    (void)T; (void)cell;
    if (B <= 0 || H <= 0) return 256;
    const size_t ctr = (size_t)2 * cdiv(B, MM_BT) * MM_CTR_STRIDE * sizeof(unsigned);
    return ctr + (size_t)8 * B * mma_kg(H) * sizeof(__nv_bfloat16);
}

extern "C" int dl4ss_rnn_layer_mma_fwd(int cell, const float *xproj, const float *whh, const float *bhn, float *y,
                                       int B, int T, int H, float *gates_save, float *cell_save, void *workspace,
                                       size_t workspace_bytes, void *stream) {
    DL4SS_CHECK_ARG(cell == DL4SS_CELL_LSTM || cell == DL4SS_CELL_GRU, "rnn_layer_mma_fwd: bad cell %d", cell);
    DL4SS_CHECK_ARG(xproj && whh && y, "rnn_layer_mma_fwd: null operand");
    DL4SS_CHECK_ARG(cell == DL4SS_CELL_LSTM || bhn, "rnn_layer_mma_fwd: GRU needs bhn");
    DL4SS_CHECK_ARG(B >= 0 && T >= 1 && H >= 1, "rnn_layer_mma_fwd: bad B/T/H %d/%d/%d", B, T, H);
    if (!rnn_mma_supported(H, cell)) {
        set_error("rnn_layer_mma_fwd: H=%d unsupported (needs a multiple of %d, at most %d, whose W_hh slice fits shared "
                  "memory); use dl4ss_rnn_layer_fwd", H, MM_HS, MM_HS * (sm_count() / 2));
        return DL4SS_EUNSUPPORTED;
    }
    if (B == 0) return DL4SS_OK;
    const size_t need = dl4ss_rnn_mma_workspace_bytes(B, T, H, cell);
    if (!workspace || workspace_bytes < need) {
        set_error("rnn_layer_mma_fwd: workspace %zu B < %zu B", workspace_bytes, need);
        return DL4SS_EWORKSPACE;
    }
    DL4SS_CHECK_ARG((((uintptr_t)workspace) & 255) == 0, "rnn_layer_mma_fwd: workspace must be 256-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    DL4SS_CUDA(cudaMemsetAsync(workspace, 0, need, st));     // counters, and the zero k-padding of the exchange rows
    RnnMmaParams p;
    p.xproj = xproj; p.whh = whh; p.bhn = bhn; p.y = y; p.gates_save = gates_save; p.cell_save = cell_save;
    p.B = B; p.T = T; p.H = H;
    p.Kg = mma_kg(H); p.pitch = mma_pitch(H); p.nks = cdiv(H, 16);
    p.nslices = H / MM_HS;
    const size_t ctr = (size_t)2 * cdiv(B, MM_BT) * MM_CTR_STRIDE * sizeof(unsigned);
    p.hx = (__nv_bfloat16 *)((unsigned char *)workspace + ctr);
    const size_t smem = mma_smem(H);
    unsigned *cp = (unsigned *)workspace;
    for (int b0 = 0; b0 < B; b0 += MM_MAX_TILES * MM_BT) {
        p.b_begin = b0;
        p.ntiles = cdiv(((B - b0) < MM_MAX_TILES * MM_BT ? (B - b0) : MM_MAX_TILES * MM_BT), MM_BT);
        p.counters = cp;
        cp += (size_t)2 * p.ntiles * MM_CTR_STRIDE;
        void *args[] = {(void *)&p};
        if (cell == DL4SS_CELL_LSTM) {
            DL4SS_CUDA(cudaFuncSetAttribute(rnn_mma_kernel<DL4SS_CELL_LSTM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            DL4SS_CUDA(cudaLaunchCooperativeKernel((void *)rnn_mma_kernel<DL4SS_CELL_LSTM>, dim3(2 * p.nslices), dim3(MM_THREADS), args, smem, st));
        } else {
            DL4SS_CUDA(cudaFuncSetAttribute(rnn_mma_kernel<DL4SS_CELL_GRU>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            DL4SS_CUDA(cudaLaunchCooperativeKernel((void *)rnn_mma_kernel<DL4SS_CELL_GRU>, dim3(2 * p.nslices), dim3(MM_THREADS), args, smem, st));
        }
        count_launch();
    }
    return DL4SS_OK;
}
