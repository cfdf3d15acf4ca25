// K3  dl4ss_rnn_layer_fwd : one bidirectional LSTM / GRU layer as ONE persistent kernel.
//
// The reference runs nn.LSTM / nn.GRU through cuDNN (TDAA_beta/main_run_sstune_EvalVer.py:282-293):
// T sequential [B,H]x[H,G*H] products plus gate math per direction.  Here the input projection
// x*W_ih^T (+biases) is hoisted into one GEMM per layer (xproj), and this kernel walks the T steps
// without returning to the host:
//   * a CTA owns (direction, batch tile, slice of HS hidden units).  Its G*HS rows of W_hh stay in
//     shared memory for all T steps; cell state c (LSTM) / h (GRU) of its cells stays in registers;
//   * per step it pulls h_{t-1} of its batch tile (written to y by the sibling slices) from L2 with
//     cp.async.cg, accumulates the recurrent product, applies the gates and writes h_t into y;
//   * siblings of one (direction, batch tile) group synchronise through one global counter per
//     group (release/acquire), so the two directions and the batch tiles never wait on each other;
//   * the xproj rows of step t+1 are prefetched with cp.async while step t computes.
// The launch is cooperative (all CTAs co-resident) so the counter waits cannot deadlock.
#include "common.cuh"
#include <cooperative_groups.h>

namespace dl4ss {

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async8(void *smem, const void *gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async4(void *smem, const void *gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

__device__ __forceinline__ unsigned ld_acquire(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_release_add(unsigned *p, unsigned v) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;\n" ::"l"(p), "r"(v) : "memory");
}

struct RnnParams {
    const float *xproj;   // [B,T,2,G*H]
    const float *whh;     // [2,G*H,H]
    const float *bhn;     // [2,H] (GRU) or null
    float *y;             // [B,T,2H]
    float *gates_save;    // [B,T,2,G*H] or null
    float *cell_save;     // [B,T,2,H] or null
    unsigned *counters;   // [2 * batch_tiles]
    int B, T, H, HP, KQ;  // HP: smem row pitch (floats), KQ: float4 quads per row (multiple of KS)
    int b_begin;          // first utterance of this launch (batch chunking)
    int batch_tiles, nslices;
};

// CELL: 0 LSTM (G=4), 1 GRU (G=3).  NW warps, U units per warp, R rows per lane, RB lanes span rows
// (KS = 32/RB lanes split the K loop).
template <int CELL, int NW, int U, int R, int RB>
__global__ void __launch_bounds__(32 * NW, 1)
rnn_layer_kernel(const RnnParams p) {
    constexpr int G = (CELL == DL4SS_CELL_LSTM) ? 4 : 3;
    constexpr int HS = NW * U;
    constexpr int BT = RB * R;
    constexpr int KS = 32 / RB;
    constexpr int NT = 32 * NW;
    extern __shared__ __align__(16) float smem[];
    const int HP = p.HP, H = p.H, T = p.T;
    float *Ws = smem;                               // [G*HS][HP]
    float *hs = Ws + (size_t)G * HS * HP;           // [BT][HP]
    float *xs = hs + (size_t)BT * HP;               // [2][BT][G*HS]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int rl = lane % RB, ksub = lane / RB;
    int bid = blockIdx.x;
    const int slice = bid % p.nslices; bid /= p.nslices;
    const int bt = bid % p.batch_tiles;
    const int dir = bid / p.batch_tiles;
    const int row0 = p.b_begin + bt * BT;           // first utterance of this tile
    const int u0 = slice * HS;                      // first hidden unit of this slice
    unsigned *counter = p.counters + dir * p.batch_tiles + bt;
    const size_t GH = (size_t)G * H;

    // ---- resident W_hh slice (zero padded to HP) and zeroed h tile
    {
        const float *wsrc = p.whh + (size_t)dir * GH * H;
        for (int i = tid; i < G * HS * HP; i += NT) {
            const int lr = i / HP, k = i - lr * HP;
            const int g = lr / HS, u = lr - g * HS;
            Ws[i] = (k < H) ? wsrc[((size_t)g * H + u0 + u) * H + k] : 0.f;
        }
        for (int i = tid; i < BT * HP; i += NT) hs[i] = 0.f;
    }

    auto prefetch_x = [&](int step, int buf) {
        const int t = dir ? (T - 1 - step) : step;
        float *dst = xs + (size_t)buf * BT * G * HS;
        if ((HS & 3) == 0 && (H & 3) == 0) {
            constexpr int V = HS / 4 > 0 ? HS / 4 : 1;
            for (int i = tid; i < BT * G * V; i += NT) {
                const int v = i % V, g = (i / V) % G, r = i / (V * G);
                const int b = row0 + r;
                if (b < p.B)
                    cp_async16(dst + (r * G + g) * HS + 4 * v,
                               p.xproj + (((size_t)b * T + t) * 2 + dir) * GH + (size_t)g * H + u0 + 4 * v);
            }
        } else if ((HS & 1) == 0 && (H & 1) == 0) {
            constexpr int V = HS / 2;
            for (int i = tid; i < BT * G * V; i += NT) {
                const int v = i % V, g = (i / V) % G, r = i / (V * G);
                const int b = row0 + r;
                if (b < p.B)
                    cp_async8(dst + (r * G + g) * HS + 2 * v,
                              p.xproj + (((size_t)b * T + t) * 2 + dir) * GH + (size_t)g * H + u0 + 2 * v);
            }
        } else {
            for (int i = tid; i < BT * G * HS; i += NT) {
                const int v = i % HS, g = (i / HS) % G, r = i / (HS * G);
                const int b = row0 + r;
                if (b < p.B)
                    cp_async4(dst + (r * G + g) * HS + v,
                              p.xproj + (((size_t)b * T + t) * 2 + dir) * GH + (size_t)g * H + u0 + v);
            }
        }
        cp_async_commit();
    };

    float state[R][U];   // LSTM: c ; GRU: h
#pragma unroll
    for (int i = 0; i < R; ++i)
#pragma unroll
        for (int u = 0; u < U; ++u) state[i][u] = 0.f;
    float bhn_r[U];
#pragma unroll
    for (int u = 0; u < U; ++u)
        bhn_r[u] = (CELL == DL4SS_CELL_GRU) ? p.bhn[(size_t)dir * H + u0 + warp * U + u] : 0.f;

    prefetch_x(0, 0);
    __syncthreads();

    for (int step = 0; step < T; ++step) {
        const int t = dir ? (T - 1 - step) : step;
        if (step + 1 < T) prefetch_x(step + 1, (step + 1) & 1);
        if (step > 0) {
            // wait until every slice of this (direction, batch tile) has published step-1
            if (tid == 0) {
                const unsigned want = (unsigned)p.nslices * (unsigned)step;
                while (ld_acquire(counter) < want) { __nanosleep(20); }
            }
            __syncthreads();
            const int tp = dir ? (t + 1) : (t - 1);
            if ((H & 3) == 0) {
                const int V = H >> 2;
                for (int i = tid; i < BT * V; i += NT) {
                    const int r = i / V, v = i - r * V;
                    const int b = row0 + r;
                    if (b < p.B)
                        cp_async16(hs + (size_t)r * HP + 4 * v,
                                   p.y + ((size_t)b * T + tp) * 2 * H + (size_t)dir * H + 4 * v);
                }
            } else {
                for (int i = tid; i < BT * H; i += NT) {
                    const int r = i / H, v = i - r * H;
                    const int b = row0 + r;
                    if (b < p.B)
                        hs[(size_t)r * HP + v] = __ldcg(p.y + ((size_t)b * T + tp) * 2 * H + (size_t)dir * H + v);
                }
            }
            cp_async_commit();
        }
        cp_async_wait_all();
        __syncthreads();

        float acc[R][U][G];
#pragma unroll
        for (int i = 0; i < R; ++i)
#pragma unroll
            for (int u = 0; u < U; ++u)
#pragma unroll
                for (int g = 0; g < G; ++g) acc[i][u][g] = 0.f;

        if (step > 0) {
            const float *hrow = hs + (size_t)rl * HP;
            const float *wrow = Ws + (size_t)(warp * U) * HP;
#pragma unroll 2
            for (int q = ksub; q < p.KQ; q += KS) {
                float4 hv[R];
#pragma unroll
                for (int i = 0; i < R; ++i)
                    hv[i] = *reinterpret_cast<const float4 *>(hrow + (size_t)(RB * i) * HP + 4 * q);
#pragma unroll
                for (int u = 0; u < U; ++u)
#pragma unroll
                    for (int g = 0; g < G; ++g) {
                        const float4 w = *reinterpret_cast<const float4 *>(wrow + (size_t)(g * HS + u) * HP + 4 * q);
#pragma unroll
                        for (int i = 0; i < R; ++i) {
                            float a = acc[i][u][g];
                            a = fmaf(hv[i].x, w.x, a);
                            a = fmaf(hv[i].y, w.y, a);
                            a = fmaf(hv[i].z, w.z, a);
                            a = fmaf(hv[i].w, w.w, a);
                            acc[i][u][g] = a;
                        }
                    }
            }
            if (KS > 1) {
#pragma unroll
                for (int i = 0; i < R; ++i)
#pragma unroll
                    for (int u = 0; u < U; ++u)
#pragma unroll
                        for (int g = 0; g < G; ++g) {
                            float a = acc[i][u][g];
#pragma unroll
                            for (int o = RB; o < 32; o <<= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
                            acc[i][u][g] = a;
                        }
            }
        }

        // ---- gates, state update, publish h_t
        const float *xb = xs + (size_t)(step & 1) * BT * G * HS;
#pragma unroll
        for (int i = 0; i < R; ++i) {
            const int r = rl + RB * i;
            const int b = row0 + r;
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int lu = warp * U + u;
                const float *xr = xb + (size_t)r * G * HS + lu;
                float hnew, gv[G];
                if constexpr (CELL == DL4SS_CELL_LSTM) {
                    const float ig = sigmoid_f(xr[0 * HS] + acc[i][u][0]);
                    const float fg = sigmoid_f(xr[1 * HS] + acc[i][u][1]);
                    const float gg = tanh_f(xr[2 * HS] + acc[i][u][2]);
                    const float og = sigmoid_f(xr[3 * HS] + acc[i][u][3]);
                    const float c = fmaf(fg, state[i][u], ig * gg);
                    state[i][u] = c;
                    hnew = og * tanh_f(c);
                    gv[0] = ig; gv[1] = fg; gv[2] = gg; gv[3] = og;
                } else {
                    const float rg = sigmoid_f(xr[0 * HS] + acc[i][u][0]);
                    const float zg = sigmoid_f(xr[1 * HS] + acc[i][u][1]);
                    const float hn = acc[i][u][2] + bhn_r[u];
                    const float ng = tanh_f(fmaf(rg, hn, xr[2 * HS]));
                    hnew = fmaf(zg, state[i][u] - ng, ng);      // (1-z)*n + z*h
                    state[i][u] = hnew;
                    gv[0] = rg; gv[1] = zg; gv[2] = ng;
                    if (p.cell_save != nullptr && ksub == 0 && b < p.B)     // GRU: keep W_hn*h + b_hn for backward
                        p.cell_save[(((size_t)b * T + t) * 2 + dir) * H + u0 + lu] = hn;
                }
                if (ksub == 0 && b < p.B) {
                    p.y[((size_t)b * T + t) * 2 * H + (size_t)dir * H + u0 + lu] = hnew;
                    if (p.gates_save != nullptr) {
#pragma unroll
                        for (int g = 0; g < G; ++g)
                            p.gates_save[(((size_t)b * T + t) * 2 + dir) * GH + (size_t)g * H + u0 + lu] = gv[g];
                    }
                    if (CELL == DL4SS_CELL_LSTM && p.cell_save != nullptr)
                        p.cell_save[(((size_t)b * T + t) * 2 + dir) * H + u0 + lu] = state[i][u];
                }
            }
        }
        __syncthreads();
        if (tid == 0) {
            __threadfence();
            red_release_add(counter, 1u);
        }
    }
}

struct RnnCfg { int NW, U, R, RB; };

template <int CELL, int NW, int U, int R, int RB>
static int launch_rnn(RnnParams p, int nchunk_rows, cudaStream_t st, int *launched_rows) {
    constexpr int G = (CELL == DL4SS_CELL_LSTM) ? 4 : 3;
    constexpr int HS = NW * U, BT = RB * R, KS = 32 / RB;
    if (p.H % HS != 0) {
        set_error("rnn_layer_fwd: H=%d is not a multiple of the %d-unit slice", p.H, HS);
        return DL4SS_EUNSUPPORTED;
    }
    int KQ = (p.H + 3) / 4;
    KQ = (KQ + KS - 1) / KS * KS;
    int HP = 4 * KQ;
    if (HP % 8 == 0) HP += 4;          // odd multiple of 4 floats: conflict-free 128-bit row reads
    p.KQ = KQ;
    p.HP = HP;
    p.nslices = p.H / HS;
    const size_t smem = ((size_t)G * HS * HP + (size_t)BT * HP + 2ull * BT * G * HS) * sizeof(float);
    if (smem > 227 * 1024) {
        set_error("rnn_layer_fwd: H=%d needs %zu B of shared memory per CTA", p.H, smem);
        return DL4SS_EUNSUPPORTED;
    }
    auto kern = rnn_layer_kernel<CELL, NW, U, R, RB>;
    DL4SS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    DL4SS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 32 * NW, smem));
    const int max_ctas = per_sm * sm_count();
    int max_tiles = max_ctas / (2 * p.nslices);
    if (max_tiles < 1) {
        set_error("rnn_layer_fwd: %d co-resident CTAs cannot hold one batch tile (%d slices x 2 directions)",
                  max_ctas, p.nslices);
        return DL4SS_EUNSUPPORTED;
    }
    int tiles = (nchunk_rows + BT - 1) / BT;
    if (tiles > max_tiles) tiles = max_tiles;
    p.batch_tiles = tiles;
    *launched_rows = tiles * BT;
    void *args[] = {(void *)&p};
    DL4SS_CUDA(cudaLaunchCooperativeKernel((void *)kern, dim3(2 * tiles * p.nslices), dim3(32 * NW), args, smem, st));
    count_launch();
    return DL4SS_OK;
}

template <int CELL>
static int rnn_dispatch(RnnParams p, int rows_left, cudaStream_t st, int *launched_rows) {
    // largest batch tile first; a configuration whose resident W_hh slice + h tile does not fit in shared memory
    // (H = 600: the speaker classifier) falls through to the next smaller one
    int rc = DL4SS_EUNSUPPORTED;
    if (rows_left > 32) rc = launch_rnn<CELL, 10, 2, 2, 32>(p, rows_left, st, launched_rows);
    if (rc == DL4SS_EUNSUPPORTED && rows_left > 8) rc = launch_rnn<CELL, 10, 2, 1, 32>(p, rows_left, st, launched_rows);
    // H = 600: a 10-unit slice (96 KB of W_hh) next to a 32-utterance h tile (77 KB) still fits; one unit per warp
    if (rc == DL4SS_EUNSUPPORTED && rows_left > 8) rc = launch_rnn<CELL, 10, 1, 1, 32>(p, rows_left, st, launched_rows);
    if (rc == DL4SS_EUNSUPPORTED) rc = launch_rnn<CELL, 5, 2, 1, 8>(p, rows_left, st, launched_rows);
    return rc;
}

}  // namespace dl4ss

using namespace dl4ss;

extern "C" size_t dl4ss_rnn_workspace_bytes(int B, int T, int H, int cell) {
    (void)T; (void)H; (void)cell;
    if (B <= 0) return 256;
    // one counter per (launch chunk, direction, batch tile); tiles are >= 8 rows
    return (size_t)(2 * ((B + 7) / 8) + 2) * sizeof(unsigned) + 256;
}

extern "C" int dl4ss_rnn_layer_fwd(int cell, const float *xproj, const float *whh, const float *bhn, float *y,
                                   int B, int T, int H, float *gates_save, float *cell_save, void *workspace,
                                   size_t workspace_bytes, void *stream) {
    DL4SS_CHECK_ARG(cell == DL4SS_CELL_LSTM || cell == DL4SS_CELL_GRU, "rnn_layer_fwd: bad cell %d", cell);
    DL4SS_CHECK_ARG(xproj && whh && y, "rnn_layer_fwd: null operand");
    DL4SS_CHECK_ARG(cell == DL4SS_CELL_LSTM || bhn, "rnn_layer_fwd: GRU needs bhn");
    DL4SS_CHECK_ARG(B >= 0 && T >= 1 && H >= 1, "rnn_layer_fwd: bad B/T/H %d/%d/%d", B, T, H);
    if (B == 0) return DL4SS_OK;
    const size_t need = dl4ss_rnn_workspace_bytes(B, T, H, cell);
    if (!workspace || workspace_bytes < need) {
        set_error("rnn_layer_fwd: workspace %zu B < %zu B", workspace_bytes, need);
        return DL4SS_EWORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    DL4SS_CUDA(cudaMemsetAsync(workspace, 0, need, st));
    RnnParams p;
    p.xproj = xproj; p.whh = whh; p.bhn = bhn; p.y = y;
    p.gates_save = gates_save; p.cell_save = cell_save;
    p.B = B; p.T = T; p.H = H; p.HP = 0; p.KQ = 0; p.batch_tiles = 0; p.nslices = 0;
    unsigned *ctr = (unsigned *)workspace;
    int b0 = 0;
    while (b0 < B) {
        p.b_begin = b0;
        p.counters = ctr;
        int done = 0;
        int rc = (cell == DL4SS_CELL_LSTM) ? rnn_dispatch<DL4SS_CELL_LSTM>(p, B - b0, st, &done)
                                           : rnn_dispatch<DL4SS_CELL_GRU>(p, B - b0, st, &done);
        if (rc) return rc;
        ctr += 2 * ((done + 7) / 8);
        b0 += done;
    }
    return DL4SS_OK;
}
