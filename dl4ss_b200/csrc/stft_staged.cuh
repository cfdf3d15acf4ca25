// Persistent, bulk-copy staged forms of K1 / K6 for hop = n_fft/2 (every reference config).  Included by stft.cu.
//
// The first-round kernels pulled their rows straight from global memory into registers, so the bytes a SM kept in
// flight were tied to how many warps sat in their load phase (16 warps / SM at 126 registers: ncu showed issue slots
// 36 % and L1 64 % busy -- latency bound, 61 % / 68 % of the HBM copy peak at the bench size).  Here one producer warp
// per CTA streams the rows of the NEXT work items into a shared-memory ring with cp.async.bulk (UBLKCP: the copy engine
// moves 16-64 KB per item with no registers and no scoreboard slots), completion is signalled on mbarriers, and the FFT
// groups only ever read shared memory.  CTAs are persistent (one resident set, items taken round robin), so the twiddle
// / window tables are built once per CTA and the last wave is an item, not a CTA, long.
//
// Rows of the [.,T,129] tensors are 516 / 1032 bytes: not multiples of 16, so an item's span starts and ends at
// arbitrary 4-byte (8-byte) offsets.  A span is fetched as its 16-byte aligned superset -- the payload lands at
// dst + (address & 15) -- clipped to the last whole 16 bytes of the tensor; the <= 12 bytes that can remain at the very
// end of a tensor are copied by the producer warp with ordinary loads before it signals the barrier.
#pragma once

namespace dl4ss {

__device__ __forceinline__ uint32_t sts_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void sbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(sts_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void sbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(sts_u32(bar)) : "memory");
}
__device__ __forceinline__ void sbar_arrive_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sts_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void sbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "SWAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra SWAIT_DONE;\n\t"
        "bra SWAIT_LOOP;\n\t"
        "SWAIT_DONE:\n\t"
        "}\n" ::"r"(sts_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ uint64_t evict_first_policy() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;\n" : "=l"(pol));
    return pol;
}
// global -> shared bulk copy (both 16-byte aligned, bytes a multiple of 16), completion counted on `bar`
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar, uint64_t pol) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(sts_u32(dst)), "l"(src), "r"(bytes), "r"(sts_u32(bar)), "l"(pol) : "memory");
}

// One span of a staged item: bytes [p, p+n) of a tensor that ends at `tend`; the payload lands at dst + (p & 15).
struct Span {
    const char *lo;        // 16-byte aligned start of the bulk part
    uint32_t bulk;         // bytes of the bulk part (multiple of 16, may be 0)
    const char *tail;      // first byte not covered by the bulk part
    uint32_t ntail;        // payload bytes after the bulk part (multiple of 4, < 16 in practice)
};
__device__ __forceinline__ Span make_span(const char *p, size_t n, const char *tend) {
    Span s;
    s.lo = reinterpret_cast<const char *>(reinterpret_cast<uintptr_t>(p) & ~(uintptr_t)15);
    const char *hi = reinterpret_cast<const char *>((reinterpret_cast<uintptr_t>(p + n) + 15) & ~(uintptr_t)15);
    const char *lim = reinterpret_cast<const char *>(reinterpret_cast<uintptr_t>(tend) & ~(uintptr_t)15);
    if (hi > lim) hi = (lim > s.lo) ? lim : s.lo;
    s.bulk = (uint32_t)(hi - s.lo);
    s.tail = hi;
    s.ntail = (p + n > hi) ? (uint32_t)(p + n - hi) : 0u;
    return s;
}
// producer warp: tail words with ordinary loads (all lanes), then lane 0 issues the bulk part
__device__ __forceinline__ void span_tail(const Span &s, char *dst, int lane) {
    if (s.ntail) {
        char *d = dst + (s.tail - s.lo);
        for (uint32_t i = 4u * lane; i < s.ntail; i += 128u)
            *reinterpret_cast<uint32_t *>(d + i) = *reinterpret_cast<const uint32_t *>(s.tail + i);
    }
}

// ------------------------------------------------------------------------------------ K6 staged
// Throughput of this kernel = (frame-pair tasks in flight per SM) / (latency of one task: ~600 dependent-ish instructions):
// the first staged form (16 groups, two 64 KB stages) ran 8 consumer warps per SM and reached 76 % of the copy peak at
// large batch.  24 groups (12 consumer warps, the register file's limit at 137 registers) need the shared memory of the
// second stage: ONE stage is enough because the consumers release it as soon as their rows are in registers, so the
// next item's rows stream in underneath the transforms and stores of the current one; the parked half frames live in
// the owning group's (by then idle) transpose buffer.
template <int MASK_KIND>
struct K6Cfg {
    // real masks / per-source spectra: 24 groups; complex masks (two floats per bin and source: 146 KB of rows for 47 frames)
    // fit with 16 groups
    static constexpr int GROUPS = (MASK_KIND == DL4SS_MASK_COMPLEX) ? 16 : 24;
    static constexpr int WARPS = GROUPS / 2;                 // consumer warps
    static constexpr int PAIRS = GROUPS - 1;                 // pairs a tile owns (the last group is the halo pair)
    static constexpr int ROWS = 2 * PAIRS + 1;               // frames staged per tile (+ the halo frame)
    static constexpr int CONSUMERS = 16 * GROUPS;
    static constexpr int THREADS = CONSUMERS + 32;           // + the producer warp
    static constexpr int XBYTES = (ROWS * NBIN * 8 + 16 + 127) & ~127;     // complex rows (+ alignment slack)
    static constexpr int MBYTES = (ROWS * NBIN * 4 + 16 + 127) & ~127;     // real rows
    // REAL: mixture rows + two real mask planes ; COMPLEX: mixture rows + two complex mask planes ; NONE: two per-source spectra
    static constexpr int A = XBYTES;                                                         // mixture / source 0
    static constexpr int B = (MASK_KIND == DL4SS_MASK_REAL) ? MBYTES : XBYTES;               // mask 0 / source 1
    static constexpr int C = (MASK_KIND == DL4SS_MASK_REAL) ? MBYTES : (MASK_KIND == DL4SS_MASK_COMPLEX ? XBYTES : 0);   // mask 1
    static constexpr int BYTES = A + B + C;
    static constexpr size_t SMEM = (size_t)BYTES + (size_t)GROUPS * DL4SS_XCH2_FLOAT4 * sizeof(float4) + 256 * sizeof(float2) +
                                   NFFT * sizeof(float) + (2 + 2 * WARPS) * sizeof(uint64_t) + 128;
};

template <int MASK_KIND>
__global__ void __launch_bounds__(K6Cfg<MASK_KIND>::THREADS, 1)
istft_h128_staged_kernel(const float *__restrict__ mask, const float2 *__restrict__ spec, int S, int T,
                         int tiles_per_src, int n_items, const char *mask_end, const char *spec_end,
                         const float *__restrict__ window, float *__restrict__ out) {
    using C = K6Cfg<MASK_KIND>;
    using St = K6Cfg<MASK_KIND>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned char *stage0 = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);   // [STAGES][St::BYTES]
    float4 *xch = reinterpret_cast<float4 *>(stage0 + St::BYTES);               // groups * 272
    float2 *tw = reinterpret_cast<float2 *>(xch + C::GROUPS * DL4SS_XCH2_FLOAT4);           // 256
    float *wlo = reinterpret_cast<float *>(tw + 256);                                        // 128
    float *whi = wlo + NFFT / 2;                                                             // 128
    uint64_t *full = reinterpret_cast<uint64_t *>(whi + NFFT / 2);                           // [STAGES]
    uint64_t *empty = full + 1;                                                     // [STAGES]
    uint64_t *xfull = empty + 1;                                                    // [consumer warps]
    uint64_t *xempty = xfull + C::WARPS;                                                    // [consumer warps]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int SP = (S + 1) >> 1;
    const int Lout = (NFFT / 2) * (T - 1);

    if (tid == 0) {
        sbar_init(&full[0], 1);
        sbar_init(&empty[0], C::CONSUMERS / 32);
        for (int i = 0; i < C::WARPS; ++i) { sbar_init(&xfull[i], 32); sbar_init(&xempty[i], 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();                                         // barriers initialised: the producer starts streaming at once
    if (tid < 256) tw[tid] = g_tw256[tid];
    if (tid < NFFT / 2) {
        // out[j] = (frame_hi[j+128]*w[j+128] + frame_lo[j]*w[j]) / (w[j]^2 + w[j+128]^2), 1/N of the inverse folded in
        const float w1 = window[tid], w2 = window[tid + NFFT / 2];
        const float e = w1 * w1 + w2 * w2;
        const float inv = ((e > 1.17549435e-38f) ? 1.0f / e : 1.0f) * (1.0f / NFFT);
        wlo[tid] = w1 * inv;
        whi[tid] = w2 * inv;
    }
    if (warp < C::CONSUMERS / 32) asm volatile("bar.sync 1, %0;" ::"n"(C::CONSUMERS) : "memory");     // tables: consumers only

    if (warp == C::CONSUMERS / 32) {
        // ================= producer warp: rows of item k into stage k % STAGES
        const uint64_t pol = evict_first_policy();
        int k = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++k) {
            const int st = 0;
            if (k >= 1) sbar_wait(&empty[st], (uint32_t)(k - 1) & 1u);
            // source pair fastest, then tile, then utterance: the CTAs resident at one time stream through one contiguous
            // region (a tile-major order, shortest items last, measured 1.5 % slower)
            int id = item;
            const int sp = id % SP; id /= SP;
            const int tile = id % tiles_per_src;
            const int b = id / tiles_per_src;
            const int s0 = 2 * sp, s1 = min(s0 + 1, S - 1);
            const int t0 = tile * 2 * C::PAIRS;
            const int nrows = min(C::ROWS, T - t0);
            unsigned char *sb = stage0 + (size_t)st * St::BYTES;
            Span a, bq, c;
            c.bulk = 0; c.ntail = 0; c.lo = c.tail = nullptr;
            if (MASK_KIND != DL4SS_MASK_NONE) {
                constexpr size_t MB = (MASK_KIND == DL4SS_MASK_COMPLEX) ? 8 : 4;      // bytes per mask bin
                const char *mbase = reinterpret_cast<const char *>(mask);
                a = make_span(reinterpret_cast<const char *>(spec + ((size_t)b * T + t0) * NBIN), (size_t)nrows * NBIN * 8, spec_end);
                bq = make_span(mbase + (((size_t)b * S + s0) * T + t0) * NBIN * MB, (size_t)nrows * NBIN * MB, mask_end);
                if (s1 != s0)
                    c = make_span(mbase + (((size_t)b * S + s1) * T + t0) * NBIN * MB, (size_t)nrows * NBIN * MB, mask_end);
            } else {
                a = make_span(reinterpret_cast<const char *>(spec + (((size_t)b * S + s0) * T + t0) * NBIN), (size_t)nrows * NBIN * 8, spec_end);
                bq.bulk = 0; bq.ntail = 0; bq.lo = bq.tail = nullptr;
                if (s1 != s0)
                    bq = make_span(reinterpret_cast<const char *>(spec + (((size_t)b * S + s1) * T + t0) * NBIN), (size_t)nrows * NBIN * 8, spec_end);
            }
            span_tail(a, reinterpret_cast<char *>(sb), lane);
            span_tail(bq, reinterpret_cast<char *>(sb + St::A), lane);
            span_tail(c, reinterpret_cast<char *>(sb + St::A + St::B), lane);
            __threadfence_block();
            __syncwarp();
            if (lane == 0) {
                sbar_arrive_tx(&full[st], a.bulk + bq.bulk + c.bulk);
                if (a.bulk) bulk_g2s(sb, a.lo, a.bulk, &full[st], pol);
                if (bq.bulk) bulk_g2s(sb + St::A, bq.lo, bq.bulk, &full[st], pol);
                if (c.bulk) bulk_g2s(sb + St::A + St::B, c.lo, c.bulk, &full[st], pol);
            }
            __syncwarp();
        }
        return;
    }

    // ================= consumers: group g owns frame pair (2g, 2g+1) of the tile; the last group is the halo
    const int g = tid >> 4, l16 = tid & 15;
    // parked half frame of group g: the first 1 KB of its own transpose buffer (idle between two transforms)
    float2 *ex = reinterpret_cast<float2 *>(xch + g * DL4SS_XCH2_FLOAT4);
    int k = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++k) {
        const int st = 0;
        int id = item;
        const int sp = id % SP; id /= SP;
        const int tile = id % tiles_per_src;
        const int b = id / tiles_per_src;
        const int s0 = 2 * sp, s1 = min(s0 + 1, S - 1);
        const bool two = (s0 + 1 < S);
        const int t0 = tile * 2 * C::PAIRS;
        const int nrows = min(C::ROWS, T - t0);
        const int ta = t0 + 2 * g, tb = ta + 1;
        // out-of-range frames are clamped: read like any other, never stored
        const int ra = min(2 * g, nrows - 1);
        const int rb = (tb < T && g < C::PAIRS) ? 2 * g + 1 : ra;      // the halo group only needs its first frame
        const bool warp_live = t0 + 2 * (g & ~1) < T;

        sbar_wait(&full[st], (uint32_t)k & 1u);
        const unsigned char *sb = stage0 + (size_t)st * St::BYTES;
        // this warp's first transpose buffer still holds the half frame it parked for the warp before: that warp must
        // have taken the previous item's values before the next transform overwrites them
        if (warp > 0 && k > 0) sbar_wait(&xempty[warp], (uint32_t)(k - 1) & 1u);
        cx2 v[16];
        if (warp_live) {
            cx2 pa[8], pb[8];
            float2 pa_n, pb_n;
            if (MASK_KIND == DL4SS_MASK_NONE) {
                const size_t o0 = ((size_t)b * S + s0) * T + t0, o1 = ((size_t)b * S + s1) * T + t0;
                const float2 *x0 = reinterpret_cast<const float2 *>(sb + ((o0 * NBIN * 8) & 15));
                const float2 *x1 = two ? reinterpret_cast<const float2 *>(sb + St::A + ((o1 * NBIN * 8) & 15)) : x0;
                const float2 *ra0 = x0 + ra * NBIN, *ra1 = x1 + ra * NBIN, *rb0 = x0 + rb * NBIN, *rb1 = x1 + rb * NBIN;
#pragma unroll
                for (int m = 0; m < 8; ++m) {
                    const float2 a0 = ra0[16 * m + l16], a1 = ra1[16 * m + l16], b0 = rb0[16 * m + l16], b1 = rb1[16 * m + l16];
                    pa[m] = cx2{make_float2(a0.x, a1.x), make_float2(a0.y, a1.y)};
                    pb[m] = cx2{make_float2(b0.x, b1.x), make_float2(b0.y, b1.y)};
                }
                pa_n = pb_n = make_float2(0.f, 0.f);
                if (l16 == 0) {
                    pa_n = make_float2(ra0[128].x, ra1[128].x);
                    pb_n = make_float2(rb0[128].x, rb1[128].x);
                }
            } else if (MASK_KIND == DL4SS_MASK_COMPLEX) {
                const size_t ox = (size_t)b * T + t0;
                const size_t om0 = ((size_t)b * S + s0) * T + t0, om1 = ((size_t)b * S + s1) * T + t0;
                const float2 *xs = reinterpret_cast<const float2 *>(sb + ((ox * NBIN * 8) & 15));
                const float2 *m0 = reinterpret_cast<const float2 *>(sb + St::A + ((om0 * NBIN * 8) & 15));
                const float2 *m1 = two ? reinterpret_cast<const float2 *>(sb + St::A + St::B + ((om1 * NBIN * 8) & 15)) : m0;
                const float2 *xa = xs + ra * NBIN, *xb = xs + rb * NBIN;
                const float2 *ma0 = m0 + ra * NBIN, *ma1 = m1 + ra * NBIN, *mb0 = m0 + rb * NBIN, *mb1 = m1 + rb * NBIN;
#pragma unroll
                for (int m = 0; m < 8; ++m) {
                    const float2 xva = xa[16 * m + l16], xvb = xb[16 * m + l16];
                    const float2 a0 = ma0[16 * m + l16], a1 = ma1[16 * m + l16];
                    const float2 b0 = mb0[16 * m + l16], b1 = mb1[16 * m + l16];
                    const float2 kar = make_float2(a0.x, a1.x), kai = make_float2(a0.y, a1.y);
                    const float2 kbr = make_float2(b0.x, b1.x), kbi = make_float2(b0.y, b1.y);
                    // reference order: re = Mr*Xr - Mi*Xi ; im = Mr*Xi + Mi*Xr
                    pa[m] = cx2{pfnma(kai, pbc(xva.y), pmul(kar, pbc(xva.x))), pfma(kai, pbc(xva.x), pmul(kar, pbc(xva.y)))};
                    pb[m] = cx2{pfnma(kbi, pbc(xvb.y), pmul(kbr, pbc(xvb.x))), pfma(kbi, pbc(xvb.x), pmul(kbr, pbc(xvb.y)))};
                }
                pa_n = pb_n = make_float2(0.f, 0.f);
                if (l16 == 0) {
                    const float2 xan = xa[128], xbn = xb[128];
                    const float2 a0 = ma0[128], a1 = ma1[128], b0 = mb0[128], b1 = mb1[128];
                    pa_n = make_float2(a0.x * xan.x - a0.y * xan.y, a1.x * xan.x - a1.y * xan.y);
                    pb_n = make_float2(b0.x * xbn.x - b0.y * xbn.y, b1.x * xbn.x - b1.y * xbn.y);
                }
            } else {
                const size_t ox = (size_t)b * T + t0;
                const size_t om0 = ((size_t)b * S + s0) * T + t0, om1 = ((size_t)b * S + s1) * T + t0;
                const float2 *xs = reinterpret_cast<const float2 *>(sb + ((ox * NBIN * 8) & 15));
                const float *m0 = reinterpret_cast<const float *>(sb + St::A + ((om0 * NBIN * 4) & 15));
                const float *m1 = two ? reinterpret_cast<const float *>(sb + St::A + St::B + ((om1 * NBIN * 4) & 15)) : m0;
                const float2 *xa = xs + ra * NBIN, *xb = xs + rb * NBIN;
                const float *ma0 = m0 + ra * NBIN, *ma1 = m1 + ra * NBIN, *mb0 = m0 + rb * NBIN, *mb1 = m1 + rb * NBIN;
#pragma unroll
                for (int m = 0; m < 8; ++m) {
                    const float2 xva = xa[16 * m + l16], xvb = xb[16 * m + l16];
                    const float2 ka = make_float2(ma0[16 * m + l16], ma1[16 * m + l16]);
                    const float2 kb = make_float2(mb0[16 * m + l16], mb1[16 * m + l16]);
                    pa[m] = cx2{pmul(ka, pbc(xva.x)), pmul(ka, pbc(xva.y))};
                    pb[m] = cx2{pmul(kb, pbc(xvb.x)), pmul(kb, pbc(xvb.y))};
                }
                pa_n = pb_n = make_float2(0.f, 0.f);
                if (l16 == 0) {
                    const float xan = xa[128].x, xbn = xb[128].x;
                    pa_n = make_float2(ma0[128] * xan, ma1[128] * xan);
                    pb_n = make_float2(mb0[128] * xbn, mb1[128] * xbn);
                }
            }
            __syncwarp();
            if (lane == 0) sbar_arrive(&empty[st]);          // the rows are in registers: the stage may be refilled
            if (l16 == 0) {   // DC bin: irfft ignores the imaginary part
                pa[0].im = make_float2(0.f, 0.f);
                pb[0].im = make_float2(0.f, 0.f);
            }
            // Z[k] = A[k] + i*B[k] for k <= 128 ; Z[256-k] = conj(A[k]) + i*conj(B[k])
            const int src = (16 - l16) & 15;
            cx2 c[8];
#pragma unroll
            for (int m = 0; m < 8; ++m) {
                v[m] = cx2{psub(pa[m].re, pb[m].im), padd(pa[m].im, pb[m].re)};
                c[m] = cx2{padd(pa[m].re, pb[m].im), psub(pb[m].re, pa[m].im)};
            }
#pragma unroll
            for (int m = 8; m < 16; ++m) {
                cx2 t;
                t.re.x = __shfl_sync(0xffffffffu, c[15 - m].re.x, src, 16);
                t.re.y = __shfl_sync(0xffffffffu, c[15 - m].re.y, src, 16);
                t.im.x = __shfl_sync(0xffffffffu, c[15 - m].im.x, src, 16);
                t.im.y = __shfl_sync(0xffffffffu, c[15 - m].im.y, src, 16);
                if (l16 == 0) {
                    if (m == 8) t = cx2{pa_n, pb_n};
                    else t = c[16 - m];
                }
                v[m] = t;
            }
            fft256x2_group<true>(v, l16, xch + g * DL4SS_XCH2_FLOAT4, tw);
        } else {
            __syncwarp();
            if (lane == 0) sbar_arrive(&empty[st]);
        }
        // park the windowed, normalised lower half of the first frame for the previous group
        __syncwarp();                                    // both groups' transforms are done with the transpose buffers
        if (warp_live) {
#pragma unroll
            for (int n1 = 0; n1 < 8; ++n1) {
                const int j = 16 * n1 + l16;
                ex[j] = pmul(v[n1].re, pbc(wlo[j]));
            }
        }
        __syncwarp();
        if (warp > 0) sbar_arrive(&xfull[warp]);

        if (g < C::PAIRS) {
            float *o0 = out + ((size_t)b * S + s0) * Lout + l16;
            float *o1 = out + ((size_t)b * S + s1) * Lout + l16;
            // block 2q: frame 2q upper half + frame 2q+1 lower half (both in registers)
            if (tb <= T - 1) {
                const size_t off = (size_t)ta * (NFFT / 2);
#pragma unroll
                for (int n1 = 0; n1 < 8; ++n1) {
                    const int j = 16 * n1 + l16;
                    const float2 r = pfma(v[n1 + 8].re, pbc(whi[j]), pmul(v[n1].im, pbc(wlo[j])));
                    K6_STORE(o0 + off + 16 * n1, r.x);
                    if (two) K6_STORE(o1 + off + 16 * n1, r.y);
                }
            }
        }
        // block 2q+1: frame 2q+1 upper half + the next pair's first frame lower half (the next group's parked values;
        // for the odd group of a warp they come from the next warp)
        if (warp < C::WARPS - 1) sbar_wait(&xfull[warp + 1], (uint32_t)k & 1u);     // (whole warp waits: keeps the warp converged)
        if (g < C::PAIRS && tb + 1 <= T - 1) {
            const float2 *nx = reinterpret_cast<const float2 *>(xch + (g + 1) * DL4SS_XCH2_FLOAT4);
            float *o0 = out + ((size_t)b * S + s0) * Lout + l16;
            float *o1 = out + ((size_t)b * S + s1) * Lout + l16;
            const size_t off = (size_t)tb * (NFFT / 2);
#pragma unroll
            for (int n1 = 0; n1 < 8; ++n1) {
                const int j = 16 * n1 + l16;
                const float2 r = pfma(v[n1 + 8].im, pbc(whi[j]), nx[j]);
                K6_STORE(o0 + off + 16 * n1, r.x);
                if (two) K6_STORE(o1 + off + 16 * n1, r.y);
            }
        }
        __syncwarp();
        if (warp < C::WARPS - 1) sbar_arrive(&xempty[warp + 1]);      // the next warp's parked half frame has been consumed
    }
}

// ------------------------------------------------------------------------------------ K1 staged
constexpr int K1S_STAGES = 2;
constexpr int K1S_CONSUMERS = K1_THREADS;                 // 8 groups x 4 frames = 32 frames per item
constexpr int K1S_THREADS = K1S_CONSUMERS + 32;
constexpr int K1S_SPAN = (K1_FT - 1) * (NFFT / 2) + NFFT + NFFT / 2;   // samples an item's frames touch (33 half frames) + one half
                                                                       // frame to the left: the reflection at the utterance end reaches it

template <typename WavT>
struct K1Stage {
    static constexpr int BYTES = (K1S_SPAN * (int)sizeof(WavT) + 16 + 127) & ~127;
};

template <typename WavT, int FEAT, bool CPLX>
__global__ void __launch_bounds__(K1S_THREADS)
stft256_staged_kernel(const WavT *__restrict__ wav, int L, int T, int tiles_per_utt, int n_items, const char *wav_end,
                      const float *__restrict__ window, float eps, int conj,
                      float *__restrict__ feat, float2 *__restrict__ cplx) {
    constexpr int hop = NFFT / 2;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned char *stage0 = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);   // [STAGES][BYTES]
    float4 *xch = reinterpret_cast<float4 *>(stage0 + K1S_STAGES * K1Stage<WavT>::BYTES);     // groups * 272 float4
    float2 *tw = reinterpret_cast<float2 *>(xch + K1_GROUPS * DL4SS_XCH2_FLOAT4);             // 256 float2
    float *win = reinterpret_cast<float *>(tw + 256);                                        // 256
    uint64_t *full = reinterpret_cast<uint64_t *>(win + NFFT);
    uint64_t *empty = full + K1S_STAGES;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int i = 0; i < K1S_STAGES; ++i) { sbar_init(&full[i], 1); sbar_init(&empty[i], K1S_CONSUMERS / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < NFFT; i += K1S_THREADS) {
        tw[i] = g_tw256[i];
        win[i] = window[i];
    }
    __syncthreads();

    if (warp == K1S_CONSUMERS / 32) {
        const uint64_t pol = evict_first_policy();
        int k = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++k) {
            const int st = k % K1S_STAGES;
            if (k >= K1S_STAGES) sbar_wait(&empty[st], (uint32_t)((k / K1S_STAGES) - 1) & 1u);
            const int b = item / tiles_per_utt;
            const int tile = item - b * tiles_per_utt;
            const int lo = max(tile * K1_FT * hop - NFFT, 0);
            const int hi = min(tile * K1_FT * hop - NFFT + K1S_SPAN, L);
            unsigned char *sb = stage0 + (size_t)st * K1Stage<WavT>::BYTES;
            const Span a = make_span(reinterpret_cast<const char *>(wav + (size_t)b * L + lo), (size_t)(hi - lo) * sizeof(WavT), wav_end);
            span_tail(a, reinterpret_cast<char *>(sb), lane);
            __threadfence_block();
            __syncwarp();
            if (lane == 0) {
                sbar_arrive_tx(&full[st], a.bulk);
                if (a.bulk) bulk_g2s(sb, a.lo, a.bulk, &full[st], pol);
            }
            __syncwarp();
        }
        return;
    }

    const int g = tid >> 4, l16 = tid & 15;
    int k = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++k) {
        const int st = k % K1S_STAGES;
        const int b = item / tiles_per_utt;
        const int tile = item - b * tiles_per_utt;
        const int f0 = tile * K1_FT + 4 * g;
        const int fr[4] = {min(f0, T - 1), min(f0 + 1, T - 1), min(f0 + 2, T - 1), min(f0 + 3, T - 1)};
        const bool ok[4] = {f0 < T, f0 + 1 < T, f0 + 2 < T, f0 + 3 < T};
        const int lo = max(tile * K1_FT * hop - NFFT, 0);                // first staged sample of the utterance
        sbar_wait(&full[st], (uint32_t)(k / K1S_STAGES) & 1u);
        const unsigned char *sb = stage0 + (size_t)st * K1Stage<WavT>::BYTES;
        const WavT *ws = reinterpret_cast<const WavT *>(sb + ((((size_t)b * L + lo) * sizeof(WavT)) & 15)) - lo;   // ws[j] = sample j
        cx2 v[16];
        {
            const int s0 = f0 * hop - NFFT / 2;
            if (ok[3] && s0 >= 0 && s0 + 5 * (NFFT / 2) <= L) {
                float h[5][8];
#pragma unroll
                for (int q = 0; q < 5; ++q)
#pragma unroll
                    for (int n = 0; n < 8; ++n) h[q][n] = (float)ws[s0 + q * (NFFT / 2) + l16 + 16 * n];
#pragma unroll
                for (int n2 = 0; n2 < 16; ++n2) {
                    const float2 wv = pbc(win[l16 + 16 * n2]);
                    const int q = n2 >> 3, n = n2 & 7;
                    v[n2].re = pmul(make_float2(h[q][n], h[q + 2][n]), wv);          // frames f0, f2
                    v[n2].im = pmul(make_float2(h[q + 1][n], h[q + 3][n]), wv);      // frames f1, f3
                }
            } else {
                const int sf[4] = {fr[0] * hop - NFFT / 2, fr[1] * hop - NFFT / 2, fr[2] * hop - NFFT / 2, fr[3] * hop - NFFT / 2};
#pragma unroll
                for (int n2 = 0; n2 < 16; ++n2) {
                    float x[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) x[i] = load_reflect(ws, sf[i] + l16 + 16 * n2, L);
                    const float2 wv = pbc(win[l16 + 16 * n2]);
                    v[n2].re = pmul(make_float2(x[0], x[2]), wv);
                    v[n2].im = pmul(make_float2(x[1], x[3]), wv);
                }
            }
        }
        __syncwarp();
        if (lane == 0) sbar_arrive(&empty[st]);
        fft256x2_group<false>(v, l16, xch + g * DL4SS_XCH2_FLOAT4, tw);
        stft_store_rows<FEAT, CPLX>(v, l16, b, T, fr, ok, eps, conj, feat, cplx);
    }
}

}  // namespace dl4ss
