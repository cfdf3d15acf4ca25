// TWO 256-point complex FFTs per 16-lane group on the sm_100a packed fp32 pipe (FADD2 / FMUL2 / FFMA2).
//
// Same 16 x 16 decomposition as fft256.cuh, but every value is a pair: `cx2` holds the real parts of
// transform A and transform B in one 64-bit register pair and the imaginary parts in another (structure of
// arrays across the two transforms).  Every butterfly, +-i rotation (a re/im renaming plus a sign) and twiddle
// product is then ONE packed instruction for both transforms: the instruction count per transform halves,
// and so do the shared-memory transposes (128-bit accesses) and the twiddle-table reads.
// With two real frames riding each complex transform, a 16-lane group moves 4 real frames per pass.
#pragma once
#include "fft256.cuh"

namespace dl4ss {

// ---- packed pairs: .x = transform A, .y = transform B
__device__ __forceinline__ unsigned long long pk2(float2 a) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a.x), "f"(a.y));
    return r;
}
__device__ __forceinline__ float2 upk2(unsigned long long r) {
    float2 a;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a.x), "=f"(a.y) : "l"(r));
    return a;
}
__device__ __forceinline__ float2 padd(float2 a, float2 b) {
    unsigned long long r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pk2(a)), "l"(pk2(b)));
    return upk2(r);
}
__device__ __forceinline__ float2 psub(float2 a, float2 b) {
    unsigned long long r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pk2(a)), "l"(pk2(b)));
    return upk2(r);
}
__device__ __forceinline__ float2 pmul(float2 a, float2 b) {
    unsigned long long r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pk2(a)), "l"(pk2(b)));
    return upk2(r);
}
__device__ __forceinline__ float2 pfma(float2 a, float2 b, float2 c) {          // a*b + c
    unsigned long long r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(pk2(a)), "l"(pk2(b)), "l"(pk2(c)));
    return upk2(r);
}
__device__ __forceinline__ float2 pfnma(float2 a, float2 b, float2 c) {         // c - a*b
    return pfma(make_float2(-a.x, -a.y), b, c);
}
__device__ __forceinline__ float2 pbc(float c) { return make_float2(c, c); }

struct cx2 {
    float2 re, im;
};
__device__ __forceinline__ cx2 operator+(cx2 a, cx2 b) { return cx2{padd(a.re, b.re), padd(a.im, b.im)}; }
__device__ __forceinline__ cx2 operator-(cx2 a, cx2 b) { return cx2{psub(a.re, b.re), psub(a.im, b.im)}; }

// multiply both transforms by W16^m (forward: exp(-2*pi*i*m/16); INV: conjugate), m compile time
template <int M, bool INV>
__device__ __forceinline__ cx2 mul_w16(cx2 a) {
    constexpr int m = M & 15;
    if constexpr (m == 0) return a;
    else if constexpr (m == 4) {          // -+ i
        return INV ? cx2{make_float2(-a.im.x, -a.im.y), a.re} : cx2{a.im, make_float2(-a.re.x, -a.re.y)};
    } else if constexpr (m == 2) {        // (1 -+ i)/sqrt2
        const float2 r = pbc(DL4SS_SQRT1_2);
        return INV ? cx2{pmul(psub(a.re, a.im), r), pmul(padd(a.re, a.im), r)}
                   : cx2{pmul(padd(a.re, a.im), r), pmul(psub(a.im, a.re), r)};
    } else if constexpr (m == 6) {        // (-1 -+ i)/sqrt2
        const float2 r = pbc(DL4SS_SQRT1_2), nr = pbc(-DL4SS_SQRT1_2);
        return INV ? cx2{pmul(padd(a.re, a.im), nr), pmul(psub(a.re, a.im), r)}
                   : cx2{pmul(psub(a.im, a.re), r), pmul(padd(a.re, a.im), nr)};
    } else {
        static_assert(m == 1 || m == 3 || m == 9, "twiddle not needed by the 4x4 split");
        constexpr float c = (m == 1) ? DL4SS_COS_PI_8 : (m == 3) ? DL4SS_SIN_PI_8 : -DL4SS_COS_PI_8;
        constexpr float s = (m == 1) ? DL4SS_SIN_PI_8 : (m == 3) ? DL4SS_COS_PI_8 : -DL4SS_SIN_PI_8;
        // forward root = (c, -s): (re*c + im*s, im*c - re*s); inverse = (c, +s): (re*c - im*s, im*c + re*s)
        return INV ? cx2{pfma(a.im, pbc(-s), pmul(a.re, pbc(c))), pfma(a.re, pbc(s), pmul(a.im, pbc(c)))}
                   : cx2{pfma(a.im, pbc(s), pmul(a.re, pbc(c))), pfma(a.re, pbc(-s), pmul(a.im, pbc(c)))};
    }
}

template <bool INV>
__device__ __forceinline__ void fft4(cx2 &a0, cx2 &a1, cx2 &a2, cx2 &a3) {
    const cx2 s02 = a0 + a2, d02 = a0 - a2, s13 = a1 + a3, d13 = a1 - a3;
    a0 = s02 + s13;
    a2 = s02 - s13;
    if (INV) {   // r = +i*d13 = (-d13.im, d13.re)
        a1 = cx2{psub(d02.re, d13.im), padd(d02.im, d13.re)};
        a3 = cx2{padd(d02.re, d13.im), psub(d02.im, d13.re)};
    } else {     // r = -i*d13 = (d13.im, -d13.re)
        a1 = cx2{padd(d02.re, d13.im), psub(d02.im, d13.re)};
        a3 = cx2{psub(d02.re, d13.im), padd(d02.im, d13.re)};
    }
}

template <bool INV>
__device__ __forceinline__ void fft16(cx2 (&v)[16]) {
#pragma unroll
    for (int n1 = 0; n1 < 4; ++n1) fft4<INV>(v[n1], v[n1 + 4], v[n1 + 8], v[n1 + 12]);
    v[5] = mul_w16<1, INV>(v[5]);
    v[6] = mul_w16<2, INV>(v[6]);
    v[7] = mul_w16<3, INV>(v[7]);
    v[9] = mul_w16<2, INV>(v[9]);
    v[10] = mul_w16<4, INV>(v[10]);
    v[11] = mul_w16<6, INV>(v[11]);
    v[13] = mul_w16<3, INV>(v[13]);
    v[14] = mul_w16<6, INV>(v[14]);
    v[15] = mul_w16<9, INV>(v[15]);
#pragma unroll
    for (int k2 = 0; k2 < 4; ++k2) fft4<INV>(v[4 * k2], v[4 * k2 + 1], v[4 * k2 + 2], v[4 * k2 + 3]);
    cx2 t;
    t = v[1]; v[1] = v[4]; v[4] = t;
    t = v[2]; v[2] = v[8]; v[8] = t;
    t = v[3]; v[3] = v[12]; v[12] = t;
    t = v[6]; v[6] = v[9]; v[9] = t;
    t = v[7]; v[7] = v[13]; v[13] = t;
    t = v[11]; v[11] = v[14]; v[14] = t;
}

#define DL4SS_XCH2_PITCH 17                          // float4 per row of the 16x16 transpose buffer
#define DL4SS_XCH2_FLOAT4 (16 * DL4SS_XCH2_PITCH)    // float4 per 16-lane group (4352 B)

// tw[k2*16 + n1] = exp(-2*pi*i*n1*k2/256) (the forward table of fft256.cuh; INV conjugates).  The two groups of a
// warp read the same table entries (64-bit loads: one wavefront per half-warp); the (w, w) broadcast pairs the
// packed pipe needs are built with register moves, which are cheaper here than wider shared-memory reads.
// xch = this group's transpose buffer (16-byte aligned).
template <bool INV>
__device__ __forceinline__ void fft256x2_group(cx2 (&v)[16], int lane16, float4 *xch, const float2 *__restrict__ tw) {
    fft16<INV>(v);
#pragma unroll
    for (int k2 = 1; k2 < 16; ++k2) {
        const float2 w = tw[k2 * 16 + lane16];
        const float2 wx = pbc(w.x), wy = pbc(w.y);
        const cx2 a = v[k2];
        if (INV) {   // multiply by conj(w)
            v[k2].re = pfma(a.im, wy, pmul(a.re, wx));
            v[k2].im = pfnma(a.re, wy, pmul(a.im, wx));
        } else {
            v[k2].re = pfnma(a.im, wy, pmul(a.re, wx));
            v[k2].im = pfma(a.re, wy, pmul(a.im, wx));
        }
    }
    __syncwarp();   // previous readers of xch are done
#pragma unroll
    for (int k2 = 0; k2 < 16; ++k2)
        xch[lane16 * DL4SS_XCH2_PITCH + k2] = make_float4(v[k2].re.x, v[k2].re.y, v[k2].im.x, v[k2].im.y);
    __syncwarp();
#pragma unroll
    for (int n1 = 0; n1 < 16; ++n1) {
        const float4 t = xch[n1 * DL4SS_XCH2_PITCH + lane16];
        v[n1].re = make_float2(t.x, t.y);
        v[n1].im = make_float2(t.z, t.w);
    }
    fft16<INV>(v);
}

__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

}  // namespace dl4ss
