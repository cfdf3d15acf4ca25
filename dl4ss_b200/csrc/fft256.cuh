// 256-point complex FFT on 16 cooperating lanes (16 points per lane), used by the STFT and
// iSTFT kernels through the two-real-frames-in-one-complex-transform packing.
//
// Decomposition n = n1 + 16*n2, k = 16*k1 + k2:
//   X[16*k1+k2] = sum_{n1} W16^(n1*k1) * [ W256^(n1*k2) * sum_{n2} x[n1+16*n2] * W16^(n2*k2) ]
// pass 1: lane n1 holds x[n1+16*n2] (n2 = register index), 16-point FFT over n2, twiddle,
//         transpose through shared memory (row pitch 17 float2: conflict-free 64-bit accesses);
// pass 2: lane k2 holds Y[n1][k2] (n1 = register index), 16-point FFT over n1 ->
//         register k1 holds X[16*k1+k2].
// INV = true uses the conjugate roots (unnormalised inverse).
#pragma once
#include "common.cuh"

namespace dl4ss {

#define DL4SS_SQRT1_2 0.70710678118654752440f
#define DL4SS_COS_PI_8 0.92387953251128675613f
#define DL4SS_SIN_PI_8 0.38268343236508977173f

// multiply by W16^m (forward: exp(-2*pi*i*m/16); inverse: conjugate), m compile time
template <int M, bool INV>
__device__ __forceinline__ float2 mul_w16(float2 a) {
    constexpr int m = M & 15;
    if constexpr (m == 0) return a;
    else if constexpr (m == 4) return INV ? make_float2(-a.y, a.x) : make_float2(a.y, -a.x);   // -+ i
    else if constexpr (m == 8) return make_float2(-a.x, -a.y);
    else if constexpr (m == 12) return INV ? make_float2(a.y, -a.x) : make_float2(-a.y, a.x);
    else if constexpr (m == 2) {   // (1 -+ i)/sqrt2
        return INV ? make_float2((a.x - a.y) * DL4SS_SQRT1_2, (a.x + a.y) * DL4SS_SQRT1_2)
                   : make_float2((a.x + a.y) * DL4SS_SQRT1_2, (a.y - a.x) * DL4SS_SQRT1_2);
    } else if constexpr (m == 6) { // (-1 -+ i)/sqrt2
        return INV ? make_float2((-a.x - a.y) * DL4SS_SQRT1_2, (a.x - a.y) * DL4SS_SQRT1_2)
                   : make_float2((a.y - a.x) * DL4SS_SQRT1_2, (-a.x - a.y) * DL4SS_SQRT1_2);
    } else {
        constexpr float c = (m == 1) ? DL4SS_COS_PI_8 : (m == 3) ? DL4SS_SIN_PI_8
                          : (m == 9) ? -DL4SS_COS_PI_8 : /* unused */ 0.f;
        constexpr float s = (m == 1) ? DL4SS_SIN_PI_8 : (m == 3) ? DL4SS_COS_PI_8
                          : (m == 9) ? -DL4SS_SIN_PI_8 : 0.f;
        static_assert(m == 1 || m == 3 || m == 9, "twiddle not needed by the 4x4 split");
        // forward root = (c, -s); inverse = (c, +s)
        return INV ? make_float2(fmaf(a.x, c, -a.y * s), fmaf(a.x, s, a.y * c))
                   : make_float2(fmaf(a.x, c, a.y * s), fmaf(a.y, c, -a.x * s));
    }
}

template <bool INV>
__device__ __forceinline__ void fft4(float2 &a0, float2 &a1, float2 &a2, float2 &a3) {
    float2 s02 = cadd(a0, a2), d02 = csub(a0, a2), s13 = cadd(a1, a3), d13 = csub(a1, a3);
    float2 r = INV ? make_float2(-d13.y, d13.x) : make_float2(d13.y, -d13.x);   // +-i * d13
    a0 = cadd(s02, s13);
    a2 = csub(s02, s13);
    a1 = cadd(d02, r);
    a3 = csub(d02, r);
}

// in-place 16-point FFT, natural order in and out.  n = n1 + 4*n2, k = 4*k1 + k2.
template <bool INV>
__device__ __forceinline__ void fft16(float2 (&v)[16]) {
    // stage 1: for each n1, FFT4 over n2 (elements n1, n1+4, n1+8, n1+12) -> slot n1+4*k2
#pragma unroll
    for (int n1 = 0; n1 < 4; ++n1) fft4<INV>(v[n1], v[n1 + 4], v[n1 + 8], v[n1 + 12]);
    // twiddle W16^(n1*k2) on slot n1+4*k2
    v[5] = mul_w16<1, INV>(v[5]);
    v[6] = mul_w16<2, INV>(v[6]);
    v[7] = mul_w16<3, INV>(v[7]);
    v[9] = mul_w16<2, INV>(v[9]);
    v[10] = mul_w16<4, INV>(v[10]);
    v[11] = mul_w16<6, INV>(v[11]);
    v[13] = mul_w16<3, INV>(v[13]);
    v[14] = mul_w16<6, INV>(v[14]);
    v[15] = mul_w16<9, INV>(v[15]);
    // stage 2: for each k2, FFT4 over n1 (slots 4*k2 + 0..3) -> slot 4*k2 + k1 holds X[4*k1+k2]
#pragma unroll
    for (int k2 = 0; k2 < 4; ++k2) fft4<INV>(v[4 * k2], v[4 * k2 + 1], v[4 * k2 + 2], v[4 * k2 + 3]);
    // un-permute (compile-time register renaming): X[4*k1+k2] sits in slot 4*k2+k1
    float2 t;
    t = v[1]; v[1] = v[4]; v[4] = t;
    t = v[2]; v[2] = v[8]; v[8] = t;
    t = v[3]; v[3] = v[12]; v[12] = t;
    t = v[6]; v[6] = v[9]; v[9] = t;
    t = v[7]; v[7] = v[13]; v[13] = t;
    t = v[11]; v[11] = v[14]; v[14] = t;
}

#define DL4SS_XCH_PITCH 17                       // float2 per row of the 16x16 transpose buffer
#define DL4SS_XCH_FLOAT2 (16 * DL4SS_XCH_PITCH)  // float2 per 16-lane group

// twiddle table layout: tw[k2*16 + n1] = exp(-2*pi*i*n1*k2/256)   (k2 = 0 row unused)
// lane = index inside the 16-lane group; xch = this group's transpose buffer.
// The caller guarantees the 16 lanes of a group sit in one warp (half-warp aligned).
template <bool INV>
__device__ __forceinline__ void fft256_group(float2 (&v)[16], int lane16, float2 *xch,
                                             const float2 *__restrict__ tw) {
    fft16<INV>(v);
#pragma unroll
    for (int k2 = 1; k2 < 16; ++k2) {
        float2 w = tw[k2 * 16 + lane16];
        if (INV) w.y = -w.y;
        v[k2] = cmul(v[k2], w);
    }
    __syncwarp();   // previous readers of xch are done
#pragma unroll
    for (int k2 = 0; k2 < 16; ++k2) xch[lane16 * DL4SS_XCH_PITCH + k2] = v[k2];
    __syncwarp();
#pragma unroll
    for (int n1 = 0; n1 < 16; ++n1) v[n1] = xch[n1 * DL4SS_XCH_PITCH + lane16];
    fft16<INV>(v);
}

__device__ __forceinline__ void fill_twiddles(float2 *tw, int tid, int nthreads) {
    for (int i = tid; i < 256; i += nthreads) {
        int k2 = i >> 4, n1 = i & 15;
        float s, c;
        sincospif(-(float)(n1 * k2) / 128.0f, &s, &c);
        tw[i] = make_float2(c, s);
    }
}

}  // namespace dl4ss
