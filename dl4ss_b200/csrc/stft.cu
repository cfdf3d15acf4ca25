// K1  dl4ss_stft_feat : frame + window + 256-pt FFT + |.| / log(|.|+eps) (+ complex spectrum)
// K6  dl4ss_mask_istft: mask x mixture + Hermitian iFFT + window + overlap-add + normalise
//
// Both are HBM-bound stages: each waveform sample / spectrum bin crosses HBM once.
// Frames are staged in shared memory so the 50-75 % frame overlap never re-reads HBM, two real
// frames ride one complex transform (real/imag packing), and the 256-point transform runs on 16
// lanes x 16 registers with one shared-memory transpose (fft256.cuh).
#include "fft256.cuh"

namespace dl4ss {

constexpr int STFT_THREADS = 256;
constexpr int STFT_GROUPS = STFT_THREADS / 16;   // 16-lane FFT groups per CTA
constexpr int STFT_FT = 2 * STFT_GROUPS;         // frames per tile (two per group)
constexpr int NFFT = 256;
constexpr int NBIN = 129;
// K1 runs 128-thread CTAs (16 frames per tile): 8 CTAs per SM in different phases overlap one CTA's staging
// loads with the others' FFTs, and the last wave is finer grained
constexpr int K1_THREADS = 128;
constexpr int K1_GROUPS = K1_THREADS / 16;
constexpr int K1_FT = 2 * K1_GROUPS;

// exp(-2*pi*i*n1*k2/256) laid out [k2][n1] (fft256.cuh), built once per device
__device__ float2 g_tw256[256];
__global__ void init_twiddles_kernel() { fill_twiddles(g_tw256, threadIdx.x, blockDim.x); }

static int ensure_twiddles(cudaStream_t st) {
    static bool done[64] = {false};
    int dev = 0;
    DL4SS_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) dev = 63;
    if (!done[dev] || dev == 63) {
        init_twiddles_kernel<<<1, 256, 0, st>>>();
        DL4SS_LAUNCH_CHECK("init_twiddles_kernel");
        done[dev] = true;
    }
    return DL4SS_OK;
}

// ------------------------------------------------------------------------------------ K1
__device__ __forceinline__ float sqrt_approx(float x) {      // MUFU.SQRT-class, max relative error 2^-23
    float r;
    asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// FEAT / CPLX are compile time: the output loop is a third of the kernel's instructions
template <typename WavT, int FEAT, bool CPLX>
__global__ void __launch_bounds__(K1_THREADS)
stft256_kernel(const WavT *__restrict__ wav, int L, int hop, int T, int tiles_per_utt,
               const float *__restrict__ window, float eps, int conj,
               float *__restrict__ feat, float2 *__restrict__ cplx) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2 *tw = reinterpret_cast<float2 *>(smem_raw);                 // 256 float2
    float2 *xch = tw + 256;                                            // groups * 272 float2
    float *win = reinterpret_cast<float *>(xch + K1_GROUPS * DL4SS_XCH_FLOAT2);   // 256

    const int tid = threadIdx.x;
    const int b = blockIdx.x / tiles_per_utt;
    const int tile = blockIdx.x - b * tiles_per_utt;
    const int t0 = tile * K1_FT;
    const int nf = min(K1_FT, T - t0);

    for (int i = tid; i < NFFT; i += K1_THREADS) {
        tw[i] = g_tw256[i];
        win[i] = window[i];
    }

    __syncthreads();                       // twiddle / window tables: the kernel's only CTA-wide sync

    const int g = tid >> 4, l16 = tid & 15;
    const int fa = 2 * g, fb = 2 * g + 1;
    const bool va = fa < nf, vb = fb < nf;

    // Every 16-lane group pulls its two frames straight from global memory (64-byte coalesced rows per n2;
    // frame t spans signal [t*hop-128, t*hop+128), reflect-padded at the utterance edges).  The 50-75 % overlap
    // between neighbouring frames is served by L1/L2 (DRAM still sees each sample once), and without a
    // staging phase the warps of a CTA never wait on each other.
    float2 v[16];
    {
        const WavT *w = wav + (size_t)b * L;
        float xa[16], xb[16];
        const int sa0 = (t0 + (va ? fa : 0)) * hop - NFFT / 2;
        const int sb0 = (t0 + (vb ? fb : 0)) * hop - NFFT / 2;
        if (sa0 >= 0 && sa0 + NFFT <= L) {
#pragma unroll
            for (int n2 = 0; n2 < 16; ++n2) xa[n2] = (float)w[sa0 + l16 + 16 * n2];
        } else {
#pragma unroll
            for (int n2 = 0; n2 < 16; ++n2) {
                int j = sa0 + l16 + 16 * n2;
                j = (j < 0) ? -j : j;
                j = (j >= L) ? 2 * (L - 1) - j : j;
                xa[n2] = (float)w[j];
            }
        }
        if (sb0 >= 0 && sb0 + NFFT <= L) {
#pragma unroll
            for (int n2 = 0; n2 < 16; ++n2) xb[n2] = (float)w[sb0 + l16 + 16 * n2];
        } else {
#pragma unroll
            for (int n2 = 0; n2 < 16; ++n2) {
                int j = sb0 + l16 + 16 * n2;
                j = (j < 0) ? -j : j;
                j = (j >= L) ? 2 * (L - 1) - j : j;
                xb[n2] = (float)w[j];
            }
        }
#pragma unroll
        for (int n2 = 0; n2 < 16; ++n2) {
            const float wv = win[l16 + 16 * n2];
            v[n2] = make_float2(xa[n2] * wv, xb[n2] * wv);
        }
    }
    fft256_group<false>(v, l16, xch + g * DL4SS_XCH_FLOAT2, tw);

    // split Z = FFT(xa + i*xb) into the two real-input spectra:
    //   XA[k] = (Z[k] + conj(Z[256-k]))/2 ,  XB[k] = (Z[k] - conj(Z[256-k]))/(2i)
    // lane holds Z[16*k1+l16] in v[k1]; Z[256-k] lives in lane (16-l16)&15, register 15-k1
    // (lane 0: own register (16-k1)&15).
    const size_t rowa = ((size_t)b * T + t0 + fa) * NBIN + l16;
    float *fpa = feat + rowa, *fpb = fpa + NBIN;
    float2 *cpa = cplx + rowa, *cpb = cpa + NBIN;
    const int src = (16 - l16) & 15;
    const float sgn = conj ? -1.0f : 1.0f;
#pragma unroll
    for (int k1 = 0; k1 < 8; ++k1) {
        float2 z = v[k1];
        float px = __shfl_sync(0xffffffffu, v[15 - k1].x, src, 16);
        float py = __shfl_sync(0xffffffffu, v[15 - k1].y, src, 16);
        if (l16 == 0) {
            px = v[(16 - k1) & 15].x;
            py = v[(16 - k1) & 15].y;
        }
        float2 xa = make_float2(0.5f * (z.x + px), 0.5f * (z.y - py));
        float2 xb = make_float2(0.5f * (z.y + py), -0.5f * (z.x - px));
        if (FEAT != DL4SS_FEAT_NONE) {
            float ma = sqrt_approx(fmaf(xa.x, xa.x, xa.y * xa.y));
            float mb = sqrt_approx(fmaf(xb.x, xb.x, xb.y * xb.y));
            if (FEAT == DL4SS_FEAT_LOG) {
                ma = logf(ma + eps);
                mb = logf(mb + eps);
            }
            if (va) fpa[16 * k1] = ma;
            if (vb) fpb[16 * k1] = mb;
        }
        if (CPLX) {
            if (va) cpa[16 * k1] = make_float2(xa.x, sgn * xa.y);
            if (vb) cpb[16 * k1] = make_float2(xb.x, sgn * xb.y);
        }
    }
    if (l16 == 0) {   // Nyquist bin: Z[128] = XA[128] + i*XB[128], both real
        float xa = v[8].x, xb = v[8].y;
        if (FEAT != DL4SS_FEAT_NONE) {
            float ma = fabsf(xa), mb = fabsf(xb);
            if (FEAT == DL4SS_FEAT_LOG) {
                ma = logf(ma + eps);
                mb = logf(mb + eps);
            }
            if (va) fpa[128] = ma;
            if (vb) fpb[128] = mb;
        }
        if (CPLX) {
            if (va) cpa[128] = make_float2(xa, 0.0f);
            if (vb) cpb[128] = make_float2(xb, 0.0f);
        }
    }
}

// Masked spectra of two (source, frame) items -> one complex inverse transform: on return v[n1].x / .y hold the
// (unnormalised, unwindowed) time samples 16*n1 + l16 of item a / item b.
template <int MASK_KIND>
__device__ __forceinline__ void masked_pair_ifft(const float *__restrict__ mask, const float2 *__restrict__ spec,
                                                 int b, int S, int T, int sa, int ta, bool va, int sb, int tb, bool vb,
                                                 int l16, float2 *xch_g, const float2 *tw, float2 (&v)[16]) {
    const int src = (16 - l16) & 15;
    float2 pa[8], pb[8], pa_n = make_float2(0.f, 0.f), pb_n = make_float2(0.f, 0.f);
    if (MASK_KIND == DL4SS_MASK_NONE) {
        const float2 *ra = spec + (((size_t)b * S + sa) * T + ta) * NBIN;
        const float2 *rb = spec + (((size_t)b * S + sb) * T + tb) * NBIN;
#pragma unroll
        for (int m = 0; m < 8; ++m) {
            pa[m] = va ? ra[16 * m + l16] : make_float2(0.f, 0.f);
            pb[m] = vb ? rb[16 * m + l16] : make_float2(0.f, 0.f);
        }
        if (l16 == 0) {
            if (va) pa_n = ra[128];
            if (vb) pb_n = rb[128];
        }
    } else {
        const float2 *xa = spec + ((size_t)b * T + ta) * NBIN;
        const float2 *xb = spec + ((size_t)b * T + tb) * NBIN;
        const size_t ma = (((size_t)b * S + sa) * T + ta) * NBIN;
        const size_t mb = (((size_t)b * S + sb) * T + tb) * NBIN;
        float2 xva[8], xvb[8];
#pragma unroll
        for (int m = 0; m < 8; ++m) {
            xva[m] = va ? xa[16 * m + l16] : make_float2(0.f, 0.f);
            xvb[m] = vb ? xb[16 * m + l16] : make_float2(0.f, 0.f);
        }
        if (MASK_KIND == DL4SS_MASK_REAL) {
#pragma unroll
            for (int m = 0; m < 8; ++m) {
                float ka = va ? mask[ma + 16 * m + l16] : 0.f;
                float kb = vb ? mask[mb + 16 * m + l16] : 0.f;
                pa[m] = make_float2(ka * xva[m].x, ka * xva[m].y);
                pb[m] = make_float2(kb * xvb[m].x, kb * xvb[m].y);
            }
            if (l16 == 0) {
                if (va) { float k = mask[ma + 128]; float2 x = xa[128]; pa_n = make_float2(k * x.x, k * x.y); }
                if (vb) { float k = mask[mb + 128]; float2 x = xb[128]; pb_n = make_float2(k * x.x, k * x.y); }
            }
        } else {
            const float2 *cma = reinterpret_cast<const float2 *>(mask) + ma;
            const float2 *cmb = reinterpret_cast<const float2 *>(mask) + mb;
#pragma unroll
            for (int m = 0; m < 8; ++m) {
                float2 ka = va ? cma[16 * m + l16] : make_float2(0.f, 0.f);
                float2 kb = vb ? cmb[16 * m + l16] : make_float2(0.f, 0.f);
                // reference order: re = Mr*Xr - Mi*Xi ; im = Mr*Xi + Mi*Xr
                pa[m] = make_float2(ka.x * xva[m].x - ka.y * xva[m].y, ka.x * xva[m].y + ka.y * xva[m].x);
                pb[m] = make_float2(kb.x * xvb[m].x - kb.y * xvb[m].y, kb.x * xvb[m].y + kb.y * xvb[m].x);
            }
            if (l16 == 0) {
                if (va) { float2 k = cma[128]; float2 x = xa[128]; pa_n = make_float2(k.x * x.x - k.y * x.y, 0.f); }
                if (vb) { float2 k = cmb[128]; float2 x = xb[128]; pb_n = make_float2(k.x * x.x - k.y * x.y, 0.f); }
            }
        }
    }
    if (l16 == 0) {   // DC bin: irfft ignores the imaginary part
        pa[0].y = 0.f;
        pb[0].y = 0.f;
    }
    // Z[k] = A[k] + i*B[k] for k <= 128 ; Z[256-k] = conj(A[k]) + i*conj(B[k])
    float2 c[8];
#pragma unroll
    for (int m = 0; m < 8; ++m) {
        v[m] = make_float2(pa[m].x - pb[m].y, pa[m].y + pb[m].x);
        c[m] = make_float2(pa[m].x + pb[m].y, pb[m].x - pa[m].y);
    }
#pragma unroll
    for (int m = 8; m < 16; ++m) {
        float cx = __shfl_sync(0xffffffffu, c[15 - m].x, src, 16);
        float cy = __shfl_sync(0xffffffffu, c[15 - m].y, src, 16);
        if (l16 == 0) {
            if (m == 8) { cx = pa_n.x; cy = pb_n.x; }
            else { cx = c[16 - m].x; cy = c[16 - m].y; }
        }
        v[m] = make_float2(cx, cy);
    }
    fft256_group<true>(v, l16, xch_g, tw);
}

// ------------------------------------------------------------------------------------ K6
// One CTA reconstructs `bpt` hop-blocks of every source of one utterance.  It transforms the
// frames that touch those blocks (a leading halo of ceil(256/hop)-1 frames is recomputed by the
// neighbouring tile instead of exchanged), parks windowed frames in shared memory, then every
// thread sums the <= ceil(256/hop) overlapping frames of its output samples and divides by the
// window sum-square envelope (librosa.istft semantics).
template <int MASK_KIND>
__global__ void __launch_bounds__(STFT_THREADS)
istft256_kernel(const float *__restrict__ mask, const float2 *__restrict__ spec, int S, int T,
                int hop, int bpt, int tiles_per_utt, int max_frames,
                const float *__restrict__ window, float *__restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2 *tw = reinterpret_cast<float2 *>(smem_raw);
    float2 *xch = tw + 256;
    float *win = reinterpret_cast<float *>(xch + STFT_GROUPS * DL4SS_XCH_FLOAT2);
    float *wsq = win + NFFT;                                           // 256: window^2 (un-normalised taps)
    float *inv_full = wsq + NFFT;                                      // 256: 1/sum of window^2 over a full frame set
    float *ybuf = inv_full + NFFT;                                          // max_frames*S*256

    const int tid = threadIdx.x;
    const int b = blockIdx.x / tiles_per_utt;
    const int tile = blockIdx.x - b * tiles_per_utt;
    const int Lout = hop * (T - 1);
    // padded sample coordinates m = n + 128
    const int m_lo = NFFT / 2 + tile * bpt * hop;
    const int m_hi = min(m_lo + bpt * hop, NFFT / 2 + Lout);
    int t_lo = (m_lo - (NFFT - 1) + hop - 1) / hop;          // ceil((m_lo-255)/hop), m_lo >= 128
    if (m_lo - (NFFT - 1) < 0) t_lo = 0;
    const int t_hi = min(T - 1, (m_hi - 1) / hop);
    const int nfr = t_hi - t_lo + 1;
    const int nitems = nfr * S;
    const int npairs = (nitems + 1) >> 1;

    tw[tid] = g_tw256[tid];
    {
        const float wt = window[tid];
        win[tid] = wt * (1.0f / NFFT);          // fold the 1/N of the inverse transform
        wsq[tid] = wt * wt;
    }
    __syncthreads();

    const int g = tid >> 4, l16 = tid & 15;
    const int rounds = (npairs + STFT_GROUPS - 1) / STFT_GROUPS;
    for (int r = 0; r < rounds; ++r) {
        const int p = r * STFT_GROUPS + g;
        const int ia = 2 * p, ib = 2 * p + 1;
        const bool va = ia < nitems, vb = ib < nitems;
        const int ta = t_lo + (va ? ia / S : 0), sa = va ? ia % S : 0;
        const int tb = t_lo + (vb ? ib / S : 0), sb = vb ? ib % S : 0;

        float2 v[16];
        masked_pair_ifft<MASK_KIND>(mask, spec, b, S, T, sa, ta, va, sb, tb, vb, l16, xch + g * DL4SS_XCH_FLOAT2, tw, v);
        float *ya = ybuf + (size_t)ia * NFFT;
        float *yb = ybuf + (size_t)ib * NFFT;
#pragma unroll
        for (int n1 = 0; n1 < 16; ++n1) {
            const int n = 16 * n1 + l16;
            const float w = win[n];
            if (va) ya[n] = v[n1].x * w;
            if (vb) yb[n] = v[n1].y * w;
        }
    }
    __syncthreads();

    // overlap-add + window-sum-square normalisation + trim
    const int span = m_hi - m_lo;
    const int hop_shift = ((hop & (hop - 1)) == 0 && hop >= 4) ? (31 - __clz(hop)) : -1;
    if (hop_shift >= 0) {
        // power-of-two hop (64 / 128 in every reference config): R = 256/hop frames cover every interior
        // sample; their window^2 sum depends only on m mod hop -> reciprocal table, no divisions
        const int R = NFFT >> hop_shift;
        if (tid < hop) {          // wsq is visible: written before the __syncthreads above
            float e = 0.f;
            for (int r = 0; r < R; ++r) e += wsq[tid + (r << hop_shift)];
            inv_full[tid] = (e > 1.17549435e-38f) ? 1.0f / e : 1.0f;
        }
        __syncthreads();
        const int span4 = span >> 2;
        for (int s = 0; s < S; ++s) {
            float *o = out + ((size_t)b * S + s) * Lout + (m_lo - NFFT / 2);
            for (int i4 = tid; i4 < span4; i4 += STFT_THREADS) {
                const int m = m_lo + (i4 << 2);
                const int t1 = m >> hop_shift;
                const int n0 = m - (t1 << hop_shift);
                float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
                if (t1 <= t_hi && t1 - (R - 1) >= t_lo) {
                    for (int r = 0; r < R; ++r) {
                        const float4 yv = *reinterpret_cast<const float4 *>(
                            ybuf + (size_t)((t1 - r - t_lo) * S + s) * NFFT + n0 + (r << hop_shift));
                        acc.x += yv.x; acc.y += yv.y; acc.z += yv.z; acc.w += yv.w;
                    }
                    const float4 iv = *reinterpret_cast<const float4 *>(inv_full + n0);
                    acc.x *= iv.x; acc.y *= iv.y; acc.z *= iv.z; acc.w *= iv.w;
                } else {          // utterance edges: partial frame set
                    float4 env = make_float4(0.f, 0.f, 0.f, 0.f);
                    for (int t = min(t_hi, t1); t >= t_lo && m - (t << hop_shift) < NFFT; --t) {
                        const int n = m - (t << hop_shift);
                        const float4 yv = *reinterpret_cast<const float4 *>(ybuf + (size_t)((t - t_lo) * S + s) * NFFT + n);
                        const float4 wv = *reinterpret_cast<const float4 *>(wsq + n);
                        acc.x += yv.x; acc.y += yv.y; acc.z += yv.z; acc.w += yv.w;
                        env.x += wv.x; env.y += wv.y; env.z += wv.z; env.w += wv.w;
                    }
                    acc.x = (env.x > 1.17549435e-38f) ? acc.x / env.x : acc.x;
                    acc.y = (env.y > 1.17549435e-38f) ? acc.y / env.y : acc.y;
                    acc.z = (env.z > 1.17549435e-38f) ? acc.z / env.z : acc.z;
                    acc.w = (env.w > 1.17549435e-38f) ? acc.w / env.w : acc.w;
                }
                *reinterpret_cast<float4 *>(o + (i4 << 2)) = acc;
            }
        }
    } else if ((hop & 3) == 0) {
        // 4 consecutive samples share their frame set (hop % 4 == 0): float4 smem reads, one float4 store
        const int span4 = span >> 2;          // span = blocks*hop is a multiple of 4
        for (int idx = tid; idx < S * span4; idx += STFT_THREADS) {
            const int s = idx / span4;
            const int i = (idx - s * span4) << 2;
            const int m = m_lo + i;
            const int t1 = min(t_hi, m / hop);
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f), env = acc;
            for (int t = t1; t >= t_lo && m - t * hop < NFFT; --t) {
                const int n = m - t * hop;
                const float4 yv = *reinterpret_cast<const float4 *>(ybuf + (size_t)((t - t_lo) * S + s) * NFFT + n);
                const float4 wv = *reinterpret_cast<const float4 *>(wsq + n);
                acc.x += yv.x; acc.y += yv.y; acc.z += yv.z; acc.w += yv.w;
                env.x += wv.x; env.y += wv.y; env.z += wv.z; env.w += wv.w;
            }
            float4 o4;
            o4.x = (env.x > 1.17549435e-38f) ? acc.x / env.x : acc.x;
            o4.y = (env.y > 1.17549435e-38f) ? acc.y / env.y : acc.y;
            o4.z = (env.z > 1.17549435e-38f) ? acc.z / env.z : acc.z;
            o4.w = (env.w > 1.17549435e-38f) ? acc.w / env.w : acc.w;
            *reinterpret_cast<float4 *>(out + ((size_t)b * S + s) * Lout + (m_lo - NFFT / 2) + i) = o4;
        }
    } else {
        for (int s = 0; s < S; ++s) {
            float *o = out + ((size_t)b * S + s) * Lout + (m_lo - NFFT / 2);
            for (int i = tid; i < span; i += STFT_THREADS) {
                const int m = m_lo + i;
                int t1 = min(t_hi, m / hop);
                float acc = 0.f, env = 0.f;
                for (int t = t1; t >= t_lo && m - t * hop < NFFT; --t) {
                    const int n = m - t * hop;
                    acc += ybuf[(size_t)((t - t_lo) * S + s) * NFFT + n];
                    env += wsq[n];
                }
                o[i] = (env > 1.17549435e-38f) ? acc / env : acc;
            }
        }
    }
}

// ------------------------------------------------------------------------------------ K6, hop = n_fft/2
// Every reference config has hop 128 = half a frame: exactly two frames cover each output sample.  A 16-lane
// group transforms the frame pair (2q, 2q+1) of ONE source, so the hop block between them is summed, normalised
// and stored straight from registers; the block after frame 2q+1 needs the lower half of the NEXT pair's first
// frame, which every group parks in an 8 KB exchange buffer (one __syncthreads).  No frame buffer, no separate
// overlap-add pass; the group after the tile's last pair is recomputed as halo by the neighbouring CTA.
constexpr int K6H_PAIRS = STFT_GROUPS - 1;       // pairs a CTA owns (the 16th group is the halo pair)

template <int MASK_KIND>
__global__ void __launch_bounds__(STFT_THREADS)
istft_h128_kernel(const float *__restrict__ mask, const float2 *__restrict__ spec, int S, int T, int tiles_per_src,
                  const float *__restrict__ window, float *__restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2 *tw = reinterpret_cast<float2 *>(smem_raw);
    float2 *xch = tw + 256;
    float *win = reinterpret_cast<float *>(xch + STFT_GROUPS * DL4SS_XCH_FLOAT2);   // 256: window / N
    float *inv = win + NFFT;                                                        // 128: 1 / (w^2[j] + w^2[j+128])
    float *exch = inv + NFFT / 2;                                                   // [groups][128]

    const int tid = threadIdx.x;
    int bid = blockIdx.x;
    const int s = bid % S; bid /= S;                 // sources of an utterance run back to back: X stays in L2
    const int tile = bid % tiles_per_src;
    const int b = bid / tiles_per_src;
    const int Lout = (NFFT / 2) * (T - 1);

    tw[tid] = g_tw256[tid];
    {
        const float wt = window[tid];
        win[tid] = wt * (1.0f / NFFT);
        if (tid < NFFT / 2) {
            const float w2 = window[tid + NFFT / 2];
            const float e = wt * wt + w2 * w2;
            inv[tid] = (e > 1.17549435e-38f) ? 1.0f / e : 1.0f;
        }
    }
    __syncthreads();

    const int g = tid >> 4, l16 = tid & 15;
    const int q = tile * K6H_PAIRS + g;              // group K6H_PAIRS is the halo pair
    const int ta = 2 * q, tb = 2 * q + 1;
    const bool va = ta < T, vb = (tb < T) && (g < K6H_PAIRS);     // the halo group only needs its first frame

    float2 v[16];
    masked_pair_ifft<MASK_KIND>(mask, spec, b, S, T, s, va ? ta : 0, va, s, vb ? tb : 0, vb, l16,
                                xch + g * DL4SS_XCH_FLOAT2, tw, v);

    // park the windowed lower half of the first frame for the previous group
    float *ex = exch + g * (NFFT / 2);
#pragma unroll
    for (int n1 = 0; n1 < 8; ++n1) {
        const int j = 16 * n1 + l16;
        ex[j] = v[n1].x * win[j];
    }
    __syncthreads();
    if (g >= K6H_PAIRS) return;

    float *o = out + ((size_t)b * S + s) * Lout;
    // block 2q: frame 2q upper half + frame 2q+1 lower half (both in registers)
    if (tb <= T - 1) {
        float *ob = o + (size_t)ta * (NFFT / 2) + l16;
#pragma unroll
        for (int n1 = 0; n1 < 8; ++n1) {
            const int j = 16 * n1 + l16;
            ob[16 * n1] = (v[n1 + 8].x * win[j + NFFT / 2] + v[n1].y * win[j]) * inv[j];
        }
    }
    // block 2q+1: frame 2q+1 upper half + the next pair's first frame lower half
    if (tb + 1 <= T - 1) {
        const float *nx = exch + (g + 1) * (NFFT / 2);
        float *ob = o + (size_t)tb * (NFFT / 2) + l16;
#pragma unroll
        for (int n1 = 0; n1 < 8; ++n1) {
            const int j = 16 * n1 + l16;
            ob[16 * n1] = (v[n1 + 8].y * win[j + NFFT / 2] + nx[j]) * inv[j];
        }
    }
}

}  // namespace dl4ss

using namespace dl4ss;

extern "C" int dl4ss_stft_feat(const void *wav, int wav_dtype, int B, int L, int n_fft, int hop,
                               const float *window, int feat_mode, float eps, int conj,
                               float *feat_out, float *cplx_out, void *stream) {
    if (B == 0) return DL4SS_OK;                 // empty batch: nothing to read or write (pointers may be null)
    DL4SS_CHECK_ARG(wav && window, "stft_feat: null wav/window");
    DL4SS_CHECK_ARG(B >= 0 && L > n_fft / 2, "stft_feat: need B>=0 and L > n_fft/2 (reflect pad), got B=%d L=%d", B, L);
    DL4SS_CHECK_ARG(hop >= 1 && hop <= n_fft, "stft_feat: hop must be in [1,n_fft], got %d", hop);
    DL4SS_CHECK_ARG(feat_mode >= DL4SS_FEAT_NONE && feat_mode <= DL4SS_FEAT_LOG, "stft_feat: bad feat_mode %d", feat_mode);
    DL4SS_CHECK_ARG(feat_mode == DL4SS_FEAT_NONE || feat_out, "stft_feat: feat_out is null");
    DL4SS_CHECK_ARG(feat_out || cplx_out, "stft_feat: no output requested");
    DL4SS_CHECK_ARG(wav_dtype == DL4SS_WAV_F32 || wav_dtype == DL4SS_WAV_F64, "stft_feat: bad wav_dtype %d", wav_dtype);
    if (n_fft != NFFT) {
        set_error("stft_feat: this build has the 256-point transform only (n_fft=%d)", n_fft);
        return DL4SS_EUNSUPPORTED;
    }
    if (B == 0) return DL4SS_OK;
    const int T = 1 + L / hop;
    const int tiles = cdiv(T, K1_FT);
    const size_t smem = 256 * sizeof(float2) + K1_GROUPS * DL4SS_XCH_FLOAT2 * sizeof(float2) +
                        NFFT * sizeof(float);
    cudaStream_t st = (cudaStream_t)stream;
    { int rc = ensure_twiddles(st); if (rc) return rc; }
    const long long grid = (long long)B * tiles;
    DL4SS_CHECK_ARG(grid < (1ll << 31), "stft_feat: grid too large");
#define LAUNCH_K1(WT, FM, CP)                                                                                  \
    do {                                                                                                        \
        DL4SS_CUDA(cudaFuncSetAttribute(stft256_kernel<WT, FM, CP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        stft256_kernel<WT, FM, CP><<<(unsigned)grid, K1_THREADS, smem, st>>>(                                   \
            (const WT *)wav, L, hop, T, tiles, window, eps, conj, feat_out, (float2 *)cplx_out);                \
    } while (0)
#define DISPATCH_K1(WT)                                                                                         \
    do {                                                                                                        \
        if (cplx_out) {                                                                                         \
            if (feat_mode == DL4SS_FEAT_NONE) LAUNCH_K1(WT, DL4SS_FEAT_NONE, true);                             \
            else if (feat_mode == DL4SS_FEAT_ABS) LAUNCH_K1(WT, DL4SS_FEAT_ABS, true);                          \
            else LAUNCH_K1(WT, DL4SS_FEAT_LOG, true);                                                           \
        } else {                                                                                                \
            if (feat_mode == DL4SS_FEAT_ABS) LAUNCH_K1(WT, DL4SS_FEAT_ABS, false);                              \
            else LAUNCH_K1(WT, DL4SS_FEAT_LOG, false);                                                          \
        }                                                                                                       \
    } while (0)
    if (wav_dtype == DL4SS_WAV_F32) DISPATCH_K1(float);
    else DISPATCH_K1(double);
#undef DISPATCH_K1
#undef LAUNCH_K1
    DL4SS_LAUNCH_CHECK("stft256_kernel");
    return DL4SS_OK;
}

extern "C" int dl4ss_mask_istft(const float *mask, int mask_kind, const float *spec, int B, int S,
                                int T, int n_fft, int hop, const float *window, float *wav_out,
                                void *stream) {
    if (B == 0) return DL4SS_OK;
    DL4SS_CHECK_ARG(spec && window && wav_out, "mask_istft: null spec/window/out");
    DL4SS_CHECK_ARG(mask_kind >= DL4SS_MASK_NONE && mask_kind <= DL4SS_MASK_COMPLEX, "mask_istft: bad mask_kind %d", mask_kind);
    DL4SS_CHECK_ARG(mask_kind == DL4SS_MASK_NONE || mask, "mask_istft: mask is null");
    DL4SS_CHECK_ARG(B >= 0 && S >= 1 && T >= 1, "mask_istft: bad B/S/T %d/%d/%d", B, S, T);
    DL4SS_CHECK_ARG(hop >= 1 && hop <= n_fft, "mask_istft: hop must be in [1,n_fft], got %d", hop);
    if (n_fft != NFFT) {
        set_error("mask_istft: this build has the 256-point transform only (n_fft=%d)", n_fft);
        return DL4SS_EUNSUPPORTED;
    }
    if (B == 0 || T == 1) return DL4SS_OK;     // hop*(T-1) == 0 samples
    cudaStream_t st0 = (cudaStream_t)stream;
    if (hop == NFFT / 2) {          // the reference's hop: fused pair kernel
        { int rc = ensure_twiddles(st0); if (rc) return rc; }
        const int pairs = (T + 1) / 2;
        const int tiles_h = cdiv(pairs, K6H_PAIRS);
        const size_t smem_h = 256 * sizeof(float2) + STFT_GROUPS * DL4SS_XCH_FLOAT2 * sizeof(float2) +
                              (NFFT + NFFT / 2) * sizeof(float) + (size_t)STFT_GROUPS * (NFFT / 2) * sizeof(float);
        const long long grid_h = (long long)B * S * tiles_h;
        DL4SS_CHECK_ARG(grid_h < (1ll << 31), "mask_istft: grid too large");
#define LAUNCH_H128(KIND)                                                                                       \
        do {                                                                                                    \
            DL4SS_CUDA(cudaFuncSetAttribute(istft_h128_kernel<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_h)); \
            istft_h128_kernel<KIND><<<(unsigned)grid_h, STFT_THREADS, smem_h, st0>>>(                           \
                mask, (const float2 *)spec, S, T, tiles_h, window, wav_out);                                    \
        } while (0)
        if (mask_kind == DL4SS_MASK_NONE) LAUNCH_H128(DL4SS_MASK_NONE);
        else if (mask_kind == DL4SS_MASK_REAL) LAUNCH_H128(DL4SS_MASK_REAL);
        else LAUNCH_H128(DL4SS_MASK_COMPLEX);
#undef LAUNCH_H128
        DL4SS_LAUNCH_CHECK("istft_h128_kernel");
        return DL4SS_OK;
    }
    const int halo = (NFFT - 1) / hop;         // frames before the first block's own frame
    const int nblocks = T - 1;
    // frames per CTA: aim at 32 (frame,source) items = one round of 16 two-frame groups
    int max_items = (S <= 16) ? 32 : 2 * S;
    int bpt = max_items / S - halo;
    while (bpt < 1) { max_items += 32; bpt = max_items / S - halo; }
    if (bpt > nblocks) bpt = nblocks;
    const int tiles = cdiv(nblocks, bpt);
    bpt = cdiv(nblocks, tiles);
    const int max_frames = bpt + halo + 1;
    const size_t smem = 256 * sizeof(float2) + STFT_GROUPS * DL4SS_XCH_FLOAT2 * sizeof(float2) +
                        3 * NFFT * sizeof(float) + (size_t)max_frames * S * NFFT * sizeof(float);
    if (smem > 220 * 1024) {
        set_error("mask_istft: S=%d hop=%d needs %zu B of shared memory", S, hop, smem);
        return DL4SS_EUNSUPPORTED;
    }
    cudaStream_t st = (cudaStream_t)stream;
    { int rc = ensure_twiddles(st); if (rc) return rc; }
    const long long grid = (long long)B * tiles;
    DL4SS_CHECK_ARG(grid < (1ll << 31), "mask_istft: grid too large");
#define LAUNCH_ISTFT(KIND)                                                                             \
    do {                                                                                               \
        DL4SS_CUDA(cudaFuncSetAttribute(istft256_kernel<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        istft256_kernel<KIND><<<(unsigned)grid, STFT_THREADS, smem, st>>>(                             \
            mask, (const float2 *)spec, S, T, hop, bpt, tiles, max_frames, window, wav_out);           \
    } while (0)
    if (mask_kind == DL4SS_MASK_NONE) LAUNCH_ISTFT(DL4SS_MASK_NONE);
    else if (mask_kind == DL4SS_MASK_REAL) LAUNCH_ISTFT(DL4SS_MASK_REAL);
    else LAUNCH_ISTFT(DL4SS_MASK_COMPLEX);
#undef LAUNCH_ISTFT
    DL4SS_LAUNCH_CHECK("istft256_kernel");
    return DL4SS_OK;
}
