// K1  dl4ss_stft_feat : frame + window + 256-pt FFT + |.| / log(|.|+eps) (+ complex spectrum)
// K6  dl4ss_mask_istft: mask x mixture + Hermitian iFFT + window + overlap-add + normalise
//
// Both are HBM-bound stages: each waveform sample / spectrum bin crosses HBM once.
// Frames are staged in shared memory so the 50-75 % frame overlap never re-reads HBM, two real
// frames ride one complex transform (real/imag packing), and the 256-point transform runs on 16
// lanes x 16 registers with one shared-memory transpose (fft256.cuh).
#include "fft256x2.cuh"
#include <stdlib.h>

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
constexpr int K1_FT = 4 * K1_GROUPS;         // four frames per group (two packed complex transforms)

// exp(-2*pi*i*n1*k2/256) laid out [k2][n1] (fft256.cuh), built once per device
__device__ float2 g_tw256[256];
__global__ void init_twiddles_kernel() { fill_twiddles(g_tw256, threadIdx.x, blockDim.x); }

// Built exactly once per device, whatever host thread or stream gets here first (std::call_once per device): the table
// is written on a private stream that is synchronised before the flag is set, so a launch on ANY stream that follows
// sees it.  The capture-mode exchange keeps a first use inside a stream capture legal.
static int ensure_twiddles(cudaStream_t) {
    static std::once_flag once[DL4SS_MAX_DEVICES];
    static int status[DL4SS_MAX_DEVICES];
    int dev = 0;
    DL4SS_CUDA(cudaGetDevice(&dev));
    DL4SS_CHECK_ARG(dev >= 0 && dev < DL4SS_MAX_DEVICES, "device ordinal %d out of range", dev);
    std::call_once(once[dev], [dev]() {
        cudaStreamCaptureMode mode = cudaStreamCaptureModeRelaxed;
        cudaThreadExchangeStreamCaptureMode(&mode);
        cudaStream_t s = nullptr;
        cudaError_t e = cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
        if (e == cudaSuccess) {
            init_twiddles_kernel<<<1, 256, 0, s>>>();
            e = cudaGetLastError();
            if (e == cudaSuccess) e = cudaStreamSynchronize(s);
            cudaStreamDestroy(s);
        }
        cudaThreadExchangeStreamCaptureMode(&mode);
        status[dev] = (int)e;
    });
    if (status[dev] != 0) {
        set_error("init_twiddles_kernel: %s", cudaGetErrorString((cudaError_t)status[dev]));
        return DL4SS_ECUDA;
    }
    return DL4SS_OK;
}

// DL4SS_STAGED=0 in the environment selects the first-round direct-load kernels (A/B measurements)
static bool staged_enabled() {
    static const bool on = [] { const char *e = getenv("DL4SS_STAGED"); return !(e && e[0] == '0'); }();
    return on;
}

// K1 is store dominated (3 output bytes per input byte) and runs best with the 16 resident warps of the direct-load form:
// the staged K1 (12 warps per SM) measured 61 % against 67 % at the bench size, so it is opt-in (DL4SS_STAGED_K1=1)
static bool staged_k1_enabled() {
    static const bool on = [] { const char *e = getenv("DL4SS_STAGED_K1"); return e && e[0] == '1'; }();
    return on;
}
// a persistent (looping) form of the direct-load K1 measured SLOWER than one item per CTA (59 % against 69 % at the bench
// size, 59 % against 89 % at large batch): fresh CTAs overlap their load phase with their neighbours' store phase better
// than a resident CTA walking its items in order; opt-in (DL4SS_PERSISTENT_K1=1)
static bool persistent_k1_enabled() {
    static const bool on = [] { const char *e = getenv("DL4SS_PERSISTENT_K1"); return e && e[0] == '1'; }();
    return on;
}

// ------------------------------------------------------------------------------------ K1
__device__ __forceinline__ float sqrt_approx(float x) {      // MUFU.SQRT-class, max relative error 2^-23 (ftz: |X|^2 below 1.2e-38 reads as 0, far under eps)
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// FEAT / CPLX are compile time: the output loop is a third of the kernel's instructions.
// A 16-lane group transforms FOUR consecutive frames: complex transform A carries frames (f0, f1) in its real /
// imaginary parts, transform B frames (f2, f3), and the two transforms run packed on the FADD2/FMUL2/FFMA2 pipe
// (fft256x2.cuh).  H128 (hop = n_fft/2, every reference config): the four frames span five half-frames, so a
// lane loads 40 samples instead of 64.
template <typename WavT>
__device__ __forceinline__ float load_reflect(const WavT *__restrict__ w, int j, int L) {
    j = (j < 0) ? -j : j;
    j = (j >= L) ? 2 * (L - 1) - j : j;
    return (float)w[j];
}

// Output stage of K1: split the two packed complex transforms into four real-input spectra and store |X| / log|X| and
// the complex rows.  v = the group's transforms after fft256x2_group; fr / ok = the four frames (clamped) and validity.
template <int FEAT, bool CPLX>
__device__ __forceinline__ void stft_store_rows(const cx2 (&v)[16], int l16, int b, int T, const int (&fr)[4], const bool (&ok)[4],
                                                float eps, int conj, float *__restrict__ feat, float2 *__restrict__ cplx) {
    // split Z = FFT(xa + i*xb) into the two real-input spectra (per transform):
    //   XA[k] = (Z[k] + conj(Z[256-k]))/2 ,  XB[k] = (Z[k] - conj(Z[256-k]))/(2i)
    // lane holds Z[16*k1+l16] in v[k1]; Z[256-k] lives in lane (16-l16)&15, register 15-k1
    // (lane 0: own register (16-k1)&15).
    size_t row[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) row[i] = ((size_t)b * T + fr[i]) * NBIN + l16;
    const int src = (16 - l16) & 15;
    const float sg = conj ? -0.5f : 0.5f;
    const float2 half2 = pbc(0.5f), sgn2 = pbc(sg), nsgn2 = pbc(-sg);
#pragma unroll
    for (int k1 = 0; k1 < 8; ++k1) {
        const cx2 z = v[k1];
        cx2 p;
        p.re.x = __shfl_sync(0xffffffffu, v[15 - k1].re.x, src, 16);
        p.re.y = __shfl_sync(0xffffffffu, v[15 - k1].re.y, src, 16);
        p.im.x = __shfl_sync(0xffffffffu, v[15 - k1].im.x, src, 16);
        p.im.y = __shfl_sync(0xffffffffu, v[15 - k1].im.y, src, 16);
        if (l16 == 0) p = v[(16 - k1) & 15];
        // 2*XA = (sre, dim) ; 2*XB = (sim, -dre)
        const float2 sre = padd(z.re, p.re), dim = psub(z.im, p.im), sim = padd(z.im, p.im), dre = psub(z.re, p.re);
        if (FEAT != DL4SS_FEAT_NONE) {
            const float2 qa = pfma(sre, sre, pmul(dim, dim)), qb = pfma(sim, sim, pmul(dre, dre));
            float m[4] = {0.5f * sqrt_approx(qa.x), 0.5f * sqrt_approx(qb.x), 0.5f * sqrt_approx(qa.y), 0.5f * sqrt_approx(qb.y)};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                if (FEAT == DL4SS_FEAT_LOG) m[i] = logf(m[i] + eps);
                if (ok[i]) feat[row[i] + 16 * k1] = m[i];
            }
        }
        if (CPLX) {
            const float2 are = pmul(sre, half2), aim = pmul(dim, sgn2), bre = pmul(sim, half2), bim = pmul(dre, nsgn2);
            if (ok[0]) cplx[row[0] + 16 * k1] = make_float2(are.x, aim.x);
            if (ok[1]) cplx[row[1] + 16 * k1] = make_float2(bre.x, bim.x);
            if (ok[2]) cplx[row[2] + 16 * k1] = make_float2(are.y, aim.y);
            if (ok[3]) cplx[row[3] + 16 * k1] = make_float2(bre.y, bim.y);
        }
    }
    if (l16 == 0) {   // Nyquist bin: Z[128] = XA[128] + i*XB[128], both real
        const float x[4] = {v[8].re.x, v[8].im.x, v[8].re.y, v[8].im.y};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (!ok[i]) continue;
            if (FEAT != DL4SS_FEAT_NONE) {
                float m = fabsf(x[i]);
                if (FEAT == DL4SS_FEAT_LOG) m = logf(m + eps);
                feat[row[i] + 128] = m;
            }
            if (CPLX) cplx[row[i] + 128] = make_float2(x[i], 0.0f);
        }
    }
}

template <typename WavT, int FEAT, bool CPLX, int HDIV>   // HDIV = n_fft / hop when that is 2 or 4, else 0
__global__ void __launch_bounds__(K1_THREADS)
stft256_kernel(const WavT *__restrict__ wav, int L, int hop, int T, int tiles_per_utt, int n_items,
               const float *__restrict__ window, float eps, int conj, int pf_dist,
               float *__restrict__ feat, float2 *__restrict__ cplx) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4 *xch = reinterpret_cast<float4 *>(smem_raw);                  // groups * 272 float4
    float2 *tw = reinterpret_cast<float2 *>(xch + K1_GROUPS * DL4SS_XCH2_FLOAT4);     // 256 float2
    float *win = reinterpret_cast<float *>(tw + 256);                    // 256

    const int tid = threadIdx.x;
    for (int i = tid; i < NFFT; i += K1_THREADS) {
        tw[i] = g_tw256[i];
        win[i] = window[i];
    }
    __syncthreads();                       // twiddle / window tables: the kernel's only CTA-wide sync

    // persistent form: gridDim.x < n_items, every CTA walks items blockIdx.x, + gridDim.x, ... (the tables above are
    // built once, the last wave is one item long); gridDim.x == n_items is the one-item-per-CTA form
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
    const int b = item / tiles_per_utt;
    const int tile = item - b * tiles_per_utt;
    const int t0 = tile * K1_FT;

    // pull the waveform span of the item `pf_dist` ahead (the next resident wave / this CTA's next item) into L2 (no
    // registers, no scoreboard): by the time it is processed, its loads are L2 hits instead of DRAM round trips
    if (item + pf_dist < n_items) {
        const int pb = (item + pf_dist) / tiles_per_utt;
        const int pt = (item + pf_dist) - pb * tiles_per_utt;
        const int j = min(max(pt * K1_FT * hop - NFFT / 2 + tid * (int)(128 / sizeof(WavT)), 0), L - 1);
        if (tid * (int)(128 / sizeof(WavT)) < (K1_FT - 1) * hop + NFFT) prefetch_l2(wav + (size_t)pb * L + j);
    }

    const int g = tid >> 4, l16 = tid & 15;
    const int f0 = t0 + 4 * g;
    // frames past the end are clamped (valid loads, results never stored); no early exit: the two groups of a
    // warp meet in __syncwarp / __shfl_sync
    const int fr[4] = {min(f0, T - 1), min(f0 + 1, T - 1), min(f0 + 2, T - 1), min(f0 + 3, T - 1)};
    const bool ok[4] = {f0 < T, f0 + 1 < T, f0 + 2 < T, f0 + 3 < T};

    // Every group pulls its frames straight from global memory (64-byte coalesced rows per n2; frame t spans signal
    // [t*hop-128, t*hop+128), reflect-padded at the utterance edges).  The overlap between neighbouring groups is
    // served by L1/L2 (DRAM sees each sample once); without a staging phase the warps never wait on each other.
    cx2 v[16];
    {
        const WavT *w = wav + (size_t)b * L;
        const int s0 = f0 * hop - NFFT / 2;
        if (HDIV == 2 && ok[3]) {
            float h[5][8];
            if (s0 >= 0 && s0 + 5 * (NFFT / 2) <= L) {
#pragma unroll
                for (int q = 0; q < 5; ++q)
#pragma unroll
                    for (int n = 0; n < 8; ++n) h[q][n] = (float)w[s0 + q * (NFFT / 2) + l16 + 16 * n];
            } else {
#pragma unroll
                for (int q = 0; q < 5; ++q)
#pragma unroll
                    for (int n = 0; n < 8; ++n) h[q][n] = load_reflect(w, s0 + q * (NFFT / 2) + l16 + 16 * n, L);
            }
#pragma unroll
            for (int n2 = 0; n2 < 16; ++n2) {
                const float2 wv = pbc(win[l16 + 16 * n2]);
                const int q = n2 >> 3, n = n2 & 7;
                v[n2].re = pmul(make_float2(h[q][n], h[q + 2][n]), wv);          // frames f0, f2
                v[n2].im = pmul(make_float2(h[q + 1][n], h[q + 3][n]), wv);      // frames f1, f3
            }
        } else if (HDIV == 4 && ok[3]) {
            // hop 64: the four frames span 448 samples = 28 sixteen-sample columns; frame i, column n2 is column 4i + n2
            float h[28];
            if (s0 >= 0 && s0 + 7 * (NFFT / 4) <= L) {
#pragma unroll
                for (int k = 0; k < 28; ++k) h[k] = (float)w[s0 + 16 * k + l16];
            } else {
#pragma unroll
                for (int k = 0; k < 28; ++k) h[k] = load_reflect(w, s0 + 16 * k + l16, L);
            }
#pragma unroll
            for (int n2 = 0; n2 < 16; ++n2) {
                const float2 wv = pbc(win[l16 + 16 * n2]);
                v[n2].re = pmul(make_float2(h[n2], h[8 + n2]), wv);               // frames f0, f2
                v[n2].im = pmul(make_float2(h[4 + n2], h[12 + n2]), wv);          // frames f1, f3
            }
        } else {
            const int sf[4] = {fr[0] * hop - NFFT / 2, fr[1] * hop - NFFT / 2, fr[2] * hop - NFFT / 2, fr[3] * hop - NFFT / 2};
            const bool interior = (sf[0] >= 0) && (sf[3] + NFFT <= L);
#pragma unroll
            for (int n2 = 0; n2 < 16; ++n2) {
                float x[4];
                if (interior) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) x[i] = (float)w[sf[i] + l16 + 16 * n2];
                } else {
#pragma unroll
                    for (int i = 0; i < 4; ++i) x[i] = load_reflect(w, sf[i] + l16 + 16 * n2, L);
                }
                const float2 wv = pbc(win[l16 + 16 * n2]);
                v[n2].re = pmul(make_float2(x[0], x[2]), wv);
                v[n2].im = pmul(make_float2(x[1], x[3]), wv);
            }
        }
    }
    fft256x2_group<false>(v, l16, xch + g * DL4SS_XCH2_FLOAT4, tw);

    stft_store_rows<FEAT, CPLX>(v, l16, b, T, fr, ok, eps, conj, feat, cplx);
    }
}

// Masked spectra of two (source, frame) items -> one complex inverse transform: on return v[n1].x / .y hold the
// (unnormalised, unwindowed) time samples 16*n1 + l16 of item a / item b.  Items flagged invalid still carry
// in-range (clamped) indices: they are loaded like any other and their samples are never stored, which keeps
// the address arithmetic out of predicated code.
template <int MASK_KIND>
__device__ __forceinline__ void masked_pair_ifft(const float *__restrict__ mask, const float2 *__restrict__ spec,
                                                 int b, int S, int T, int sa, int ta, int sb, int tb,
                                                 int l16, float2 *xch_g, const float2 *tw, float2 (&v)[16]) {
    const int src = (16 - l16) & 15;
    float2 pa[8], pb[8], pa_n = make_float2(0.f, 0.f), pb_n = make_float2(0.f, 0.f);
    if (MASK_KIND == DL4SS_MASK_NONE) {
        const float2 *ra = spec + (((size_t)b * S + sa) * T + ta) * NBIN;
        const float2 *rb = spec + (((size_t)b * S + sb) * T + tb) * NBIN;
#pragma unroll
        for (int m = 0; m < 8; ++m) {
            pa[m] = ra[16 * m + l16];
            pb[m] = rb[16 * m + l16];
        }
        if (l16 == 0) {
            pa_n = ra[128];
            pb_n = rb[128];
        }
    } else {
        const float2 *xa = spec + ((size_t)b * T + ta) * NBIN;
        const float2 *xb = spec + ((size_t)b * T + tb) * NBIN;
        const size_t ma = (((size_t)b * S + sa) * T + ta) * NBIN;
        const size_t mb = (((size_t)b * S + sb) * T + tb) * NBIN;
        float2 xva[8], xvb[8];
#pragma unroll
        for (int m = 0; m < 8; ++m) {
            xva[m] = xa[16 * m + l16];
            xvb[m] = xb[16 * m + l16];
        }
        if (MASK_KIND == DL4SS_MASK_REAL) {
#pragma unroll
            for (int m = 0; m < 8; ++m) {
                float ka = mask[ma + 16 * m + l16];
                float kb = mask[mb + 16 * m + l16];
                pa[m] = make_float2(ka * xva[m].x, ka * xva[m].y);
                pb[m] = make_float2(kb * xvb[m].x, kb * xvb[m].y);
            }
            if (l16 == 0) {
                { float k = mask[ma + 128]; float2 x = xa[128]; pa_n = make_float2(k * x.x, k * x.y); }
                { float k = mask[mb + 128]; float2 x = xb[128]; pb_n = make_float2(k * x.x, k * x.y); }
            }
        } else {
            const float2 *cma = reinterpret_cast<const float2 *>(mask) + ma;
            const float2 *cmb = reinterpret_cast<const float2 *>(mask) + mb;
#pragma unroll
            for (int m = 0; m < 8; ++m) {
                float2 ka = cma[16 * m + l16];
                float2 kb = cmb[16 * m + l16];
                // reference order: re = Mr*Xr - Mi*Xi ; im = Mr*Xi + Mi*Xr
                pa[m] = make_float2(ka.x * xva[m].x - ka.y * xva[m].y, ka.x * xva[m].y + ka.y * xva[m].x);
                pb[m] = make_float2(kb.x * xvb[m].x - kb.y * xvb[m].y, kb.x * xvb[m].y + kb.y * xvb[m].x);
            }
            if (l16 == 0) {
                { float2 k = cma[128]; float2 x = xa[128]; pa_n = make_float2(k.x * x.x - k.y * x.y, 0.f); }
                { float2 k = cmb[128]; float2 x = xb[128]; pb_n = make_float2(k.x * x.x - k.y * x.y, 0.f); }
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
        masked_pair_ifft<MASK_KIND>(mask, spec, b, S, T, sa, ta, sb, tb, l16, xch + g * DL4SS_XCH_FLOAT2, tw, v);
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
// group inverse-transforms the frame pair (2q, 2q+1) of TWO sources at once (the two packed transforms of
// fft256x2.cuh; the mixture spectrum is loaded once for both), so the hop block between the two frames is summed,
// normalised and stored straight from registers; the block after frame 2q+1 needs the lower half of the NEXT
// pair's first frame, which every group parks in a 16 KB exchange buffer (one __syncthreads).  No frame buffer, no
// separate overlap-add pass; the pair after the tile's last one is recomputed as halo by the neighbouring CTA.
// the separated waveforms are final outputs: streaming (evict-first) stores keep them from displacing the rows the
// L2 prefetch pulled in one wave ahead (mask + iSTFT 69.1 -> 70.0 % of the measured HBM peak at 30 s x 512)
#define K6_STORE(p, v) __stcs((p), (v))
constexpr int K6H_PAIRS = STFT_GROUPS - 1;       // pairs a CTA owns (the 16th group is the halo pair)

template <int MASK_KIND>
__global__ void __launch_bounds__(STFT_THREADS, 2)
istft_h128_kernel(const float *__restrict__ mask, const float2 *__restrict__ spec, int S, int T, int tiles_per_src,
                  int pf_dist, const float *__restrict__ window, float *__restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4 *xch = reinterpret_cast<float4 *>(smem_raw);                              // groups * 272
    float2 *tw = reinterpret_cast<float2 *>(xch + STFT_GROUPS * DL4SS_XCH2_FLOAT4);  // 256
    float2 *exch = tw + 256;                                                         // [groups][128] (src0, src1)
    float *wlo = reinterpret_cast<float *>(exch + STFT_GROUPS * (NFFT / 2));         // 128: w[j] / (N * env[j])
    float *whi = wlo + NFFT / 2;                                                     // 128: w[j+128] / (N * env[j])

    const int tid = threadIdx.x;
    const int SP = (S + 1) >> 1;
    int bid = blockIdx.x;
    const int sp = bid % SP; bid /= SP;              // source pairs of an utterance run back to back: X stays in L2
    const int tile = bid % tiles_per_src;
    const int b = bid / tiles_per_src;
    const int s0 = 2 * sp, s1 = min(s0 + 1, S - 1);
    const bool two = (s0 + 1 < S);
    const int Lout = (NFFT / 2) * (T - 1);

    const int g = tid >> 4, l16 = tid & 15;
    // pull the rows of the CTA one resident wave ahead into L2 (no registers, no scoreboard)
    if (blockIdx.x + pf_dist < gridDim.x) {
        int pid = blockIdx.x + pf_dist;
        const int psp = pid % SP; pid /= SP;
        const int ptile = pid % tiles_per_src, pb = pid / tiles_per_src;
        const int pta = min(2 * (ptile * K6H_PAIRS + g), T - 1);
        const int rows = min(2, T - pta);            // frames 2q, 2q+1 are adjacent rows
        const char *px = reinterpret_cast<const char *>(spec + (MASK_KIND == DL4SS_MASK_NONE
                             ? (((size_t)pb * S + 2 * psp) * T + pta) * NBIN : ((size_t)pb * T + pta) * NBIN));
        const int xbytes = rows * NBIN * 8;
        if (l16 * 128 < xbytes) prefetch_l2(px + l16 * 128);
        if (l16 == 0 && 16 * 128 < xbytes) prefetch_l2(px + 16 * 128);
        if (MASK_KIND != DL4SS_MASK_NONE) {
            const int mb = (MASK_KIND == DL4SS_MASK_COMPLEX) ? 8 : 4, mbytes = rows * NBIN * mb;
            for (int sidx = 2 * psp; sidx < min(2 * psp + 2, S); ++sidx) {
                const char *pm = reinterpret_cast<const char *>(mask) + (((size_t)pb * S + sidx) * T + pta) * NBIN * mb;
                if (l16 * 128 < mbytes) prefetch_l2(pm + l16 * 128);
                if (l16 == 0 && 16 * 128 < mbytes) prefetch_l2(pm + 16 * 128);
            }
        } else if (2 * psp + 1 < S) {
            const char *p1 = px + (size_t)T * NBIN * 8;
            if (l16 * 128 < xbytes) prefetch_l2(p1 + l16 * 128);
            if (l16 == 0 && 16 * 128 < xbytes) prefetch_l2(p1 + 16 * 128);
        }
    }
    tw[tid] = g_tw256[tid];
    if (tid < NFFT / 2) {
        // out[j] = (frame_hi[j+128]*w[j+128] + frame_lo[j]*w[j]) / (w[j]^2 + w[j+128]^2), 1/N of the inverse folded in
        const float w1 = window[tid], w2 = window[tid + NFFT / 2];
        const float e = w1 * w1 + w2 * w2;
        const float inv = ((e > 1.17549435e-38f) ? 1.0f / e : 1.0f) * (1.0f / NFFT);
        wlo[tid] = w1 * inv;
        whi[tid] = w2 * inv;
    }
    __syncthreads();

    const int q = tile * K6H_PAIRS + g;              // group K6H_PAIRS is the halo pair
    const int ta = 2 * q, tb = 2 * q + 1;
    // out-of-range frames are clamped: loaded like any other, never stored
    const int tac = min(ta, T - 1);
    const int tbc = (tb < T && g < K6H_PAIRS) ? tb : tac;     // the halo group only needs its first frame

    // warps whose two groups both lie past the last frame (last tile of a source) skip the transform; the test is
    // warp-uniform because the groups meet in __syncwarp / __shfl_sync
    const bool warp_live = 2 * (tile * K6H_PAIRS + (g & ~1)) < T;
    cx2 v[16];
    if (warp_live) {
        // pa / pb: masked spectra of frame a / b, packed over the two sources
        cx2 pa[8], pb[8];
        float2 pa_n, pb_n;                           // Nyquist bin (real part only), lane 0
        if (MASK_KIND == DL4SS_MASK_NONE) {
            const float2 *ra0 = spec + (((size_t)b * S + s0) * T + tac) * NBIN, *ra1 = spec + (((size_t)b * S + s1) * T + tac) * NBIN;
            const float2 *rb0 = spec + (((size_t)b * S + s0) * T + tbc) * NBIN, *rb1 = spec + (((size_t)b * S + s1) * T + tbc) * NBIN;
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
        } else {
            const float2 *xa = spec + ((size_t)b * T + tac) * NBIN;
            const float2 *xb = spec + ((size_t)b * T + tbc) * NBIN;
            const size_t ma0 = (((size_t)b * S + s0) * T + tac) * NBIN, ma1 = (((size_t)b * S + s1) * T + tac) * NBIN;
            const size_t mb0 = (((size_t)b * S + s0) * T + tbc) * NBIN, mb1 = (((size_t)b * S + s1) * T + tbc) * NBIN;
            float2 xva[8], xvb[8];
            if (SP == 1) {      // one source pair: the mixture rows are read exactly once, evict-first like the masks
#pragma unroll
                for (int m = 0; m < 8; ++m) {
                    xva[m] = __ldcs(xa + 16 * m + l16);
                    xvb[m] = __ldcs(xb + 16 * m + l16);
                }
            } else {
#pragma unroll
                for (int m = 0; m < 8; ++m) {
                    xva[m] = xa[16 * m + l16];
                    xvb[m] = xb[16 * m + l16];
                }
            }
            if (MASK_KIND == DL4SS_MASK_REAL) {
#pragma unroll
                for (int m = 0; m < 8; ++m) {
                    const float2 ka = make_float2(__ldcs(mask + ma0 + 16 * m + l16), __ldcs(mask + ma1 + 16 * m + l16));
                    const float2 kb = make_float2(__ldcs(mask + mb0 + 16 * m + l16), __ldcs(mask + mb1 + 16 * m + l16));
                    pa[m] = cx2{pmul(ka, pbc(xva[m].x)), pmul(ka, pbc(xva[m].y))};
                    pb[m] = cx2{pmul(kb, pbc(xvb[m].x)), pmul(kb, pbc(xvb[m].y))};
                }
                pa_n = pb_n = make_float2(0.f, 0.f);
                if (l16 == 0) {
                    const float xan = xa[128].x, xbn = xb[128].x;
                    pa_n = make_float2(mask[ma0 + 128] * xan, mask[ma1 + 128] * xan);
                    pb_n = make_float2(mask[mb0 + 128] * xbn, mask[mb1 + 128] * xbn);
                }
            } else {
                const float2 *cm = reinterpret_cast<const float2 *>(mask);
#pragma unroll
                for (int m = 0; m < 8; ++m) {
                    const float2 a0 = cm[ma0 + 16 * m + l16], a1 = cm[ma1 + 16 * m + l16];
                    const float2 b0 = cm[mb0 + 16 * m + l16], b1 = cm[mb1 + 16 * m + l16];
                    const float2 kar = make_float2(a0.x, a1.x), kai = make_float2(a0.y, a1.y);
                    const float2 kbr = make_float2(b0.x, b1.x), kbi = make_float2(b0.y, b1.y);
                    // reference order: re = Mr*Xr - Mi*Xi ; im = Mr*Xi + Mi*Xr
                    pa[m] = cx2{pfnma(kai, pbc(xva[m].y), pmul(kar, pbc(xva[m].x))), pfma(kai, pbc(xva[m].x), pmul(kar, pbc(xva[m].y)))};
                    pb[m] = cx2{pfnma(kbi, pbc(xvb[m].y), pmul(kbr, pbc(xvb[m].x))), pfma(kbi, pbc(xvb[m].x), pmul(kbr, pbc(xvb[m].y)))};
                }
                pa_n = pb_n = make_float2(0.f, 0.f);
                if (l16 == 0) {
                    const float2 xan = xa[128], xbn = xb[128];
                    const float2 a0 = cm[ma0 + 128], a1 = cm[ma1 + 128], b0 = cm[mb0 + 128], b1 = cm[mb1 + 128];
                    pa_n = make_float2(a0.x * xan.x - a0.y * xan.y, a1.x * xan.x - a1.y * xan.y);
                    pb_n = make_float2(b0.x * xbn.x - b0.y * xbn.y, b1.x * xbn.x - b1.y * xbn.y);
                }
            }
        }
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
        // v[n1].re = samples 16*n1+l16 of frame a (src0, src1) ; v[n1].im = frame b

        // park the windowed, normalised lower half of the first frame for the previous group
        float2 *ex = exch + g * (NFFT / 2);
#pragma unroll
        for (int n1 = 0; n1 < 8; ++n1) {
            const int j = 16 * n1 + l16;
            ex[j] = pmul(v[n1].re, pbc(wlo[j]));
        }
    }
    // The parked half frame is consumed by the group before: inside a warp that is a __syncwarp, across warps only
    // warp w-1 waits for warp w (named barrier 1+w, 64 threads: 32 arrive, 32 sync) - no CTA-wide barrier, so a
    // warp whose loads landed late holds back one neighbour instead of all eight.
    const int warp = tid >> 5;
    __syncwarp();
    if (warp > 0) asm volatile("bar.arrive %0, 64;" ::"r"(warp + 1) : "memory");
    if (warp < STFT_THREADS / 32 - 1) asm volatile("bar.sync %0, 64;" ::"r"(warp + 2) : "memory");
    if (g >= K6H_PAIRS) return;

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
    // block 2q+1: frame 2q+1 upper half + the next pair's first frame lower half
    if (tb + 1 <= T - 1) {
        const float2 *nx = exch + (g + 1) * (NFFT / 2);
        const size_t off = (size_t)tb * (NFFT / 2);
#pragma unroll
        for (int n1 = 0; n1 < 8; ++n1) {
            const int j = 16 * n1 + l16;
            const float2 r = pfma(v[n1 + 8].im, pbc(whi[j]), nx[j]);
            K6_STORE(o0 + off + 16 * n1, r.x);
            if (two) K6_STORE(o1 + off + 16 * n1, r.y);
        }
    }
}


// ------------------------------------------------------------------------------------ K6, hop = n_fft/4
// hop 64 (BASELINE configs[0], Torch_multi/config.py FRAME_SHIFT = 64): four frames cover every output sample.  A 16-lane group
// inverse-transforms FOUR consecutive frames (4q .. 4q+3) of ONE source: the two packed complex transforms carry (f0, f1) and
// (f2, f3).  Output sample j = 16a + l16 of a 64-sample block needs samples 64r + j of four different frames -- all of them
// in registers of the SAME lane (sample 16*n1 + l16 sits in v[n1]) -- so the overlap-add is plain register arithmetic:
//     block 4q+3 = f0[192+j] + f1[128+j] + f2[64+j] + f3[j]                       complete inside the group
//     block 4q+4 = f1[192+j] + f2[128+j] + f3[64+j]           + next.f0[j]
//     block 4q+5 = f2[192+j] + f3[128+j]                      + next.(f0[64+j] + f1[j])
//     block 4q+6 = f3[192+j]                                  + next.(f0[128+j] + f1[64+j] + f2[j])
// where the three `next` sums (192 floats) are parked by the following group in shared memory; the group after the tile's last
// one is recomputed as halo.  Blocks are counted in padded samples (frame t starts at 64 t); output block k = padded block k + 2
// (the n_fft/2 trim); frames outside [0, T) contribute nothing and the window sum-square envelope of the (at most two)
// blocks they touch is summed from the frames that exist, as librosa.istft does.
constexpr int K6Q_GROUPS = STFT_GROUPS - 1;      // output groups per CTA (the 16th is the halo)

template <int MASK_KIND>
__global__ void __launch_bounds__(STFT_THREADS, 2)
istft_h64_kernel(const float *__restrict__ mask, const float2 *__restrict__ spec, int S, int T, int tiles_per_src,
                 const float *__restrict__ window, float *__restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4 *xch = reinterpret_cast<float4 *>(smem_raw);                              // groups * 272
    float2 *tw = reinterpret_cast<float2 *>(xch + STFT_GROUPS * DL4SS_XCH2_FLOAT4);  // 256
    float *win = reinterpret_cast<float *>(tw + 256);                                // 256: w[n] / N
    float *inv_full = win + NFFT;                                                    // 64: 1 / sum_r w^2[64 r + j]
    // 71.3 KB in all.  Registers (126) hold it at two CTAs per SM: capping at 80 for three (24 warps) spills 150-270 bytes into
    // the transform and measured 0.30 / 0.39 of the HBM peak instead of 0.47 / 0.58 (real / cRM masks).

    const int tid = threadIdx.x;
    int bid = blockIdx.x;
    const int tile = bid % tiles_per_src; bid /= tiles_per_src;
    const int s = bid % S;
    const int b = bid / S;
    const int Lout = 64 * (T - 1);
    const int g = tid >> 4, l16 = tid & 15;

    tw[tid] = g_tw256[tid];
    win[tid] = window[tid] * (1.0f / NFFT);
    if (tid < 64) {
        const float w0 = window[tid], w1 = window[tid + 64], w2 = window[tid + 128], w3 = window[tid + 192];
        const float e = w0 * w0 + w1 * w1 + w2 * w2 + w3 * w3;
        inv_full[tid] = (e > 1.17549435e-38f) ? 1.0f / e : 1.0f;
    }
    __syncthreads();

    const int q = tile * K6Q_GROUPS + g - 1;         // frames 4q .. 4q+3; q = -1 only parks nothing and outputs padded block 2
    int fr[4];
    bool fv[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int t = 4 * q + i;
        fv[i] = (t >= 0 && t < T);
        fr[i] = min(max(t, 0), T - 1);               // clamped: loaded like any other frame, zeroed below
    }
    // warps whose groups hold no frame at all skip the transform (warp-uniform: both groups of a warp are tested)
    const int q_lo = tile * K6Q_GROUPS + (g & ~1) - 1;
    const bool warp_live = (4 * q_lo < T) && (4 * (q_lo + 1) + 3 >= 0);
    cx2 v[16];
    if (warp_live) {
        // pa: masked spectra of frames (f0 | f2), pb: (f1 | f3), packed over the two transforms
        cx2 pa[8], pb[8];
        float2 pa_n = make_float2(0.f, 0.f), pb_n = make_float2(0.f, 0.f);
        const size_t xrow = (MASK_KIND == DL4SS_MASK_NONE) ? ((size_t)b * S + s) * T : (size_t)b * T;
        const float2 *x0 = spec + (xrow + fr[0]) * NBIN, *x1 = spec + (xrow + fr[1]) * NBIN;
        const float2 *x2 = spec + (xrow + fr[2]) * NBIN, *x3 = spec + (xrow + fr[3]) * NBIN;
        const size_t mrow = ((size_t)b * S + s) * T;
        if (MASK_KIND == DL4SS_MASK_NONE) {
#pragma unroll
            for (int m = 0; m < 8; ++m) {
                const float2 a0 = x0[16 * m + l16], a1 = x1[16 * m + l16], a2 = x2[16 * m + l16], a3 = x3[16 * m + l16];
                pa[m] = cx2{make_float2(a0.x, a2.x), make_float2(a0.y, a2.y)};
                pb[m] = cx2{make_float2(a1.x, a3.x), make_float2(a1.y, a3.y)};
            }
            if (l16 == 0) {
                pa_n = make_float2(x0[128].x, x2[128].x);
                pb_n = make_float2(x1[128].x, x3[128].x);
            }
        } else if (MASK_KIND == DL4SS_MASK_REAL) {
            const float *m0 = mask + (mrow + fr[0]) * NBIN, *m1 = mask + (mrow + fr[1]) * NBIN;
            const float *m2 = mask + (mrow + fr[2]) * NBIN, *m3 = mask + (mrow + fr[3]) * NBIN;
#pragma unroll
            for (int m = 0; m < 8; ++m) {
                const float2 a0 = x0[16 * m + l16], a1 = x1[16 * m + l16], a2 = x2[16 * m + l16], a3 = x3[16 * m + l16];
                const float k0 = __ldcs(m0 + 16 * m + l16), k1 = __ldcs(m1 + 16 * m + l16);
                const float k2 = __ldcs(m2 + 16 * m + l16), k3 = __ldcs(m3 + 16 * m + l16);
                pa[m] = cx2{make_float2(k0 * a0.x, k2 * a2.x), make_float2(k0 * a0.y, k2 * a2.y)};
                pb[m] = cx2{make_float2(k1 * a1.x, k3 * a3.x), make_float2(k1 * a1.y, k3 * a3.y)};
            }
            if (l16 == 0) {
                pa_n = make_float2(m0[128] * x0[128].x, m2[128] * x2[128].x);
                pb_n = make_float2(m1[128] * x1[128].x, m3[128] * x3[128].x);
            }
        } else {
            const float2 *cm = reinterpret_cast<const float2 *>(mask);
            const float2 *m0 = cm + (mrow + fr[0]) * NBIN, *m1 = cm + (mrow + fr[1]) * NBIN;
            const float2 *m2 = cm + (mrow + fr[2]) * NBIN, *m3 = cm + (mrow + fr[3]) * NBIN;
#pragma unroll
            for (int m = 0; m < 8; ++m) {
                const float2 a0 = x0[16 * m + l16], a1 = x1[16 * m + l16], a2 = x2[16 * m + l16], a3 = x3[16 * m + l16];
                const float2 k0 = m0[16 * m + l16], k1 = m1[16 * m + l16], k2 = m2[16 * m + l16], k3 = m3[16 * m + l16];
                // reference order: re = Mr*Xr - Mi*Xi ; im = Mr*Xi + Mi*Xr
                pa[m] = cx2{make_float2(k0.x * a0.x - k0.y * a0.y, k2.x * a2.x - k2.y * a2.y),
                            make_float2(k0.x * a0.y + k0.y * a0.x, k2.x * a2.y + k2.y * a2.x)};
                pb[m] = cx2{make_float2(k1.x * a1.x - k1.y * a1.y, k3.x * a3.x - k3.y * a3.y),
                            make_float2(k1.x * a1.y + k1.y * a1.x, k3.x * a3.y + k3.y * a3.x)};
            }
            if (l16 == 0) {
                const float2 a0 = x0[128], a1 = x1[128], a2 = x2[128], a3 = x3[128];
                const float2 k0 = m0[128], k1 = m1[128], k2 = m2[128], k3 = m3[128];
                pa_n = make_float2(k0.x * a0.x - k0.y * a0.y, k2.x * a2.x - k2.y * a2.y);
                pb_n = make_float2(k1.x * a1.x - k1.y * a1.y, k3.x * a3.x - k3.y * a3.y);
            }
        }
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
        __syncwarp();     // the group's transpose buffer is reused below as its parking slot
    } else {
#pragma unroll
        for (int n1 = 0; n1 < 16; ++n1) v[n1] = cx2{make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
    }
    // windowed samples 16*n1 + l16 of the four frames (absent frames: zero): f0 = re.x, f1 = im.x, f2 = re.y, f3 = im.y
    const float z0 = fv[0] ? 1.f : 0.f, z1 = fv[1] ? 1.f : 0.f, z2 = fv[2] ? 1.f : 0.f, z3 = fv[3] ? 1.f : 0.f;
#pragma unroll
    for (int n1 = 0; n1 < 16; ++n1) {
        const float w = win[16 * n1 + l16];
        v[n1].re.x *= w * z0; v[n1].im.x *= w * z1; v[n1].re.y *= w * z2; v[n1].im.y *= w * z3;
    }
    // park what this group's frames add to the previous group's three open blocks
    float *ex = reinterpret_cast<float *>(xch + g * DL4SS_XCH2_FLOAT4);      // [3][64], the (now idle) transpose buffer
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const int j = 16 * a + l16;
        ex[j] = v[a].re.x;                                               // f0[j]
        ex[64 + j] = v[4 + a].re.x + v[a].im.x;                          // f0[64+j] + f1[j]
        ex[128 + j] = v[8 + a].re.x + v[4 + a].im.x + v[a].re.y;         // f0[128+j] + f1[64+j] + f2[j]
    }
    // consumed by the group before: __syncwarp inside a warp; across warps warp w-1 waits for warp w (named barriers)
    const int warp = tid >> 5;
    __syncwarp();
    if (warp > 0) asm volatile("bar.arrive %0, 64;" ::"r"(warp + 1) : "memory");
    if (warp < STFT_THREADS / 32 - 1) asm volatile("bar.sync %0, 64;" ::"r"(warp + 2) : "memory");
    if (g >= K6Q_GROUPS) return;

    const float *nx = reinterpret_cast<const float *>(xch + (g + 1) * DL4SS_XCH2_FLOAT4);
    float *o = out + ((size_t)b * S + s) * Lout;
    // padded block pb = 4q+3+i, i = 0..3 -> output block k = pb - 2
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int pbk = 4 * q + 3 + i;
        const int k = pbk - 2;
        if (k < 0 || k > T - 2) continue;
        // contributing frames pbk-3 .. pbk: all four exist in the interior
        const bool full = (pbk - 3 >= 0) && (pbk <= T - 1);
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const int j = 16 * a + l16;
            float acc;
            if (i == 0) acc = v[12 + a].re.x + v[8 + a].im.x + v[4 + a].re.y + v[a].im.y;
            else if (i == 1) acc = v[12 + a].im.x + v[8 + a].re.y + v[4 + a].im.y + nx[j];
            else if (i == 2) acc = v[12 + a].re.y + v[8 + a].im.y + nx[64 + j];
            else acc = v[12 + a].im.y + nx[128 + j];
            float inv;
            if (full) inv = inv_full[j];
            else {
                float e = 0.f;
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const int t = pbk - r;
                    if (t >= 0 && t < T) { const float wt = __ldg(window + 64 * r + j); e += wt * wt; }
                }
                inv = (e > 1.17549435e-38f) ? 1.0f / e : 1.0f;
            }
            // win holds w/N; librosa divides the overlap-added w*frame by sum w^2
            K6_STORE(o + (size_t)k * 64 + j, acc * inv);
        }
    }
}

}  // namespace dl4ss

#include "stft_staged.cuh"

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
    const size_t smem = K1_GROUPS * DL4SS_XCH2_FLOAT4 * sizeof(float4) + 256 * sizeof(float2) + NFFT * sizeof(float);
    const int pf_dist = 4 * sm_count();          // CTAs resident at once (4 per SM: registers / shared memory; forcing 5 with 96 registers spills and measured 64 % against 69 %)
    const bool h128 = (hop == NFFT / 2);
    cudaStream_t st = (cudaStream_t)stream;
    { int rc = ensure_twiddles(st); if (rc) return rc; }
    const long long grid = (long long)B * tiles;
    DL4SS_CHECK_ARG(grid < (1ll << 31), "stft_feat: grid too large");
    if (h128 && staged_k1_enabled() && (((uintptr_t)wav) & 15) == 0) {
        // persistent CTAs fed by bulk copies (stft_staged.cuh)
        const size_t esz = (wav_dtype == DL4SS_WAV_F32) ? sizeof(float) : sizeof(double);
        const char *wav_end = (const char *)wav + (size_t)B * L * esz;
#define LAUNCH_K1S(WT, FM, CP)                                                                                  \
        do {                                                                                                    \
            auto kern = stft256_staged_kernel<WT, FM, CP>;                                                      \
            const size_t sm = K1S_STAGES * K1Stage<WT>::BYTES + K1_GROUPS * DL4SS_XCH2_FLOAT4 * sizeof(float4) + \
                              256 * sizeof(float2) + NFFT * sizeof(float) + 2 * K1S_STAGES * sizeof(uint64_t) + 128; \
            DL4SS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));       \
            int per_sm = 0;                                                                                     \
            DL4SS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, K1S_THREADS, sm));          \
            long long g = (long long)(per_sm > 0 ? per_sm : 1) * sm_count();                                    \
            if (g > grid) g = grid;                                                                             \
            kern<<<(unsigned)g, K1S_THREADS, sm, st>>>((const WT *)wav, L, T, tiles, (int)grid, wav_end, window, eps, conj, \
                                                       feat_out, (float2 *)cplx_out);                          \
        } while (0)
#define DISPATCH_K1S(WT)                                                                                        \
        do {                                                                                                    \
            if (cplx_out) {                                                                                     \
                if (feat_mode == DL4SS_FEAT_NONE) LAUNCH_K1S(WT, DL4SS_FEAT_NONE, true);                        \
                else if (feat_mode == DL4SS_FEAT_ABS) LAUNCH_K1S(WT, DL4SS_FEAT_ABS, true);                     \
                else LAUNCH_K1S(WT, DL4SS_FEAT_LOG, true);                                                      \
            } else {                                                                                            \
                if (feat_mode == DL4SS_FEAT_ABS) LAUNCH_K1S(WT, DL4SS_FEAT_ABS, false);                         \
                else LAUNCH_K1S(WT, DL4SS_FEAT_LOG, false);                                                     \
            }                                                                                                   \
        } while (0)
        if (wav_dtype == DL4SS_WAV_F32) DISPATCH_K1S(float);
        else DISPATCH_K1S(double);
#undef DISPATCH_K1S
#undef LAUNCH_K1S
        DL4SS_LAUNCH_CHECK("stft256_staged_kernel");
        return DL4SS_OK;
    }
    // persistent form: one resident set of CTAs (4 per SM) walks the items
    const bool pers = persistent_k1_enabled() && grid > pf_dist;
    const long long grid_k1 = pers ? pf_dist : grid;
    const int pf_k1 = pf_dist;
#define LAUNCH_K1(WT, FM, CP)                                                                                  \
    do {                                                                                                        \
        if (h128) {                                                                                             \
            DL4SS_CUDA(cudaFuncSetAttribute(stft256_kernel<WT, FM, CP, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
            stft256_kernel<WT, FM, CP, 2><<<(unsigned)grid_k1, K1_THREADS, smem, st>>>(                         \
                (const WT *)wav, L, hop, T, tiles, (int)grid, window, eps, conj, pf_k1, feat_out, (float2 *)cplx_out);   \
        } else if (hop == NFFT / 4) {                                                                           \
            DL4SS_CUDA(cudaFuncSetAttribute(stft256_kernel<WT, FM, CP, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
            stft256_kernel<WT, FM, CP, 4><<<(unsigned)grid_k1, K1_THREADS, smem, st>>>(                         \
                (const WT *)wav, L, hop, T, tiles, (int)grid, window, eps, conj, pf_k1, feat_out, (float2 *)cplx_out);   \
        } else {                                                                                                \
            DL4SS_CUDA(cudaFuncSetAttribute(stft256_kernel<WT, FM, CP, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
            stft256_kernel<WT, FM, CP, 0><<<(unsigned)grid_k1, K1_THREADS, smem, st>>>(                         \
                (const WT *)wav, L, hop, T, tiles, (int)grid, window, eps, conj, pf_k1, feat_out, (float2 *)cplx_out);   \
        }                                                                                                       \
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
        const size_t smem_h = STFT_GROUPS * DL4SS_XCH2_FLOAT4 * sizeof(float4) + 256 * sizeof(float2) +
                              (size_t)STFT_GROUPS * (NFFT / 2) * sizeof(float2) + NFFT * sizeof(float);
        const int pf_dist = 2 * sm_count();      // CTAs resident at once (2 per SM: 128 registers x 256 threads); 1, 3 and 4 waves
                                                 // ahead measured equal / slower (69.0 / 68.5 / 67.2 % at 5 s x 4096)
        const long long grid_h = (long long)B * ((S + 1) / 2) * tiles_h;
        DL4SS_CHECK_ARG(grid_h < (1ll << 31), "mask_istft: grid too large");
        if (staged_enabled() && (((uintptr_t)spec) & 15) == 0 && (((uintptr_t)mask) & 15) == 0) {
            // persistent CTAs fed by bulk copies (stft_staged.cuh)
            const int pairs_tile = (mask_kind == DL4SS_MASK_COMPLEX) ? K6Cfg<DL4SS_MASK_COMPLEX>::PAIRS : K6Cfg<DL4SS_MASK_REAL>::PAIRS;
            const int tiles_s = cdiv(T - 1, 2 * pairs_tile);         // a tile owns 2*PAIRS hop blocks
            const long long items = (long long)B * ((S + 1) / 2) * tiles_s;
            DL4SS_CHECK_ARG(items < (1ll << 31), "mask_istft: too many work items");
            const char *spec_end = (const char *)spec + (size_t)B * (mask_kind == DL4SS_MASK_NONE ? S : 1) * T * NBIN * 8;
            const char *mask_end = mask ? (const char *)mask + (size_t)B * S * T * NBIN * (mask_kind == DL4SS_MASK_COMPLEX ? 8 : 4) : nullptr;
            const long long g = items < sm_count() ? items : sm_count();
#define LAUNCH_K6S(KIND)                                                                                        \
            do {                                                                                                \
                auto kern = istft_h128_staged_kernel<KIND>;                                                     \
                const size_t sm = K6Cfg<KIND>::SMEM;                                                            \
                DL4SS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));   \
                kern<<<(unsigned)g, K6Cfg<KIND>::THREADS, sm, st0>>>(mask, (const float2 *)spec, S, T, tiles_s, (int)items, \
                                                                     mask_end, spec_end, window, wav_out);      \
            } while (0)
            if (mask_kind == DL4SS_MASK_NONE) LAUNCH_K6S(DL4SS_MASK_NONE);
            else if (mask_kind == DL4SS_MASK_REAL) LAUNCH_K6S(DL4SS_MASK_REAL);
            else LAUNCH_K6S(DL4SS_MASK_COMPLEX);
#undef LAUNCH_K6S
            DL4SS_LAUNCH_CHECK("istft_h128_staged_kernel");
            return DL4SS_OK;
        }
#define LAUNCH_H128(KIND)                                                                                       \
        do {                                                                                                    \
            DL4SS_CUDA(cudaFuncSetAttribute(istft_h128_kernel<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_h)); \
            istft_h128_kernel<KIND><<<(unsigned)grid_h, STFT_THREADS, smem_h, st0>>>(                           \
                mask, (const float2 *)spec, S, T, tiles_h, pf_dist, window, wav_out);                                    \
        } while (0)
        if (mask_kind == DL4SS_MASK_NONE) LAUNCH_H128(DL4SS_MASK_NONE);
        else if (mask_kind == DL4SS_MASK_REAL) LAUNCH_H128(DL4SS_MASK_REAL);
        else LAUNCH_H128(DL4SS_MASK_COMPLEX);
#undef LAUNCH_H128
        DL4SS_LAUNCH_CHECK("istft_h128_kernel");
        return DL4SS_OK;
    }
    if (hop == NFFT / 4 && T >= 4) {        // hop 64: four frames of one source per group, overlap-add in registers
        { int rc = ensure_twiddles(st0); if (rc) return rc; }
        const int qmax = (T - 3) / 4;                         // last group with an output block (4q+3 <= T)
        const int ngroups = qmax + 2;                         // q = -1 .. qmax
        const int tiles_q = cdiv(ngroups, K6Q_GROUPS);
        const size_t smem_q = STFT_GROUPS * DL4SS_XCH2_FLOAT4 * sizeof(float4) + 256 * sizeof(float2) + (NFFT + 64) * sizeof(float);
        const long long grid_q = (long long)B * S * tiles_q;
        DL4SS_CHECK_ARG(grid_q < (1ll << 31), "mask_istft: grid too large");
#define LAUNCH_H64(KIND)                                                                                        \
        do {                                                                                                    \
            DL4SS_CUDA(cudaFuncSetAttribute(istft_h64_kernel<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_q)); \
            istft_h64_kernel<KIND><<<(unsigned)grid_q, STFT_THREADS, smem_q, st0>>>(                            \
                mask, (const float2 *)spec, S, T, tiles_q, window, wav_out);                                    \
        } while (0)
        if (mask_kind == DL4SS_MASK_NONE) LAUNCH_H64(DL4SS_MASK_NONE);
        else if (mask_kind == DL4SS_MASK_REAL) LAUNCH_H64(DL4SS_MASK_REAL);
        else LAUNCH_H64(DL4SS_MASK_COMPLEX);
#undef LAUNCH_H64
        DL4SS_LAUNCH_CHECK("istft_h64_kernel");
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
