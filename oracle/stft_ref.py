"""librosa-semantics STFT / iSTFT restated in numpy (oracle, test infra only).

Reference call sites (the library itself, librosa 0.4-0.6 era, is not vendored):
  stft : TDAA_beta/predata_fromList.py:166,179,194,199 ; Torch_multi/predata_multiAims.py:168,185,200,205
  istft: TDAA_beta/main_run_sstune_EvalVer.py:64-65 ; TDAA_beta/main_run_sstune_cRM_EvalVer.py:96-99
Algorithm restated (published librosa.core.spectrum behaviour):
  stft : centre=True -> reflect-pad n_fft/2 ; frames at hop ; x periodic Hann (scipy get_window
         fftbins=True) ; FFT in float64 ; keep 1+n_fft/2 bins ; cast complex64 ; shape [F,T],
         T = 1 + L//hop.
  istft: per frame Hermitian-extend, ifft().real, x window, overlap-add into a float32 buffer,
         divide by the window sum-square where > tiny(float32), trim n_fft/2 both sides,
         length hop*(T-1).
Pinned: against outputs of the reference's own numpy stft / istft (Cocktail/software/DL4SS_Keras/
test_stft_istft.py:9-63 -> tests/golden/ref_stft_*.npz, tests/test_oracle_refpin.py; `center=False` is
the restatement of those functions) and against torch.stft/istft + scipy (tests/test_oracle.py).
librosa itself is absent from the image: its values are not compared directly.
"""
import numpy as np

TINY_F32 = float(np.finfo(np.float32).tiny)


def get_window(kind, n_fft):
    """Window table in float64.

    'hann' : scipy.signal.get_window('hann', N, fftbins=True) -- librosa default.
    'sine' : [sin(i*pi/N)] -- Torch_multi/config.py:240 (log-spectral / Keras-era mode).
    'sqrt_hann': sqrt of periodic hann (commented alternative Torch_multi/config.py:242).
    array  : used as-is (must have n_fft taps).
    """
    if isinstance(kind, str):
        i = np.arange(n_fft, dtype=np.float64)
        if kind == 'hann':
            return 0.5 - 0.5 * np.cos(2.0 * np.pi * i / n_fft)
        if kind == 'sine':
            return np.sin(i * np.pi / n_fft)
        if kind == 'sqrt_hann':
            return np.sqrt(0.5 - 0.5 * np.cos(2.0 * np.pi * i / n_fft))
        if kind == 'sqrt_hanning':                      # sqrt(np.hanning(M)): the SYMMETRIC Hann of the reference's own numpy
            return np.sqrt(np.hanning(n_fft))           # stft/istft, Cocktail/software/DL4SS_Keras/test_stft_istft.py:9-10
        if kind == 'ones' or kind == 'boxcar':
            return np.ones(n_fft)
        raise ValueError('unknown window %r' % (kind,))
    w = np.asarray(kind, dtype=np.float64)
    if w.shape != (n_fft,):
        raise ValueError('window must have n_fft taps')
    return w


def num_frames(L, hop):
    return 1 + L // hop


def stft_ref(y, n_fft=256, hop=128, window='hann', conj=False, center=True):
    """y [L] real -> complex64 [F, T] (librosa layout).

    center=False is the reference's own pure-numpy transform (Cocktail/software/DL4SS_Keras/test_stft_istft.py:13-35):
    no padding, frames start at range(0, L - n_fft, hop) (a frame ending exactly at L is NOT taken), rfft of
    window * frame, kept in complex128 as numpy does.  tests/golden/ref_stft_*.npz hold that function's outputs."""
    y = np.asarray(y, dtype=np.float64)
    if y.ndim != 1:
        raise ValueError('stft_ref takes one utterance')
    if not center:
        w = get_window(window, n_fft)
        starts = np.arange(0, y.shape[0] - n_fft, hop)
        idx = np.arange(n_fft)[:, None] + starts[None, :]
        S = np.fft.rfft(y[idx] * w[:, None], axis=0)
        return np.conj(S) if conj else S
    if y.shape[0] <= n_fft // 2:
        raise ValueError('reflect padding needs L > n_fft/2')
    w = get_window(window, n_fft)
    yp = np.pad(y, n_fft // 2, mode='reflect')
    T = 1 + (yp.shape[0] - n_fft) // hop
    idx = np.arange(n_fft)[:, None] + hop * np.arange(T)[None, :]
    frames = yp[idx] * w[:, None]                       # [n_fft, T]
    S = np.fft.fft(frames, axis=0)[:1 + n_fft // 2]
    if conj:                                            # librosa <= 0.5 "match DPWE phase" flavour
        S = np.conj(S)
    return S.astype(np.complex64)


def window_sumsquare(window, T, n_fft, hop):
    w = get_window(window, n_fft).astype(np.float32) ** 2
    n = n_fft + hop * (T - 1)
    x = np.zeros(n, dtype=np.float32)
    for i in range(T):
        s = i * hop
        x[s:s + n_fft] += w
    return x


def istft_ref(S, hop=128, window='hann', center=True):
    """S complex [F, T] -> float32 [hop*(T-1)].

    center=False is the reference's own pure-numpy inverse (Cocktail/software/DL4SS_Keras/test_stft_istft.py:38-63):
    a float64 buffer of T*hop samples, frames n = 0.. placed at range(0, T*hop - n_fft, hop) (so the last n_fft/hop
    frames are never used), irfft * window overlap-added, divided by the summed squared window where that is
    non-zero; nothing is trimmed.  Returns float64 [T*hop]."""
    S = np.asarray(S)
    F, T = S.shape
    n_fft = 2 * (F - 1)
    w = get_window(window, n_fft)
    if not center:
        x = np.zeros(T * hop)
        wsum = np.zeros(T * hop)
        for n, i in enumerate(range(0, T * hop - n_fft, hop)):
            x[i:i + n_fft] += np.fft.irfft(S[:, n]).real * w
            wsum[i:i + n_fft] += w ** 2.0
        pos = wsum != 0
        x[pos] /= wsum[pos]
        return x
    n = n_fft + hop * (T - 1)
    y = np.zeros(n, dtype=np.float32)
    full = np.concatenate([S, np.conj(S[-2:0:-1])], axis=0)     # [n_fft, T]
    yt = np.fft.ifft(full.astype(np.complex128), axis=0).real * w[:, None]
    for i in range(T):
        s = i * hop
        y[s:s + n_fft] = y[s:s + n_fft] + yt[:, i]
    wss = window_sumsquare(window, T, n_fft, hop)
    nz = wss > TINY_F32
    y[nz] /= wss[nz]
    return y[n_fft // 2: n - n_fft // 2]


# ---------------------------------------------------------------- features (a3, a4)
EPS_LOG = float(np.spacing(1))          # TDAA_beta/predata_fromList.py:192


def features_ref(wav, n_fft=256, hop=128, log_spectral=False, window=None, conj=False):
    """One utterance -> dict(mix_feas [T,F] f32, mix_phase [T,F] c64, mix_mag [T,F,2] f32).

    Follows TDAA_beta/predata_fromList.py:188-200 (mix_feas / mix_phase) and
    TDAA_beta/predata_fromList_cRM_123.py:37-41,250-255 (convert2 -> mix_mag).
    The log branch uses `window` (the reference passes config.WINDOWS = sine table there);
    mix_phase / mix_mag always use the default Hann window, as in the reference.
    """
    S = stft_ref(wav, n_fft, hop, 'hann', conj).T              # [T,F]
    if log_spectral:
        Sw = stft_ref(wav, n_fft, hop, window if window is not None else 'sine', conj).T
        feas = np.log(np.abs(Sw) + EPS_LOG)
    else:
        feas = np.abs(S)
    return {
        'mix_feas': feas.astype(np.float32),
        'mix_phase': S,
        'mix_mag': convert2(S),
    }


def convert2(array):
    """complex [T,F] -> float32 [T,F,2]  (TDAA_beta/predata_fromList_cRM_123.py:37-41)."""
    out = np.zeros(array.shape + (2,), dtype=np.float32)
    out[..., 0] = np.real(array)
    out[..., 1] = np.imag(array)
    return out


# ---------------------------------------------------------------- waveform preprocessing (a1)
def preprocess_source(signal, max_len, gain_db):
    """TDAA_beta/predata_fromList.py:140-159: crop, -mean, /max|.|, zero-pad, gain 10^(dB/20)."""
    s = np.array(signal, dtype=np.float64)
    if s.shape[0] > max_len:
        s = s[:max_len]
    s = s - np.mean(s)
    s = s / np.max(np.abs(s))
    if s.shape[0] < max_len:
        s = np.append(s, np.zeros(max_len - s.shape[0]))
    return (10.0 ** (gain_db / 20.0)) * s
