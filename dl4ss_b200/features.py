"""Batched feature step and reconstruction on the GPU (host side of K1 / K6).

`stft_features` produces, for a whole batch in one launch, exactly the arrays the reference's
generators build one utterance at a time with librosa on the CPU
(TDAA_beta/predata_fromList.py:166-200, TDAA_beta/predata_fromList_cRM_123.py:215-255):
    mix_feas  [B,T,F]   float32   |STFT|  or  log(|STFT| + eps)
    mix_mag   [B,T,F,2] float32   convert2(STFT)  (re, im)
    mix_phase [B,T,F]   complex64 (a view of mix_mag)
`mask_istft` is the reconstruction half of `bss_eval` / `bss_eval_cRM`
(TDAA_beta/main_run_sstune_EvalVer.py:55-65, ...cRM_EvalVer.py:96-99) without the wav files.
"""
import math

import numpy as np
import torch

from . import _lib
from . import config

_window_cache = {}


def window_tensor(window, n_fft, device):
    """fp32 device table of the analysis/synthesis window.

    'hann' = periodic Hann (librosa default, scipy get_window fftbins=True); 'sine' = the table
    of Torch_multi/config.py:240; 'sqrt_hann'; or any sequence of n_fft taps (config.WINDOWS)."""
    if torch.is_tensor(window):
        w = window.to(device=device, dtype=torch.float32).contiguous()
        if w.numel() != n_fft:
            raise ValueError('window must have n_fft taps')
        return w
    key = None
    if isinstance(window, str):
        key = (window, n_fft, str(device))
        if key in _window_cache:
            return _window_cache[key]
        i = np.arange(n_fft, dtype=np.float64)
        if window == 'hann':
            w = 0.5 - 0.5 * np.cos(2.0 * np.pi * i / n_fft)
        elif window == 'sine':
            w = np.sin(i * np.pi / n_fft)
        elif window == 'sqrt_hann':
            w = np.sqrt(0.5 - 0.5 * np.cos(2.0 * np.pi * i / n_fft))
        elif window in ('ones', 'boxcar'):
            w = np.ones(n_fft)
        else:
            raise ValueError('unknown window %r' % (window,))
    else:
        w = np.asarray(window, dtype=np.float64)
        if w.shape != (n_fft,):
            raise ValueError('window must have n_fft taps')
    t = torch.from_numpy(w.astype(np.float32)).to(device)
    if key is not None:
        _window_cache[key] = t
    return t


def num_frames(L, hop):
    return 1 + L // hop


def stft_features(wav, n_fft=None, hop=None, window='hann', feat='abs', eps=None, want_complex=True,
                  conj=False, out_feat=None, out_cplx=None):
    """wav [B,L] CUDA float32/float64 -> (feat [B,T,F] or None, cplx [B,T,F,2] or None)."""
    lib = _lib.load()
    n_fft = config.FRAME_LENGTH if n_fft is None else n_fft
    hop = config.FRAME_SHIFT if hop is None else hop
    if wav.dim() == 1:
        wav = wav.unsqueeze(0)
    if wav.dtype == torch.float64:
        wdt = _lib.WAV_F64
    elif wav.dtype == torch.float32:
        wdt = _lib.WAV_F32
    else:
        raise RuntimeError('stft_features: wav must be float32 or float64')
    B, L = wav.shape
    T, F = num_frames(L, hop), n_fft // 2 + 1
    mode = {None: _lib.FEAT_NONE, 'none': _lib.FEAT_NONE, 'abs': _lib.FEAT_ABS, 'log': _lib.FEAT_LOG}[feat]
    dev = wav.device
    feat_t = None
    if mode != _lib.FEAT_NONE:
        feat_t = out_feat if out_feat is not None else torch.empty(B, T, F, device=dev, dtype=torch.float32)
    cplx_t = None
    if want_complex:
        cplx_t = out_cplx if out_cplx is not None else torch.empty(B, T, F, 2, device=dev, dtype=torch.float32)
    w = window_tensor(window, n_fft, dev)
    rc = lib.dl4ss_stft_feat(_lib.ptr(wav, None, 'wav'), wdt, B, L, n_fft, hop, _lib.ptr(w), mode,
                             float(config.EPS_LOG if eps is None else eps), int(bool(conj)),
                             _lib.ptr(feat_t), _lib.ptr(cplx_t), _lib.stream())
    _lib.check(rc, 'dl4ss_stft_feat')
    return feat_t, cplx_t


def mask_istft(mask, spec, hop=None, window='hann', n_fft=None, out=None):
    """Masked reconstruction.

    mask [B,S,T,F]   (real)    x spec [B,T,F,2] mixture
    mask [B,S,T,F,2] (complex) x spec [B,T,F,2] mixture
    mask None                    spec [B,S,T,F,2] per-source spectra
    -> wav [B,S,hop*(T-1)] float32 (librosa.istft semantics, centre trimmed)."""
    lib = _lib.load()
    hop = config.FRAME_SHIFT if hop is None else hop
    if torch.is_complex(spec):
        spec = torch.view_as_real(spec)
    if mask is None:
        B, S, T, F, _ = spec.shape
        kind = _lib.MASK_NONE
    else:
        B, S, T, F = mask.shape[:4]
        kind = _lib.MASK_COMPLEX if mask.dim() == 5 else _lib.MASK_REAL
        if tuple(spec.shape) != (B, T, F, 2):
            raise RuntimeError('mask_istft: spec must be the mixture spectrum [B,T,F,2]')
    nf = 2 * (F - 1) if n_fft is None else n_fft
    if out is None:
        out = torch.empty(B, S, hop * (T - 1), device=spec.device, dtype=torch.float32)
    w = window_tensor(window, nf, spec.device)
    rc = lib.dl4ss_mask_istft(_lib.ptr(mask, name='mask'), kind, _lib.ptr(spec, name='spec'), B, S, T, nf, hop,
                              _lib.ptr(w), _lib.ptr(out), _lib.stream())
    _lib.check(rc, 'dl4ss_mask_istft')
    return out


def prepare_batch(mix_wav, n_fft=None, hop=None, is_log_spectral=None, window=None, sources=None):
    """GPU counterpart of one `prepare_data('once', ...)` yield for already-mixed waveforms.

    mix_wav [B,L] (CUDA; float64 like the reference's `mix_wav`, or float32).
    Returns the hot-path keys of the reference dict (TDAA_beta/predata_fromList.py:224-234):
      mix_wav, mix_feas, mix_phase (complex64 view), mix_mag ([B,T,F,2]) and, when the clean
      `sources` [B,S,L] are given, `multi_spk_fea` [B,S,T,F] (|STFT| of each source, the MSE
      targets) / `multi_spk_mag` [B,S,T,F,2] (cRM targets).
    The reference recomputes the mixture STFT 2-3 times; here one launch yields all views."""
    is_log = config.IS_LOG_SPECTRAL if is_log_spectral is None else is_log_spectral
    if is_log:
        win = config.WINDOWS if window is None else window
        if isinstance(win, int):            # un-initialised config (Torch_multi/config.py:133): sine table
            win = 'sine'
        feas, _ = stft_features(mix_wav, n_fft, hop, win, 'log', want_complex=False)
        _, cplx = stft_features(mix_wav, n_fft, hop, 'hann', None, want_complex=True)
    else:
        feas, cplx = stft_features(mix_wav, n_fft, hop, 'hann', 'abs', want_complex=True)
    out = {'mix_wav': mix_wav, 'mix_feas': feas, 'mix_mag': cplx, 'mix_phase': torch.view_as_complex(cplx)}
    if sources is not None:
        B, S, L = sources.shape
        f, c = stft_features(sources.reshape(B * S, L), n_fft, hop, 'hann', 'abs', want_complex=True)
        out['multi_spk_fea'] = f.view(B, S, f.shape[1], f.shape[2])
        out['multi_spk_mag'] = c.view(B, S, c.shape[1], c.shape[2], 2)
    return out


def premix(sources, gains_db, lengths=None, keep_sources=True, shifts=None):
    """GPU counterpart of the generators' per-source preprocessing + mixing
    (TDAA_beta/predata_fromList.py:140-177): sources [B,S,L] fp32 CUDA (raw, zero beyond `lengths`), gains_db
    [B,S], lengths int [B,S] or None, shifts int [B,S] or None (the AUGMENT_DATA circular shift, :150-153)
    -> dict(mix_wav [B,L], sources [B,S,L] preprocessed) for `prepare_batch`."""
    lib = _lib.load()
    B, S, L = sources.shape
    dev = sources.device
    gains_db = gains_db.to(device=dev, dtype=torch.float32).contiguous()
    ln = None if lengths is None else lengths.to(device=dev, dtype=torch.int32).contiguous()
    sh = None if shifts is None else shifts.to(device=dev, dtype=torch.int32).contiguous()
    mix = torch.empty(B, L, device=dev, dtype=torch.float32)
    out = torch.empty_like(sources) if keep_sources else None
    rc = lib.dl4ss_premix_shift_fwd(_lib.ptr(sources, name='sources'), _lib.ptr(ln, torch.int32, 'lengths'),
                                    _lib.ptr(sh, torch.int32, 'shifts'), _lib.ptr(gains_db, name='gains_db'), B, S, L,
                                    _lib.ptr(out), _lib.ptr(mix), _lib.stream())
    _lib.check(rc, 'dl4ss_premix_shift_fwd')
    return {'mix_wav': mix, 'sources': out}
