"""BSS-Eval v3 `bss_eval_sources` restated (oracle, test infra only).

Reference call site: Torch_multi/bss_test.py:5,55 (`from separation import bss_eval_sources`,
i.e. mir_eval.separation, un-vendored, version unpinned; absent from this image).
Published algorithm restated (Vincent, Gribonval, Fevotte 2006; mir_eval.separation):
  for every (estimate j, reference i): project the estimate on the span of all references
  delayed by 0..flen-1 samples (flen=512, least squares via the Toeplitz Gram matrix of
  reference auto/cross-correlations computed with FFTs), split into s_target / e_interf /
  e_artif, SDR = 10log10(|s_target|^2 / |e_interf+e_artif|^2), SIR, SAR; choose the permutation
  with the best mean SIR.
**Parity unpinned** (no mir_eval here, no reference fixtures): pinned by identities in
tests/test_oracle_bss.py (scaled/filtered copy -> very high SDR; equal-power uncorrelated mix
-> ~0 dB; permutation recovered).  It is applied identically to oracle and CUDA outputs, so
the 0.01 dB criterion compares like with like.
"""
import itertools
import numpy as np
from scipy.linalg import toeplitz
from scipy.signal import fftconvolve


def _project(reference_sources, estimated_source, flen):
    nsrc, nsampl = reference_sources.shape
    reference_sources = np.hstack((reference_sources, np.zeros((nsrc, flen - 1))))
    estimated_source = np.hstack((estimated_source, np.zeros(flen - 1)))
    n_fft = int(2 ** np.ceil(np.log2(nsampl + flen - 1.)))
    sf = np.fft.rfft(reference_sources, n=n_fft, axis=1)
    sef = np.fft.rfft(estimated_source, n=n_fft)
    G = np.zeros((nsrc * flen, nsrc * flen))
    for i in range(nsrc):
        for j in range(nsrc):
            ssf = sf[i] * np.conj(sf[j])
            ssf = np.real(np.fft.irfft(ssf))
            ss = toeplitz(np.hstack((ssf[0], ssf[-1:-flen:-1])), r=ssf[:flen])
            G[i * flen: (i + 1) * flen, j * flen: (j + 1) * flen] = ss
            G[j * flen: (j + 1) * flen, i * flen: (i + 1) * flen] = ss.T
    D = np.zeros(nsrc * flen)
    for i in range(nsrc):
        ssef = sf[i] * np.conj(sef)
        ssef = np.real(np.fft.irfft(ssef))
        D[i * flen: (i + 1) * flen] = np.hstack((ssef[0], ssef[-1:-flen:-1]))
    try:
        C = np.linalg.solve(G, D).reshape(flen, nsrc, order='F')
    except np.linalg.LinAlgError:
        C = np.linalg.lstsq(G, D, rcond=None)[0].reshape(flen, nsrc, order='F')
    sproj = np.zeros(nsampl + flen - 1)
    for i in range(nsrc):
        sproj += fftconvolve(C[:, i], reference_sources[i])[:nsampl + flen - 1]
    return sproj


def _bss_decomp_mtifilt(reference_sources, estimated_source, j, flen):
    nsampl = estimated_source.size
    # s_target = s_true + e_spat in mir_eval's notation: the projection on the target alone
    s_target = _project(reference_sources[j, np.newaxis, :], estimated_source, flen)
    e_interf = _project(reference_sources, estimated_source, flen) - s_target
    e_artif = -s_target - e_interf
    e_artif[:nsampl] += estimated_source
    return s_target, e_interf, e_artif


def _safe_db(num, den):
    if den == 0:
        return np.inf
    return 10 * np.log10(num / den)


def _bss_source_crit(s_target, e_interf, e_artif):
    s_filt = s_target
    sdr = _safe_db(np.sum(s_filt ** 2), np.sum((e_interf + e_artif) ** 2))
    sir = _safe_db(np.sum(s_filt ** 2), np.sum(e_interf ** 2))
    sar = _safe_db(np.sum((s_filt + e_interf) ** 2), np.sum(e_artif ** 2))
    return sdr, sir, sar


def bss_eval_sources(reference_sources, estimated_sources, compute_permutation=True, flen=512):
    """-> (sdr[nsrc], sir[nsrc], sar[nsrc], perm[nsrc]) like mir_eval.separation.bss_eval_sources."""
    reference_sources = np.atleast_2d(np.asarray(reference_sources, dtype=np.float64))
    estimated_sources = np.atleast_2d(np.asarray(estimated_sources, dtype=np.float64))
    if reference_sources.shape != estimated_sources.shape:
        raise ValueError('shape mismatch')
    nsrc = estimated_sources.shape[0]
    if compute_permutation:
        sdr = np.empty((nsrc, nsrc)); sir = np.empty((nsrc, nsrc)); sar = np.empty((nsrc, nsrc))
        for jest in range(nsrc):
            for jtrue in range(nsrc):
                s_true, e_interf, e_artif = _bss_decomp_mtifilt(
                    reference_sources, estimated_sources[jest], jtrue, flen)
                sdr[jest, jtrue], sir[jest, jtrue], sar[jest, jtrue] = _bss_source_crit(
                    s_true, e_interf, e_artif)
        perms = list(itertools.permutations(list(range(nsrc))))
        mean_sir = np.empty(len(perms))
        dum = np.arange(nsrc)
        for (i, perm) in enumerate(perms):
            mean_sir[i] = np.mean(sir[perm, dum])
        popt = perms[int(np.argmax(mean_sir))]
        idx = (popt, dum)
        return sdr[idx], sir[idx], sar[idx], np.asarray(popt)
    sdr = np.empty(nsrc); sir = np.empty(nsrc); sar = np.empty(nsrc)
    for j in range(nsrc):
        s_true, e_interf, e_artif = _bss_decomp_mtifilt(reference_sources, estimated_sources[j], j, flen)
        sdr[j], sir[j], sar[j] = _bss_source_crit(s_true, e_interf, e_artif)
    return sdr, sir, sar, np.arange(nsrc)


def pcm16_roundtrip(x):
    """soundfile default WAV subtype emulation: the reference writes PCM16 wavs and re-reads them
    before SDR (TDAA_beta/main_run_sstune_EvalVer.py:72 ; Torch_multi/bss_test.py:34)."""
    x = np.asarray(x, dtype=np.float64)
    q = np.round(np.clip(x, -1.0, 1.0 - 2.0 ** -15) * 32768.0)
    return q / 32768.0
