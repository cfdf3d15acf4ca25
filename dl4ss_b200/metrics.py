"""On-device BSS-Eval: SDR / SIR / SAR of separated waveforms without leaving the GPU.

The reference scores a batch by writing every separated and clean waveform to PCM16 wav files
(TDAA_beta/main_run_sstune_EvalVer.py:45-74), reading them back and calling
mir_eval.separation.bss_eval_sources per utterance on the CPU (Torch_multi/bss_test.py:12-61): by far the
slowest part of its evaluation loop.  `bss_eval_sources_batch` scores the whole batch where the waveforms
already are:

  1. `dl4ss_xcorr_f64` (csrc/xcorr.cu): all reference auto/cross-correlations over lags -(flen-1)..flen-1 and
     all reference x estimate correlations over lags 0..flen-1, summed directly in fp64 (mir_eval: 2^17-point FFTs).
  2. The Gram matrix G of the flen-tap projection (block Toeplitz, S*flen square) is gathered from (1) and
     Cholesky-factored in fp64 (torch.linalg = cuSOLVER: a library call, as numpy.linalg.solve is in mir_eval).
  3. Nothing is filtered or convolved: with y = L^-1 D the energies BSS-Eval needs are quadratic forms,
         |P|^2 = |L^-1 D|^2              (projection of the estimate on ALL delayed references)
         |s_t|^2 = |L_tt^-1 D_t|^2        (projection on reference t alone)
         |e_interf|^2 = |P|^2 - |s_t|^2,  |e_artif|^2 = |est|^2 - |P|^2,  |e_interf + e_artif|^2 = |est|^2 - |s_t|^2
     (exact identities of least-squares projections: <est, P> = |P|^2, <P, s_t> = |s_t|^2), so
         SDR = 10 log10(|s_t|^2 / (|est|^2 - |s_t|^2)),  SIR = 10 log10(|s_t|^2 / (|P|^2 - |s_t|^2)),
         SAR = 10 log10(|P|^2 / (|est|^2 - |P|^2)).
  4. The permutation with the best mean SIR is picked per utterance, as mir_eval does.
"""
import itertools

import torch

from . import _lib


def xcorr_f64(x, y, nlags, lag0):
    """x [B,Sx,N], y [B,Sy,N] CUDA fp32 -> out [B,Sx,Sy,nlags] fp64, out[..., k] = sum_m x[m] * y[m + lag0 + k]."""
    lib = _lib.load()
    B, Sx, N = x.shape
    Sy = y.shape[1]
    if y.shape[0] != B or y.shape[2] != N:
        raise RuntimeError('xcorr_f64: x %s and y %s must share batch and length' % (tuple(x.shape), tuple(y.shape)))
    out = torch.empty(B, Sx, Sy, nlags, device=x.device, dtype=torch.float64)
    rc = lib.dl4ss_xcorr_f64(_lib.ptr(x, name='x'), _lib.ptr(y, name='y'), B, Sx, Sy, N, nlags, lag0,
                             _lib.ptr(out, torch.float64), _lib.stream())
    _lib.check(rc, 'dl4ss_xcorr_f64')
    return out


def pcm16_roundtrip(x):
    """What the reference's sf.write / sf.read pair does to a waveform before scoring
    (TDAA_beta/main_run_sstune_EvalVer.py:72, Torch_multi/bss_test.py:34): PCM16 quantisation."""
    return torch.round(torch.clamp(x, -1.0, 1.0 - 2.0 ** -15) * 32768.0) / 32768.0


def bss_eval_sources_batch(reference_sources, estimated_sources, compute_permutation=True, flen=512):
    """reference_sources, estimated_sources [B,S,N] CUDA fp32 -> (sdr, sir, sar, perm), each [B,S]
    (fp64, fp64, fp64, int64), row b equal to mir_eval.separation.bss_eval_sources(ref[b], est[b])."""
    ref = reference_sources.contiguous()
    est = estimated_sources.contiguous()
    if ref.shape != est.shape or ref.dim() != 3:
        raise ValueError('bss_eval_sources_batch: reference and estimate must both be [B,S,N]')
    rr = xcorr_f64(ref, ref, 2 * flen - 1, -(flen - 1))          # [B,S,S,2flen-1]: r_ij[tau] at index flen-1-tau
    rd = xcorr_f64(ref, est, flen, 0)                            # [B,S,Se,flen]: D_i[a] for estimate e
    ee = (est.double() ** 2).sum(-1)                             # |est_e|^2 [B,S]
    return bss_from_correlations(rr, rd, ee, flen, compute_permutation)


def bss_from_correlations(rr, rd, ee, flen=512, compute_permutation=True):
    """Steps 2-4 of the module docstring (device-agnostic tensor algebra in fp64).
    rr [B,S,S,2flen-1] = sum_m ref_i[m] ref_j[m - (flen-1) + k] ; rd [B,S,S,flen] = sum_m ref_i[m] est_e[m + a] ;
    ee [B,S] = |est_e|^2."""
    B, S = ee.shape
    dev = ee.device
    a = torch.arange(flen, device=dev)
    toe = (flen - 1) + a[:, None] - a[None, :]                   # G_ij[a,c] = rr[i,j, flen-1 + a - c]
    G = rr[:, :, :, toe]                                         # [B,S,S,flen,flen]
    Gd = torch.diagonal(G, dim1=1, dim2=2).permute(0, 3, 1, 2)   # [B,S,flen,flen]: single-reference systems
    Gf = G.permute(0, 1, 3, 2, 4).reshape(B, S * flen, S * flen)
    D = rd.permute(0, 1, 3, 2)                                   # [B,S(i),flen,Se]
    Lf, info_f = torch.linalg.cholesky_ex(Gf)
    Ld, info_d = torch.linalg.cholesky_ex(Gd.contiguous())
    yf = torch.linalg.solve_triangular(Lf, D.reshape(B, S * flen, S), upper=False)
    yd = torch.linalg.solve_triangular(Ld, D.contiguous(), upper=False)
    P = (yf ** 2).sum(1)                                         # [B,Se]
    st = (yd ** 2).sum(2)                                        # [B,St,Se]
    bad = (info_f != 0)[:, None] | (info_d != 0).any(1)[:, None]  # a silent / duplicated reference: no projection
    ee_ = ee[:, None, :]
    P_ = P[:, None, :]
    ten = 10.0 / torch.log(torch.tensor(10.0, dtype=torch.float64, device=dev))

    def db(num, den):
        r = ten * (torch.log(num) - torch.log(den.clamp_min(0)))
        return torch.where(den <= 0, torch.full_like(r, float('inf')), r)

    sdr = db(st, ee_ - st)                                       # [B, true t, est e]
    sir = db(st, P_ - st)
    sar = db(P_, ee_ - P_).expand(B, S, S)
    nan = torch.full_like(sdr, float('nan'))
    sdr, sir, sar = [torch.where(bad[:, :, None], nan, v) for v in (sdr, sir, sar)]
    dum = torch.arange(S, device=dev)
    if not compute_permutation:
        return sdr[:, dum, dum], sir[:, dum, dum], sar[:, dum, dum], dum.expand(B, S).clone()
    perms = torch.tensor(list(itertools.permutations(range(S))), device=dev)          # [P,S]: est perm[t] <-> true t
    mean_sir = sir[:, dum[None, :], perms].mean(-1)                                    # [B,P]
    best = perms[torch.argmax(mean_sir, dim=1)]                                        # [B,S]
    bi = torch.arange(B, device=dev)[:, None]
    return sdr[bi, dum[None, :], best], sir[bi, dum[None, :], best], sar[bi, dum[None, :], best], best
