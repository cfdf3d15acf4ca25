"""CPU tests of the boundary: the C-ABI library loads and exports every symbol include/*.h declares
(no compute calls -- there is no GPU here), the ctypes signatures cover the header one to one,
the product path refuses CPU tensors (no fallback), and the host-side mirror of the reference's
module interfaces (names, constructor arguments, state-dict keys, glue reshapes) behaves."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, 'include', 'dl4ss_b200.h')


def header_functions():
    src = open(HEADER).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    names = re.findall(r'\b(dl4ss_[a-z0-9_]+)\s*\(', src)
    return sorted(set(names))


def test_library_exports_every_declared_symbol():
    from dl4ss_b200 import _lib
    lib = _lib.load()
    names = header_functions()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), 'libdl4ss_b200.so does not export %s' % n
    # the ctypes table mirrors the header one to one
    assert sorted(_lib.SIGNATURES) == names
    assert lib.dl4ss_version() >= 100
    assert isinstance(_lib.launch_count(), int)


def test_header_is_plain_c():
    """The boundary is a C ABI: the header must compile as C (no torch / C++ types)."""
    import subprocess
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, 't.c')
        with open(c, 'w') as f:
            f.write('#include "dl4ss_b200.h"\nint main(void){return dl4ss_version()>0?0:1;}\n')
        subprocess.check_call(['gcc', '-std=c99', '-Wall', '-Werror', '-I', os.path.join(ROOT, 'include'),
                               '-c', c, '-o', os.path.join(d, 't.o')])


def test_argument_errors_are_reported_without_a_gpu():
    """Argument validation happens before any CUDA call: status code + dl4ss_last_error()."""
    from dl4ss_b200 import _lib
    lib = _lib.load()
    rc = lib.dl4ss_stft_feat(None, 0, 1, 1000, 256, 128, None, 1, 0.0, 0, None, None, None)
    assert rc == -1 and 'null' in _lib.last_error()
    one = ctypes.c_void_p(16)
    rc = lib.dl4ss_stft_feat(one, 0, 1, 1000, 512, 128, one, 1, 0.0, 0, one, one, None)
    assert rc == -2 and '256' in _lib.last_error()              # DL4SS_EUNSUPPORTED
    rc = lib.dl4ss_stft_feat(one, 0, 1, 100, 256, 128, one, 1, 0.0, 0, one, one, None)
    assert rc == -1                                             # L <= n_fft/2: reflect pad impossible
    rc = lib.dl4ss_mask_istft(None, 1, one, 1, 2, 10, 256, 128, one, one, None)
    assert rc == -1 and 'mask' in _lib.last_error()
    rc = lib.dl4ss_rnn_layer_fwd(7, one, one, None, one, 1, 1, 300, None, None, None, 0, None)
    assert rc == -1 and 'cell' in _lib.last_error()
    assert lib.dl4ss_rnn_workspace_bytes(256, 313, 300, 0) >= 256
    assert lib.dl4ss_split_bf16_bytes(10, 129) == 2 * 10 * 192 * 2
    with pytest.raises(RuntimeError):
        _lib.check(-1, 'x')


def test_no_cpu_fallback():
    import dl4ss_b200 as d
    with pytest.raises(RuntimeError, match='CUDA'):
        d.stft_features(torch.zeros(1, 4000))
    with pytest.raises(RuntimeError, match='CUDA'):
        d.linear_fwd(torch.zeros(2, 3), torch.zeros(4, 3))
    m = d.MIX_SPEECH(129, 10, cell='lstm', num_layers=1)
    with pytest.raises(RuntimeError, match='CUDA'):
        m(torch.zeros(1, 10, 129))
    # nothing under dl4ss_b200/ imports the oracle
    pkg = os.path.join(ROOT, 'dl4ss_b200')
    for f in os.listdir(pkg):
        if f.endswith('.py'):
            assert not re.search(r'^\s*(from|import)\s+oracle', open(os.path.join(pkg, f)).read(), re.M), f


def test_state_dict_keys_match_reference():
    """Checkpoint compatibility: one state-dict per module, the reference's key names."""
    import dl4ss_b200 as d
    from oracle import modules_ref as mr
    d.config.HIDDEN_UNITS, d.config.EMBEDDING_SIZE = 300, 50
    for cplx in (False, True):
        d.config.is_ComlexMask = cplx
        rc = mr.RefConfig(is_ComlexMask=cplx)
        pairs = [(d.MIX_SPEECH(129, 313, cell='lstm', num_layers=4), mr.MIX_SPEECH(rc, 129, 313, 'lstm', 4)),
                 (d.MIX_SPEECH(129, 313, cell='gru', num_layers=2), mr.MIX_SPEECH(rc, 129, 313, 'gru', 2)),
                 (d.ATTENTION(50, 'dot'), mr.ATTENTION(rc, 50, 'dot')),
                 (d.SPEECH_EMBEDDING(101, 50, 2), mr.SPEECH_EMBEDDING(rc, 101, 50, 2)),
                 (d.ADDJUST(600, 50), mr.ADDJUST(rc, 600, 50))]
        for ours, ref in pairs:
            so, sr_ = ours.state_dict(), ref.state_dict()
            assert list(so.keys()) == list(sr_.keys())
            assert all(so[k].shape == sr_[k].shape for k in so)
            ours.load_state_dict(sr_)
    d.config.is_ComlexMask = False
    keys = d.MIX_SPEECH(129, 313).state_dict().keys()
    assert 'layer.weight_ih_l0' in keys and 'layer.weight_hh_l1_reverse' in keys and 'Linear.weight' in keys
    assert d.MIX_SPEECH(129, 313).Linear.weight.shape == (129 * 50, 600)


def test_deferred_embedding_follows_reference_glue():
    """The reshapes the reference applies to MIX_SPEECH's output (EvalVer.py:453-455) work on the
    stand-in, anything else is refused."""
    import dl4ss_b200 as d
    B, T, F, E, S = 3, 7, 129, 50, 2
    de = d.DeferredEmbedding(torch.zeros(B, T, 600), torch.zeros(F * E, 600), torch.zeros(F * E), F, E)
    assert de.size() == (B, T, F, E) and de.size(0) == B and de.dim() == 4 and de.copies == 1
    h5 = de.view(B, 1, T, F, E).expand(B, S, T, F, E).contiguous().view(-1, T, F, E)
    assert h5.shape == (B * S, T, F, E) and h5.copies == S
    with pytest.raises(RuntimeError):
        de.view(B * T, F * E)


def test_top_k_mask_matches_oracle():
    import dl4ss_b200 as d
    from oracle import modules_ref as mr
    torch.manual_seed(0)
    p = torch.rand(6, 101)
    for alpha, k in ((0.5, 2), (-0.5, 2), (0.99, 3), (0.0, 1)):
        assert torch.equal(d.top_k_mask(p, alpha, k), mr.top_k_mask(p, alpha, k))


def test_config_mirrors_reference_globals():
    import dl4ss_b200.config as c
    assert c.FRAME_LENGTH == 256 and c.FRAME_SHIFT == 128 and c.MAX_LEN == 40000      # Torch_multi/config.py:114-130
    assert c.EMBEDDING_SIZE == 50 and c.FRAME_RATE == 8000
    assert c.cRM_k == 10.0 and c.cRM_C == 0.1
    assert c.EPS_LOG == float(np.spacing(1))
    assert len(c.sine_window()) == 256 and c.sine_window()[0] == 0.0


def test_window_tensor_and_frames():
    from dl4ss_b200 import features
    from oracle import stft_ref as sr
    for kind in ('hann', 'sine', 'sqrt_hann'):
        w = features.window_tensor(kind, 256, torch.device('cpu'))
        assert np.allclose(w.numpy(), sr.get_window(kind, 256).astype(np.float32))
    with pytest.raises(ValueError):
        features.window_tensor('blackman', 256, torch.device('cpu'))
    with pytest.raises(ValueError):
        features.window_tensor([1.0] * 5, 256, torch.device('cpu'))
    assert features.num_frames(40000, 128) == 313 and features.num_frames(17040, 128) == 134


def _np_xcorr(x, y, nlags, lag0):
    """out[b,i,j,k] = sum_m x[b,i,m] * y[b,j,m+lag0+k] in float64 (y = 0 outside its support)."""
    B, Sx, N = x.shape
    Sy = y.shape[1]
    pad = nlags + abs(lag0) + 1
    out = np.zeros((B, Sx, Sy, nlags))
    for b in range(B):
        for j in range(Sy):
            yy = np.concatenate([np.zeros(pad), y[b, j].astype(np.float64), np.zeros(pad)])
            for i in range(Sx):
                xi = x[b, i].astype(np.float64)
                for k in range(nlags):
                    out[b, i, j, k] = np.dot(xi, yy[pad + lag0 + k: pad + lag0 + k + N])
    return out


@pytest.mark.parametrize('S,perm', [(2, True), (3, True), (2, False)])
def test_bss_closed_form_equals_oracle(S, perm):
    """dl4ss_b200.metrics.bss_from_correlations (Gram-matrix quadratic forms, no filtering) against the oracle's
    explicit BSS-Eval decomposition (oracle/bss_eval_ref.py) from the same float64 correlations: SDR/SIR/SAR agree
    to 1e-9 dB and the permutation is the same."""
    from oracle import bss_eval_ref as be
    from dl4ss_b200 import metrics
    rng = np.random.RandomState(3)
    B, N, flen = 2, 3000, 32
    ref = np.stack([[np.convolve(rng.randn(N), np.ones(4) / 4, 'same') for _ in range(S)] for _ in range(B)]).astype(np.float32)
    est = (ref[:, ::-1] * 0.7 + 0.25 * ref + 0.2 * rng.randn(B, S, N)).astype(np.float32)
    rr = torch.from_numpy(_np_xcorr(ref, ref, 2 * flen - 1, -(flen - 1)))
    rd = torch.from_numpy(_np_xcorr(ref, est, flen, 0))
    ee = torch.from_numpy((est.astype(np.float64) ** 2).sum(-1))
    sdr, sir, sar, pm = metrics.bss_from_correlations(rr, rd, ee, flen, perm)
    for b in range(B):
        o = be.bss_eval_sources(ref[b], est[b], perm, flen)
        assert list(pm[b].numpy()) == list(o[3])
        for got, want in zip((sdr, sir, sar), o[:3]):
            assert np.abs(got[b].numpy() - want).max() < 1e-9


def test_bench_reference_arm_prints_one_json_line():
    """bench.py --impl reference (the CPU oracle arm the driver runs beside ours): stdout is exactly ONE JSON line carrying
    the contract's keys, whatever else the run prints (library banners go to stderr)."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '1',
                          '--ref-utts', '1'], capture_output=True, text=True, timeout=300, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, out.stdout[:500]
    line = json.loads(lines[0])
    assert line['impl'] == 'reference' and line['metric'] == 'separated_audio_seconds_per_second'
    for k in ('value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'config', 'cpu_baseline', 'e2e'):
        assert k in line, k
    assert line['e2e']['h2d_bytes_per_step'] == 0 and line['cpu_baseline']['kind'] == 'port'
