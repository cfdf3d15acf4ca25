"""Shared helpers for the parity tests (oracle = checker, dl4ss_b200 = thing under test)."""
import copy

import numpy as np
import torch


def rel_err(a, b):
    """max|a-b| / max|b| -- the 'relative to the peak' error used for spectra and waveforms."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def set_config(cfg_mod, **kw):
    old = {k: getattr(cfg_mod, k) for k in kw}
    for k, v in kw.items():
        setattr(cfg_mod, k, v)
    return old


def build_pair(cell, layers, F, T, complex_mask, self_tune=True, seed=1, num_spk=101, E=50, H=300, mode='dot'):
    """Oracle modules (torch CPU) and dl4ss_b200 modules (CUDA) sharing bit-identical seeded weights."""
    import dl4ss_b200 as d
    from oracle import modules_ref as mr
    torch.manual_seed(seed)
    rc = mr.RefConfig(HIDDEN_UNITS=H, EMBEDDING_SIZE=E, NUM_LAYERS=layers, is_ComlexMask=complex_mask,
                      is_SelfTune=self_tune)
    ref = {
        'cfg': rc,
        'mix': mr.MIX_SPEECH(rc, F, T, cell, layers),
        'emb': mr.SPEECH_EMBEDDING(rc, num_spk, E, 2),
        'att': mr.ATTENTION(rc, E, mode),
        'adj': mr.ADDJUST(rc, 2 * H, E) if self_tune else None,
    }
    d.config.HIDDEN_UNITS, d.config.EMBEDDING_SIZE, d.config.NUM_LAYERS = H, E, layers
    d.config.is_ComlexMask, d.config.is_SelfTune = complex_mask, self_tune
    dev = torch.device('cuda:0')
    ours = {
        'mix': d.MIX_SPEECH(F, T, cell=cell, num_layers=layers).to(dev),
        'emb': d.SPEECH_EMBEDDING(num_spk, E, 2).to(dev),
        'att': d.ATTENTION(E, mode).to(dev),
        'adj': d.ADDJUST(2 * H, E).to(dev) if self_tune else None,
    }
    for k in ('mix', 'emb', 'att', 'adj'):
        if ref[k] is not None:
            ours[k].load_state_dict(copy.deepcopy(ref[k].state_dict()))
    return ref, ours
