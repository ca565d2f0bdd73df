"""CPU-only checks of the host side: the C-ABI library loads and exports every symbol the headers
declare, host tables match the reference's, and the product fails loudly without CUDA / library."""
import ctypes
import glob
import os
import re

import numpy as np
import pytest
import torch

import meta_viterbinet_b200 as mvn
from meta_viterbinet_b200 import _lib, channel_taps
from conftest import ROOT, load_golden


def _declared_symbols():
    names = set()
    for h in glob.glob(os.path.join(ROOT, 'include', '*.h')):
        text = open(h).read()
        text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
        names |= set(re.findall(r'\b(mvn_[a-z0-9_]+)\s*\(', text))
    return sorted(names)


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    declared = _declared_symbols()
    assert len(declared) >= 15
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(raw, name), f'{name} declared in include/ but not exported'
    # and the ctypes prototypes cover all of them
    assert set(declared) <= set(_lib.exported_symbols())
    assert lib.mvn_version() >= 100


def test_argument_errors_surface_without_gpu():
    lib = _lib.load()
    rc = lib.mvn_acs_decode(None, 4, 8, 12, 8, 0, None, None, None, None)   # L out of range
    assert rc == 1
    assert b'memory_length' in lib.mvn_last_error()
    rc = lib.mvn_acs_decode(None, 4, 8, 4, 9, 0, None, None, None, None)    # n_stages > T
    assert rc == 1
    rc = lib.mvn_va_decode(None, 0, 8, 4, 8, None, 1, 0, None, None, 0, 0, None, None)
    assert rc == 1
    with pytest.raises(_lib.MVNError):
        _lib.check(rc)


def test_missing_library_fails_loudly(monkeypatch):
    monkeypatch.setattr(_lib, '_lib', None)
    monkeypatch.setattr(_lib, 'LIB_PATH', '/nonexistent/libmvn_b200.so')
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        _lib.load()


@pytest.mark.skipif(torch.cuda.is_available(), reason='checks the no-GPU behaviour')
def test_no_cpu_fallback_without_cuda():
    with pytest.raises(RuntimeError, match='CUDA'):
        mvn.ops.acs_decode(torch.zeros(2, 3, 16))
    with pytest.raises(RuntimeError, match='CUDA'):
        mvn.VNETDetector(16, {'val': 8, 'train': 8})


def test_no_product_module_imports_the_oracle():
    pkg = os.path.join(ROOT, 'meta-viterbinet_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle', text, flags=re.M), f


def test_transition_table_matches_reference():
    g = load_golden('acs')
    for L in range(3, 9):
        assert np.array_equal(mvn.create_transition_table(2 ** L), g[f'table_L{L}'])


@pytest.mark.parametrize('name,kw', [
    ('L4_fade1_ecc', dict(fading=True, fading_taps_type=1)),
    ('L4_fade2', dict(fading=True, fading_taps_type=2)),
    ('L3_static', dict(fading=False)), ('L8_static', dict(fading=False))])
def test_taps_and_state_table_bit_identical(name, kw):
    g = load_golden('va')
    L = int(g[f'{name}_meta'][0])
    h_ref = g[f'{name}_h']
    h = channel_taps.channel_taps(L, 0.2, 'time_decay', 0, indices=np.arange(h_ref.shape[0]), **kw)
    assert np.array_equal(h, h_ref)
    one = channel_taps.estimate_channel(L, 0.2, 'time_decay', 0, index=7, **kw)
    assert one.shape == (1, L) and np.array_equal(one[0], h_ref[7])
    table = channel_taps.state_priors_table(h, L)          # [n_h, S]
    assert np.array_equal(table.T.view(np.uint32), g[f'{name}_sp'].view(np.uint32))


def test_cost2100_taps_from_directory(tmp_path, monkeypatch):
    import scipy.io
    g = load_golden('va')
    h_ref = g['L4_cost2100_ecc_h']
    full = np.zeros((300, 4))
    full[:h_ref.shape[0]] = h_ref
    for i in range(4):
        scipy.io.savemat(tmp_path / f'h_{i}.mat', {'h_channel_response_mag': full[:, i].reshape(1, -1)})
    monkeypatch.setattr(channel_taps, 'COST2100_DIR', str(tmp_path))
    h = channel_taps.channel_taps(4, 0.2, 'cost2100', 0, indices=np.arange(h_ref.shape[0]))
    assert np.array_equal(h, h_ref)
    with pytest.raises(ValueError):
        channel_taps.channel_taps(4, 0.2, 'nope')
