"""CPU test: the C-ABI library loads and exports every symbol include/mmk_b200.h declares (no compute calls)."""
import os
import re

import pytest

from conftest import ROOT


def _declared_symbols():
    with open(os.path.join(ROOT, "include", "mmk_b200.h")) as f:
        src = f.read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mmk_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from mimikit_b200 import _capi
    if not os.path.exists(_capi.LIB_PATH):
        from mimikit_b200 import build
        build.build()
    lib = _capi.lib()
    declared = _declared_symbols()
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in mmk_b200.h but not exported"
        assert name in _capi.PROTOTYPES, f"{name} has no ctypes prototype"
    assert lib.mmk_abi_version() == 1


def test_host_only_entry_points_match_oracle():
    import numpy as np
    import torch
    from mimikit_b200 import features
    from oracle import restate
    for L in [2048, 2100, 5000, 22050, 220500, 3071, 3072]:
        for center in (True, False):
            for n_fft, hop in [(2048, 512), (512, 128)]:
                nf, kept = features.stft_n_frames(L, n_fft, hop, center, "end")
                assert kept == restate.stft_target_length(L, n_fft, hop, center)
                assert nf == restate.stft_n_frames(L, n_fft, hop, center)
    fb = features.mel_filterbank(2048, 128, 0., None, False).numpy()
    np.testing.assert_allclose(fb, restate.mel_filterbank(2048, 128, 0., None, False), rtol=0, atol=1e-7)
    fb = features.mel_filterbank(1024, 40, 50., 8000., True).numpy()
    np.testing.assert_allclose(fb, restate.mel_filterbank(1024, 40, 50., 8000., True), rtol=0, atol=1e-7)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from mimikit_b200 import MuLawCompress, _capi
    with pytest.raises(_capi.MmkError):
        MuLawCompress()(torch.zeros(16))
