"""CPU-only checks of the C-ABI library: it loads, exports every symbol include/zns.h declares,
its host-side VQT basis equals the oracle's, and device entry points fail loudly without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from zeronotesamba_b200 import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FMIN = 440.0 * 2.0 ** ((12 - 69) / 12.0)


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "zns.h")).read()
    declared = set(re.findall(r"^(?:int|const char\*)\s+(zns_[a-z0-9_]+)\s*\(", header, flags=re.M))
    lib = L.lib()
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/zns.h but not exported"
    assert set(L.EXPORTS) == declared, set(L.EXPORTS) ^ declared
    assert lib.zns_version() == 100


def test_decimator_taps_equal_oracle():
    from oracle import vqt_oracle as vo
    t = np.zeros(32)
    L.check(L.lib().zns_vqt_decimator_taps_host(t.ctypes.data))
    assert np.allclose(t, vo.decimator_taps()[:32], rtol=1e-12, atol=1e-18)


@pytest.mark.parametrize("gamma,mode", [(-1.0, "vqt"), (0.0, "cqt")])
def test_basis_equals_oracle(gamma, mode):
    from oracle import vqt_oracle as vo
    g = vo.default_gamma() if mode == "vqt" else 0.0
    for octave in range(8):
        re_ = np.zeros(12 * 1024, np.float32)
        im_ = np.zeros(12 * 1024, np.float32)
        nf = C.c_int(0)
        L.check(L.lib().zns_vqt_basis_host(16000, 96, 12, FMIN, gamma, octave, re_.ctypes.data, im_.ctypes.data, C.byref(nf)))
        ker, n_fft = vo.octave_time_kernels(octave, 16000.0, g)
        assert nf.value == n_fft
        got = re_[:12 * n_fft].reshape(12, n_fft) + 1j * im_[:12 * n_fft].reshape(12, n_fft)
        scale = np.abs(ker).max()
        assert np.abs(got - ker).max() < 2e-7 * scale, (octave, np.abs(got - ker).max() / scale)


def test_frames_and_argument_errors():
    lib = L.lib()
    assert lib.zns_vqt_num_frames(160000, 256) == 626
    assert lib.zns_vqt_num_frames(480000, 256) == 1876
    assert lib.zns_vqt_num_frames(80001, 256) == 313
    h = C.c_void_p()
    rc = lib.zns_vqt_plan_create(16000, 100, 96, 12, FMIN, -1.0, 1, 16000, C.byref(h))
    assert rc == 1 and b"hop_length" in lib.zns_last_error()
    rc = lib.zns_vqt_plan_create(8000, 256, 96, 12, FMIN, -1.0, 1, 16000, C.byref(h))
    assert rc != 0
    with pytest.raises(L.ZnsError):
        L.check(lib.zns_adam_flat(None, None, None, None, 4, 1e-6, 0.9, 0.999, 1e-8, 1, None, 1.0, None))


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = C.c_void_p()
    rc = L.lib().zns_vqt_plan_create(16000, 256, 96, 12, FMIN, -1.0, 1, 16000, C.byref(h))
    assert rc == 2, "plan creation must fail with ZNS_ERR_CUDA when no device is present"
