"""CPU oracle for the VQT/CQT front-end -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module; the product path
(``zeronotesamba_b200``) never does.

PARITY UNPINNED: the arithmetic the reference runs for this path lives in un-vendored
third-party wheels -- ``librosa == 0.8.1`` and ``resampy == 0.4.2``
(/root/reference/pyproject.toml:34, /root/reference/poetry.lock:1183-1184,2188-2189) -- which
are not installed here and not installable (no network).  The reference itself ships no test,
golden vector or fixture for ``generate_XQT``.  This file restates the published algorithm of
those two packages for the exact call the reference makes
(/root/reference/zeroNoteSamba/processing/input_rep.py:27-34 ``librosa.cqt`` and :42-49
``librosa.vqt`` with hop_length=256, fmin=C0, n_bins=96, bins_per_octave=12), followed by the
reference's own post-processing ``log(abs(.) + 1e-9)`` (input_rep.py:22,36-37,51-52).

Two variants:
  * ``vqt_ref_f32``   -- follows the reference's dtypes and accumulation order (float32 signal,
    float64 weights with a float32 running sum in the resampler, float64 rFFT rounded to
    complex64, complex64 sparse x dense contraction, 1 % frequency-domain sparsification).
  * ``vqt_truth_f64`` -- the same mathematics in float64 throughout (sparsification kept by
    default, switchable), used to bound how much of any disagreement is fp32 noise.

Restated pieces (names are the third-party functions they follow):
  librosa.core.constantq.vqt / cqt / __cqt_filter_fft / __cqt_response / __trim_stack /
  __early_downsample_count, librosa.filters.constant_q / constant_q_lengths /
  window_bandwidth("hann"), librosa.core.audio.resample, librosa.core.spectrum.stft
  (window="ones", center=True, reflect), librosa.util.sparsify_rows / normalize / pad_center /
  fix_length, resampy.core.resample + resampy.interpn._resample_loop with the "kaiser_fast"
  filter (sinc_window(num_zeros=16, precision=9, rolloff=0.85, kaiser beta=8.555504641634386)).
"""
from __future__ import annotations

import functools
from typing import List, Tuple

import numpy as np
import scipy.signal
import scipy.sparse

# ---- constants of the reference call (input_rep.py:18-24) -------------------------------------
HOP_LENGTH = 256
N_OCTAVES = 8
BINS_PER_OCTAVE = 12
N_BINS = N_OCTAVES * BINS_PER_OCTAVE
EPS = 10e-10  # sic: 1e-9 (input_rep.py:22)
# librosa.note_to_hz("C0") = 440 * 2**((12 - 69) / 12)
FMIN_C0 = 440.0 * 2.0 ** ((12 - 69) / 12.0)
HANN_BANDWIDTH = 1.50018310546875  # librosa.filters.WINDOW_BANDWIDTHS["hann"]
BW_FASTEST = 0.85  # resampy kaiser_fast roll-off == librosa.core.audio.BW_FASTEST
KAISER_FAST_BETA = 8.555504641634386
KAISER_FAST_ZEROS = 16
KAISER_FAST_PRECISION = 9


# ---- alternative readings of the three points SURVEY.md Appendix A could not pin against the absent wheels -------------
# Consulted by the functions below; the defaults are this oracle's reading.  tests/test_oracle_vqt.py perturbs each one and
# bounds how far the output moves (a sensitivity test: none of them can reach the parity tolerance).
#   tap_wings:        (left, right) tap counts of resampy's 2:1 loop -- (32, 31) reads table entries j = 0..-31 and +1..+31;
#                     (32, 32) would also read entry 32 (j = +32), (31, 31) would drop j = -31
#   c64_before_scale: constant_q filters cast to complex64 BEFORE the len / n_fft scaling (True) or after it (False)
#   sparsify_ties:    among entries whose magnitude EQUALS the threshold, sparsify_rows keeps all of them ("ge": librosa's
#                     `mags >= threshold`) or only the first ("first")
VARIANT = {"tap_wings": (32, 31), "c64_before_scale": True, "sparsify_ties": "ge"}


# ---- resampy "kaiser_fast" half window (resampy.filters.sinc_window) -------------------------
@functools.lru_cache(maxsize=None)
def kaiser_fast_half_window() -> np.ndarray:
    num_bits = 2 ** KAISER_FAST_PRECISION
    n = num_bits * KAISER_FAST_ZEROS
    sinc_win = BW_FASTEST * np.sinc(BW_FASTEST * np.linspace(0, KAISER_FAST_ZEROS, num=n + 1, endpoint=True))
    taper = scipy.signal.get_window(("kaiser", KAISER_FAST_BETA), 2 * n + 1, fftbins=False)[n:]
    return taper * sinc_win  # float64, length 8193


@functools.lru_cache(maxsize=None)
def decimator_taps() -> np.ndarray:
    """The 33 table entries a 2->1 resample touches: interp_win[j*256] * sample_ratio (0.5).

    resampy._resample_loop with sample_ratio 0.5: index_step = int(0.5 * 512) = 256, eta = 0;
    left wing uses entries 0..31 (i_max = 8193 // 256 = 32), right wing entries 1..31
    (offset 256, k_max = (8193 - 256) // 256 = 31).  Entry 32 is never read.
    """
    return (0.5 * kaiser_fast_half_window())[::256].copy()


def resample_2to1_f32(x: np.ndarray) -> np.ndarray:
    """librosa.core.audio.resample(x, 2, 1, res_type="kaiser_fast", fix=True, scale=True), float32.

    Accumulation order and rounding follow resampy's loop: output is a float32 array, every tap
    does ``y[t] += weight(f64) * x[n](f32)`` i.e. a float64 add rounded back to float32; left wing
    j = 0,-1,...,-31 first, then right wing j = +1..+31.  Then fix_length to ceil(N/2) and an
    in-place float32 division by sqrt(0.5).
    """
    x = np.asarray(x, dtype=np.float32)
    n_in = x.shape[0]
    n_out = int(n_in * 1 / 2)
    h = decimator_taps()
    xz = np.zeros(n_in + 64, dtype=np.float32)  # zero extension == truncated wings
    xz[32:32 + n_in] = x
    acc = np.zeros(n_out, dtype=np.float32)
    centre = 32 + 2 * np.arange(n_out)
    n_left, n_right = VARIANT["tap_wings"]
    for i in range(n_left):  # left wing: x[n - i]
        acc = (acc.astype(np.float64) + h[i] * xz[centre - i].astype(np.float64)).astype(np.float32)
    for k in range(n_right):  # right wing: x[n + k + 1]
        acc = (acc.astype(np.float64) + h[k + 1] * xz[centre + k + 1].astype(np.float64)).astype(np.float32)
    n_fix = int(np.ceil(n_in * 0.5))
    y = np.zeros(n_fix, dtype=np.float32)
    y[:n_out] = acc
    y /= np.float32(np.sqrt(0.5))
    return y


def resample_2to1_f64(x: np.ndarray) -> np.ndarray:
    x = np.asarray(x, dtype=np.float64)
    n_in = x.shape[0]
    n_out = n_in // 2
    h = decimator_taps()
    full = np.concatenate([h[31:0:-1], h[:32]])  # taps j = -31..31
    xz = np.zeros(n_in + 64, dtype=np.float64)
    xz[32:32 + n_in] = x
    # y[t] = sum_j full[j+31] * x[2t + j]
    idx = (32 + 2 * np.arange(n_out))[:, None] + np.arange(-31, 32)[None, :]
    acc = xz[idx] @ full
    y = np.zeros(int(np.ceil(n_in * 0.5)), dtype=np.float64)
    y[:n_out] = acc
    return y / np.sqrt(0.5)


# ---- filter basis (librosa.filters.constant_q + __cqt_filter_fft + sparsify_rows) -------------
def default_gamma(bins_per_octave: int = BINS_PER_OCTAVE) -> float:
    alpha = 2.0 ** (1.0 / bins_per_octave) - 1.0
    return 24.7 * alpha / 0.108


def constant_q_lengths(sr: float, fmin: float, n_bins: int, gamma: float) -> np.ndarray:
    alpha = 2.0 ** (1.0 / BINS_PER_OCTAVE) - 1.0
    q = 1.0 / alpha
    freq = fmin * (2.0 ** (np.arange(n_bins, dtype=float) / BINS_PER_OCTAVE))
    return q * sr / (freq + gamma / alpha)


def sparsify_rows(x: np.ndarray, quantile: float) -> np.ndarray:
    """Dense result of librosa.util.sparsify_rows (zeros where the reference drops entries)."""
    mags = np.abs(x)
    norms = np.sum(mags, axis=1, keepdims=True)
    mag_sort = np.sort(mags, axis=1)
    cumulative_mag = np.cumsum(mag_sort / norms, axis=1)
    threshold_idx = np.argmin(cumulative_mag < quantile, axis=1)
    out = np.zeros_like(x)
    for i, j in enumerate(threshold_idx):
        keep = mags[i] >= mag_sort[i, j]
        if VARIANT["sparsify_ties"] == "first":
            ties = np.flatnonzero(mags[i] == mag_sort[i, j])
            keep[ties[1:]] = False
        out[i, keep] = x[i, keep]
    return out


def octave_fft_basis(octave: int, sr: float, gamma: float, sparsity: float = 0.01,
                     f32_faithful: bool = True) -> Tuple[np.ndarray, int, np.ndarray]:
    """fft_basis (12, n_fft/2+1) for octave ``octave`` (0 = top), n_fft, lengths at octave rate.

    Follows __cqt_filter_fft(my_sr, fmin_t * 2**-i, ...) and the ``*= sqrt(2**i)`` rescale in vqt.
    """
    top_freqs = FMIN_C0 * 2.0 ** (np.arange(N_BINS, dtype=float) / BINS_PER_OCTAVE)
    fmin_t = np.min(top_freqs[-BINS_PER_OCTAVE:])
    my_sr = sr / (2.0 ** octave)
    fmin_i = fmin_t * 2.0 ** (-octave)
    lengths = constant_q_lengths(my_sr, fmin_i, BINS_PER_OCTAVE, gamma)
    freqs = fmin_i * (2.0 ** (np.arange(BINS_PER_OCTAVE, dtype=float) / BINS_PER_OCTAVE))
    filts: List[np.ndarray] = []
    for ilen, freq in zip(lengths, freqs):
        sig = np.exp(np.arange(-ilen // 2, ilen // 2, dtype=float) * 1j * 2 * np.pi * freq / my_sr)
        sig = sig * scipy.signal.get_window("hann", len(sig), fftbins=True)
        sig = sig / np.sum(np.abs(sig))  # util.normalize(norm=1)
        filts.append(sig)
    n_fft = int(2.0 ** (np.ceil(np.log2(max(lengths)))))
    cdtype = np.complex64 if f32_faithful else np.complex128
    basis = np.zeros((BINS_PER_OCTAVE, n_fft), dtype=cdtype if VARIANT["c64_before_scale"] else np.complex128)
    for k, filt in enumerate(filts):
        lpad = int((n_fft - len(filt)) // 2)  # util.pad_center
        basis[k, lpad:lpad + len(filt)] = filt
    # ``basis *= lengths[:, None] / n_fft``: in-place on a complex64 array with a float64 array
    # operand -> product formed in complex128, rounded back to complex64.
    basis = (basis.astype(np.complex128) * (lengths[:, np.newaxis] / float(n_fft))).astype(cdtype)
    # numpy's pocketfft always transforms in double precision
    fft_basis = np.fft.fft(basis.astype(np.complex128), n=n_fft, axis=1)[:, : (n_fft // 2) + 1]
    if sparsity > 0:
        fft_basis = sparsify_rows(fft_basis, sparsity)
    fft_basis = fft_basis.astype(cdtype)  # lil_matrix(dtype=complex64) storage
    # ``fft_basis[:] *= np.sqrt(2 ** i)``: scalar operand -> stays in the array's precision
    rscalar = np.float32(np.sqrt(2.0 ** octave)) if f32_faithful else np.sqrt(2.0 ** octave)
    fft_basis = fft_basis * rscalar
    return fft_basis.astype(cdtype), n_fft, lengths


def _check_no_early_downsample(sr: float, gamma: float) -> None:
    """vqt()'s resampler choice and __early_downsample_count for this configuration."""
    alpha = 2.0 ** (1.0 / BINS_PER_OCTAVE) - 1.0
    q = 1.0 / alpha
    fmax_t = FMIN_C0 * 2.0 ** ((N_BINS - 1) / BINS_PER_OCTAVE)
    filter_cutoff = fmax_t * (1 + 0.5 * HANN_BANDWIDTH / q) + 0.5 * gamma
    nyquist = sr / 2.0
    if not filter_cutoff < BW_FASTEST * nyquist:
        raise NotImplementedError("oracle covers the kaiser_fast branch only (sr=16000 path)")
    count1 = max(0, int(np.ceil(np.log2(BW_FASTEST * nyquist / filter_cutoff)) - 1) - 1)
    num_twos = 8  # hop 256
    count2 = max(0, num_twos - N_OCTAVES + 1)
    if min(count1, count2) != 0:
        raise NotImplementedError("oracle covers the no-early-downsample configuration only")


def _stft_ones(y: np.ndarray, n_fft: int, hop: int, cdtype) -> np.ndarray:
    """librosa.stft(y, n_fft, hop, window="ones", center=True, pad_mode="reflect")."""
    ypad = np.pad(y, int(n_fft // 2), mode="reflect")
    n_frames = 1 + (len(ypad) - n_fft) // hop
    idx = (np.arange(n_frames) * hop)[None, :] + np.arange(n_fft)[:, None]
    frames = ypad[idx].astype(np.float64)
    return np.fft.rfft(frames, axis=0).astype(cdtype)


def xqt_complex(y: np.ndarray, sr: int = 16000, mode: str = "vqt", f32_faithful: bool = True,
                sparsity: float = 0.01) -> np.ndarray:
    """Complex V (96, 1 + N // 256) as librosa.vqt / librosa.cqt would return it (scale=True)."""
    if mode == "vqt":
        gamma = default_gamma()
    elif mode == "cqt":
        gamma = 0.0
    else:
        raise Exception("Mode can only be vqt or cqt!")  # input_rep.py:56-57
    _check_no_early_downsample(float(sr), gamma)
    rdtype = np.float32 if f32_faithful else np.float64
    cdtype = np.complex64 if f32_faithful else np.complex128
    my_y = np.asarray(y, dtype=rdtype)
    my_hop = HOP_LENGTH
    resp = []
    for i in range(N_OCTAVES):
        if i > 0:
            my_y = resample_2to1_f32(my_y) if f32_faithful else resample_2to1_f64(my_y)
            my_hop //= 2
        fft_basis, n_fft, _ = octave_fft_basis(i, float(sr), gamma, sparsity, f32_faithful)
        d = _stft_ones(my_y, n_fft, my_hop, cdtype)
        if f32_faithful:
            # scipy CSR (complex64) x dense (complex64): accumulation in complex64, ascending bin order
            c = scipy.sparse.csr_matrix(fft_basis).dot(np.asfortranarray(d))
        else:
            c = fft_basis @ d
        resp.append(np.asarray(c))
    max_col = min(c.shape[-1] for c in resp)  # __trim_stack
    out = np.empty((N_BINS, max_col), dtype=cdtype, order="F")
    end = N_BINS
    for c in resp:
        out[end - BINS_PER_OCTAVE:end] = c[:, :max_col]
        end -= BINS_PER_OCTAVE
    lengths = constant_q_lengths(float(sr), FMIN_C0, N_BINS, gamma)
    out /= np.sqrt(lengths[:, np.newaxis]).astype(rdtype)
    return out


def vqt_ref_f32(y: np.ndarray, sr: int = 16000, mode: str = "vqt") -> np.ndarray:
    """Oracle for generate_XQT(y, sr, mode) (input_rep.py:11-57): float32 (96, 1 + N // 256)."""
    v = np.abs(xqt_complex(y, sr, mode, f32_faithful=True))
    return np.log(v + np.float32(EPS)).astype(np.float32) if v.dtype == np.float32 else np.log(v + EPS).astype(np.float32)


def vqt_truth_f64(y: np.ndarray, sr: int = 16000, mode: str = "vqt", sparsity: float = 0.01) -> np.ndarray:
    """Same mathematics in float64 (returns float64 log-magnitudes)."""
    v = np.abs(xqt_complex(y, sr, mode, f32_faithful=False, sparsity=sparsity))
    return np.log(v + EPS)


def vqt_magnitude_f64(y: np.ndarray, sr: int = 16000, mode: str = "vqt") -> np.ndarray:
    return np.abs(xqt_complex(y, sr, mode, f32_faithful=False))


# ---- time-domain form of the same basis (what a device plan needs; used to cross-check) -------
def octave_time_kernels(octave: int, sr: float, gamma: float) -> Tuple[np.ndarray, int]:
    """g_i[k, n] = sum_b basis_i[k, b] e^{-2 pi i b n / n_fft}: C_i[k,t] = sum_n g_i[k,n] ypad_i[t hop_i + n].

    Float64 from the float32-faithful (sparsified, complex64) frequency-domain basis.
    """
    fft_basis, n_fft, _ = octave_fft_basis(octave, sr, gamma, 0.01, True)
    b = np.arange(n_fft // 2 + 1)[:, None]
    n = np.arange(n_fft)[None, :]
    e = np.exp(-2j * np.pi * b * n / n_fft)
    return fft_basis.astype(np.complex128) @ e, n_fft


# ---- RMS stem gate (row f3): librosa.feature.rms + check_CL_clips --------------------------------
def rms_frames(y: np.ndarray, frame_length: int = 2048, hop_length: int = 512) -> np.ndarray:
    """librosa.feature.rms(y=y, frame_length=2048, hop_length=512) (center=True, reflect), shape (F,)."""
    y = np.asarray(y, dtype=np.float32)
    ypad = np.pad(y, int(frame_length // 2), mode="reflect")
    n_frames = 1 + (len(ypad) - frame_length) // hop_length
    idx = (np.arange(n_frames) * hop_length)[None, :] + np.arange(frame_length)[:, None]
    return np.sqrt(np.mean(np.abs(ypad[idx]) ** 2, axis=0))


def rms_fraction(anchor: np.ndarray, positive: np.ndarray) -> float:
    """Fraction used by check_CL_clips (/root/reference/zeroNoteSamba/processing/stem_check.py:33-45)."""
    stem, ros = rms_frames(anchor), rms_frames(positive)
    ok = (stem > ros / 2).astype(int) * (stem < ros * 4).astype(int)
    return float(np.sum(ok) / len(ok))
