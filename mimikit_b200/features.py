"""Feature functionals on the hot path — same names, fields and call contract as the reference's
`mimikit/features/functionals.py` (`Functional.__call__` dispatching on the input type, `.inv`, `.unit`,
`.elem_type`, functionals.py:81-111), computed by the sm_100a kernels behind the C ABI.

Inputs may be CUDA tensors (zero-copy), CPU tensors or numpy arrays (host buffers: copied to the device, computed
there and copied back — the output lives where the input lived).  There is no CPU implementation.
"""
import ctypes
import dataclasses as dtc
from typing import Optional, Union

import numpy as np
import torch

from . import _capi

__all__ = ["Continuous", "Discrete", "Sample", "Frame", "Functional", "Identity", "Compose", "Normalize", "RemoveDC",
           "MuLawCompress", "MuLawExpand", "STFT", "MagSpec", "MelSpec", "Resample", "mel_filterbank", "stft_n_frames"]

N_FFT = 2048
HOP_LENGTH = 512
SR = 22050
Q_LEVELS = 256


@dtc.dataclass
class Continuous:  # functionals.py:64-68
    min_value: Union[float, int]
    max_value: Union[float, int]
    size: int


@dtc.dataclass
class Discrete:  # functionals.py:71-73
    size: int


@dtc.dataclass(frozen=True)
class Sample:  # item_spec.py:24-29
    sr: Optional[int]


@dtc.dataclass(frozen=True)
class Frame:  # item_spec.py:32-38
    frame_size: int
    hop_length: int
    padding: Optional[bool] = None


def _to_device(x, dtype=None):
    """-> (cuda tensor, restore) where restore maps a cuda result back to the caller's container."""
    _capi.require_cuda()
    if isinstance(x, np.ndarray):
        t = torch.from_numpy(np.ascontiguousarray(x))
        return t.cuda(non_blocking=False), (lambda r: r.cpu().numpy())
    if not isinstance(x, torch.Tensor):
        raise TypeError(f"expected np.ndarray or torch.Tensor, got {type(x)}")
    if x.is_cuda:
        return x, (lambda r: r)
    return x.cuda(), (lambda r: r.cpu())


class Functional:
    """functionals.py:76-111."""

    @property
    def unit(self):
        return None

    @property
    def elem_type(self):
        return None

    def torch_func(self, inputs: torch.Tensor):
        raise NotImplementedError

    def np_func(self, inputs: np.ndarray):
        return self.torch_func(inputs)

    def __call__(self, inputs):
        if not isinstance(inputs, (np.ndarray, torch.Tensor)):
            raise KeyError(type(inputs))  # the reference's dict dispatch raises KeyError
        return self.torch_func(inputs)


@dtc.dataclass
class Identity(Functional):
    """functionals.py:158-173."""

    def torch_func(self, inputs):
        return inputs

    def __call__(self, inputs):
        return inputs

    @property
    def inv(self):
        return Identity()


def _rows(x):
    """(2-D contiguous view of x's rows along the last dim, n_rows, row_len)."""
    L = x.shape[-1] if x.dim() > 0 else 1
    x2 = x.reshape(-1, L).contiguous()
    return x2, x2.shape[0], L


@dtc.dataclass
class Normalize(Functional):
    """functionals.py:236-253 — torch_func is F.normalize(inputs, p, dim): x / max(||x||_p, 1e-12).  The B200 path
    implements the reference's defaults (p = inf, dim = -1: peak normalisation of every clip): a row-maximum pass and
    one IEEE division per sample, bit-exact with torch on CPU.  `norms` of the last call are kept for inspection."""
    p: float = float('inf')
    dim: int = -1

    @property
    def elem_type(self):
        return Continuous(-1., 1., 1)

    def _check(self, x):
        if self.p != float('inf') or self.dim not in (-1, x.dim() - 1):
            raise NotImplementedError("the B200 path implements Normalize(p=inf, dim=-1) (the reference defaults) only")
        if x.dtype != torch.float32:
            raise TypeError("Normalize: the B200 path computes in fp32 (the reference's dtype for audio)")

    def torch_func(self, inputs):
        x, restore = _to_device(inputs)
        self._check(x)
        x2, n_rows, L = _rows(x)
        out = torch.empty_like(x2)
        norms = torch.empty((n_rows,), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            for r0 in range(0, n_rows, 65535):
                r1 = min(n_rows, r0 + 65535)
                _capi.check(_capi.lib().mmk_normalize_inf(x2[r0:].data_ptr(), out[r0:].data_ptr(), norms[r0:].data_ptr(),
                                                          r1 - r0, L, x2.stride(0), _capi.stream_ptr()))
        self.norms = norms.reshape(x.shape[:-1])
        return restore(out.reshape(x.shape))

    @property
    def inv(self):
        return Identity()


@dtc.dataclass
class RemoveDC(Functional):
    """functionals.py:216-233 — np_func semantics (scipy.signal.lfilter([1, -1], [1, -0.99], x) in float64, cast back),
    for numpy AND torch inputs: the reference's torch_func passes lfilter's arguments in the wrong order and cannot run
    (documented deviation).  Bit-exact with the reference's np_func: one lane per clip evaluates scipy's chain."""

    count_recomputed = False      # diagnostic: when set, `recomputed_rows` is filled after each call (synchronises)
    recomputed_rows = 0

    @property
    def elem_type(self):
        return None

    def torch_func(self, inputs):
        x, restore = _to_device(inputs)
        if x.dtype != torch.float32:
            raise TypeError("RemoveDC: the B200 path takes fp32 audio (the filter itself runs in fp64, as scipy's does)")
        x2, n_rows, L = _rows(x)
        out = torch.empty_like(x2)
        lib = _capi.lib()
        nbytes = int(lib.mmk_remove_dc_scratch_bytes(n_rows, L))      # > 0: rows are long enough to be split in time
        scratch = torch.empty((nbytes + 7) // 8, dtype=torch.float64, device=x.device) if nbytes else None
        with torch.cuda.device(x.device):
            _capi.check(lib.mmk_remove_dc(x2.data_ptr(), out.data_ptr(), n_rows, L, x2.stride(0),
                                          scratch.data_ptr() if nbytes else None, nbytes, _capi.stream_ptr()))
            if self.count_recomputed and nbytes:
                n = ctypes.c_int64(0)
                _capi.check(lib.mmk_remove_dc_recomputed_rows(scratch.data_ptr(), n_rows, L, ctypes.byref(n), _capi.stream_ptr()))
                self.recomputed_rows = n.value
        return restore(out.reshape(x.shape))

    @property
    def inv(self):
        return Identity()


class Compose(Functional):
    """functionals.py:196-213 — applies the functionals in order.  An adjacent (Normalize(), MuLawCompress) pair — the
    reference's usual waveform -> class-index pipeline — runs as ONE fused two-pass launch (row maxima; divide and
    quantise: 16 B per sample instead of 24), with the same bits as the unfused composition."""

    def __init__(self, *functionals):
        self.functionals = functionals

    @property
    def elem_type(self):
        return self.functionals[-1].elem_type if self.functionals else None

    def __call__(self, inputs):
        x, fs, i = inputs, self.functionals, 0
        while i < len(fs):
            f = fs[i]
            if (isinstance(f, Normalize) and f.p == float('inf') and f.dim == -1 and i + 1 < len(fs)
                    and isinstance(fs[i + 1], MuLawCompress) and isinstance(x, (np.ndarray, torch.Tensor))
                    and (x.dtype == torch.float32 if isinstance(x, torch.Tensor) else x.dtype == np.float32)):
                x = fs[i + 1].normalized(x, f)
                i += 2
            else:
                x = f(x)
                i += 1
        return x

    def torch_func(self, inputs):
        return self(inputs)

    @property
    def inv(self):
        return Compose(*(f.inv for f in reversed(self.functionals)))


@dtc.dataclass
class Resample(Functional):
    """functionals.py:291-310 — torch_func = torchaudio.functional.resample.  Orchestration-side (EnsembleGenerator resamples
    around every event): equal rates are the identity, anything else is handed to torchaudio on the tensor's device; it is
    library code next to the hot path, not part of it."""
    orig_sr: int = 22050
    target_sr: int = 16000

    @property
    def unit(self):
        return Sample(self.target_sr)

    def torch_func(self, inputs):
        if self.orig_sr == self.target_sr:
            return inputs
        as_np = isinstance(inputs, np.ndarray)
        x = torch.from_numpy(inputs) if as_np else inputs
        try:
            import torchaudio.functional as AF
        except ImportError as e:       # pragma: no cover
            raise NotImplementedError("Resample between different rates needs torchaudio") from e
        y = AF.resample(x, self.orig_sr, self.target_sr)
        return y.numpy() if as_np else y

    @property
    def inv(self):
        return Resample(self.target_sr, self.orig_sr)


@dtc.dataclass
class MuLawCompress(Functional):
    """functionals.py:313-342.  Bit-exact with the reference's torch_func on CPU (fp32)."""
    q_levels: int = Q_LEVELS
    compression: float = 1.

    @property
    def elem_type(self):
        return Discrete(self.q_levels)

    def torch_func(self, inputs, out_dtype=torch.int64):
        x, restore = _to_device(inputs)
        if not x.is_floating_point():
            x = x.to(torch.float)            # functionals.py:332-333
        if x.dtype != torch.float32:
            raise TypeError("MuLawCompress: the B200 path computes in fp32 (the reference's dtype for audio)")
        x = x.contiguous()
        if x.data_ptr() % 16:                # a contiguous slice such as x[1:]: the kernels load 16-byte vectors
            x = x.clone()
        lib = _capi.lib()
        if out_dtype == torch.int64:
            out = torch.empty(x.shape, dtype=torch.int64, device=x.device)
            fn = lib.mmk_mulaw_compress
        elif out_dtype == torch.uint8:
            out = torch.empty(x.shape, dtype=torch.uint8, device=x.device)
            fn = lib.mmk_mulaw_compress_u8
        else:
            raise TypeError("out_dtype must be torch.int64 (drop-in) or torch.uint8")
        with torch.cuda.device(x.device):
            _capi.check(fn(x.data_ptr(), out.data_ptr(), x.numel(), int(self.q_levels), float(self.compression),
                           _capi.stream_ptr()))
        return restore(out)

    def normalized(self, inputs, normalize: "Normalize" = None):
        """MuLawCompress(Normalize()(x)) in one fused two-pass launch (mmk_normalize_mulaw_compress)."""
        x, restore = _to_device(inputs)
        (normalize or Normalize())._check(x)
        x2, n_rows, L = _rows(x)
        out = torch.empty(x2.shape, dtype=torch.int64, device=x.device)
        norms = torch.empty((n_rows,), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            for r0 in range(0, n_rows, 65535):
                r1 = min(n_rows, r0 + 65535)
                _capi.check(_capi.lib().mmk_normalize_mulaw_compress(
                    x2[r0:].data_ptr(), out[r0:].data_ptr(), norms[r0:].data_ptr(), r1 - r0, L, x2.stride(0),
                    int(self.q_levels), float(self.compression), _capi.stream_ptr()))
        if normalize is not None:
            normalize.norms = norms.reshape(x.shape[:-1])
        return restore(out.reshape(x.shape))

    @property
    def inv(self):
        return MuLawExpand(self.q_levels, self.compression)


@dtc.dataclass
class MuLawExpand(Functional):
    """functionals.py:345-373."""
    q_levels: int = Q_LEVELS
    compression: float = 1.

    @property
    def elem_type(self):
        return Continuous(-1., 1., 1)

    def torch_func(self, inputs):
        q, restore = _to_device(inputs)
        if q.dtype != torch.int64:
            q = q.to(torch.int64)
        q = q.contiguous()
        if q.data_ptr() % 16:
            q = q.clone()
        out = torch.empty(q.shape, dtype=torch.float32, device=q.device)
        with torch.cuda.device(q.device):
            _capi.check(_capi.lib().mmk_mulaw_expand(q.data_ptr(), out.data_ptr(), q.numel(), int(self.q_levels),
                                                     float(self.compression), _capi.stream_ptr()))
        return restore(out)

    @property
    def inv(self):
        return MuLawCompress(self.q_levels, self.compression)


_ALIGN = {None: 0, "end": 1, "start": 2}


def stft_n_frames(length, n_fft=N_FFT, hop_length=HOP_LENGTH, center=True, alignment="end"):
    """(n_frames, kept_length) after STFT._fix_length (functionals.py:468-486; item_spec.convert :58-98)."""
    nf, kept = ctypes.c_int64(), ctypes.c_int64()
    _capi.check(_capi.lib().mmk_stft_n_frames(int(length), int(n_fft), int(hop_length), int(bool(center)),
                                              _ALIGN[alignment], ctypes.byref(nf), ctypes.byref(kept)))
    return nf.value, kept.value


def mel_filterbank(n_fft=N_FFT, n_mels=128, fmin=0., fmax=None, htk=False):
    """The librosa Slaney-normalised mel basis the reference reaches from functionals.py:665-668 (sr is always
    librosa's default 22050 there).  Host-side; returns (n_mels, n_fft/2+1) fp32 CPU tensor."""
    out = torch.empty(n_mels, n_fft // 2 + 1, dtype=torch.float32)
    _capi.check(_capi.lib().mmk_mel_filterbank(int(n_fft), int(n_mels), float(fmin),
                                               float(-1.0 if fmax is None else fmax), int(bool(htk)),
                                               out.data_ptr()))
    return out


def _stft_mag_mel(x, n_fft, hop, center, alignment, want_mag, fb):
    """x: cuda fp32 (..., L) -> (mag or None, mel or None)."""
    lead = x.shape[:-1]
    L = x.shape[-1]
    x2 = x.reshape(-1, L).contiguous()
    n_clips = x2.shape[0]
    n_frames, _ = stft_n_frames(L, n_fft, hop, center, alignment)
    nb = n_fft // 2 + 1
    mag = torch.empty((n_clips, n_frames, nb), dtype=torch.float32, device=x.device) if want_mag else None
    mel = torch.empty((n_clips, n_frames, fb.shape[0]), dtype=torch.float32, device=x.device) if fb is not None else None
    with torch.cuda.device(x.device):
        _capi.check(_capi.lib().mmk_stft_mag_mel(
            x2.data_ptr(), n_clips, L, x2.stride(0), int(n_fft), int(hop), int(bool(center)), _ALIGN[alignment],
            mag.data_ptr() if mag is not None else None, fb.data_ptr() if fb is not None else None,
            fb.shape[0] if fb is not None else 0, mel.data_ptr() if mel is not None else None, _capi.stream_ptr()))
    if mag is not None:
        mag = mag.reshape(*lead, n_frames, nb)
    if mel is not None:
        mel = mel.reshape(*lead, n_frames, fb.shape[0])
    return mag, mel


@dtc.dataclass
class STFT(Functional):
    """functionals.py:450-528 — only coordinate='mag' is on the hot path (MagSpec); the reference's default 'pol' and 'car'
    raise NotImplementedError, as does any window other than 'hann' (the reference's torch path always uses hann; its numpy
    path with window=None is rectangular — not built)."""
    n_fft: int = N_FFT
    hop_length: int = HOP_LENGTH
    coordinate: str = 'pol'
    center: bool = True
    window: Optional[str] = "hann"
    pad_mode: str = "constant"
    alignment: Optional[str] = "end"

    @property
    def unit(self):
        return Frame(self.n_fft, self.hop_length, padding=self.center)

    @property
    def elem_type(self):
        return Continuous(0., float("inf"), 1 + self.n_fft // 2)

    def torch_func(self, inputs):
        if self.coordinate != "mag":
            raise NotImplementedError("the B200 path implements coordinate='mag' (MagSpec) only")
        if self.window != "hann":
            raise NotImplementedError("the B200 path implements window='hann' only")
        if self.pad_mode != "constant":
            raise NotImplementedError("the B200 path implements pad_mode='constant' (the reference default) only")
        x, restore = _to_device(inputs)
        if x.dtype != torch.float32:
            x = x.to(torch.float32)
        mag, _ = _stft_mag_mel(x, self.n_fft, self.hop_length, self.center, self.alignment, True, None)
        return restore(mag)


@dtc.dataclass
class MagSpec(Functional):
    """functionals.py:576-606."""
    n_fft: int = N_FFT
    hop_length: int = HOP_LENGTH
    center: bool = True
    window: Optional[str] = "hann"
    pad_mode: str = "constant"
    alignment: Optional[str] = "end"

    @property
    def stft(self):
        return STFT(self.n_fft, self.hop_length, "mag", self.center, self.window, self.pad_mode,
                    alignment=self.alignment)

    @property
    def unit(self):
        return Frame(self.n_fft, self.hop_length, padding=self.center)

    @property
    def elem_type(self):
        return Continuous(0., float("inf"), 1 + self.n_fft // 2)

    def torch_func(self, inputs):
        return self.stft.torch_func(inputs)

    def mel(self, inputs, melspec: "MelSpec", return_mag=False):
        """Fused waveform -> (magnitudes,) mel in one kernel: the magnitudes never leave shared memory unless
        asked for.  Equivalent to melspec(self(inputs)) in the reference."""
        x, restore = _to_device(inputs)
        if x.dtype != torch.float32:
            x = x.to(torch.float32)
        fb = melspec.filterbank(self.n_fft, x.device)
        mag, mel = _stft_mag_mel(x, self.n_fft, self.hop_length, self.center, self.alignment, return_mag, fb)
        return (restore(mag), restore(mel)) if return_mag else restore(mel)


@dtc.dataclass
class MelSpec(Functional):
    """functionals.py:649-676 ("expects a MagSpec as inputs").  The reference only implements np_func (librosa); here
    numpy arrays and torch tensors both work: `MelSpec()(mag)` launches the stand-alone sparse mel kernel
    (csrc/mel_apply.cu), `MagSpec.mel(x, MelSpec())` the fused STFT -> |.| -> mel kernel."""
    n_mels: int = 128
    fmin: float = 0.
    fmax: Optional[float] = None
    htk: bool = False

    @property
    def elem_type(self):
        return Continuous(0., float("inf"), self.n_mels)

    def filterbank(self, n_fft, device=None):
        """Host filterbank (device=None) or its cached copy on `device` (no per-call H2D copy on the hot path)."""
        key = (n_fft, self.n_mels, self.fmin, self.fmax, self.htk)
        cache = MelSpec._fb_cache
        if key not in cache:
            cache[key] = mel_filterbank(n_fft, self.n_mels, self.fmin, self.fmax, self.htk)
        if device is None:
            return cache[key]
        dkey = key + (str(torch.device(device)),)
        if dkey not in cache:
            cache[dkey] = cache[key].to(device)
        return cache[dkey]

    def torch_func(self, inputs):
        """(..., frames, n_fft/2+1) magnitudes -> (..., frames, n_mels): `inputs @ mel_basis.T`, what the reference's
        np_func computes through librosa.feature.melspectrogram(S=inputs.T).T (functionals.py:665-668; its torch_func
        is an unimplemented `pass`).  n_fft is read off the bin count, as librosa does (2 * (n_bins - 1))."""
        x, restore = _to_device(inputs)
        if x.dtype != torch.float32:
            x = x.to(torch.float32)
        if x.dim() < 1 or x.shape[-1] < 2:
            raise ValueError(f"expected (..., frames, n_bins) magnitudes, got {tuple(x.shape)}")
        n_bins = x.shape[-1]
        fb = self.filterbank(2 * (n_bins - 1), x.device)
        x2 = x.reshape(-1, n_bins)
        if x2.stride(-1) != 1:
            x2 = x2.contiguous()
        out = torch.empty((x2.shape[0], self.n_mels), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            _capi.check(_capi.lib().mmk_mel_apply(x2.data_ptr(), x2.shape[0], n_bins, x2.stride(0), fb.data_ptr(),
                                                  self.n_mels, out.data_ptr(), _capi.stream_ptr()))
        return restore(out.reshape(*x.shape[:-1], self.n_mels))

    @property
    def inv(self):
        return Identity()


MelSpec._fb_cache = {}
