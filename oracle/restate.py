"""TEST INFRASTRUCTURE ONLY (oracle) — CPU restatement, in numpy fp32, of the reference's hot path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
module.  The product package (mimikit_b200/) never does; it fails loudly without its CUDA library.

Every function cites the reference lines it follows (paths relative to /root/reference).  Pinning status:
  * mu-law            : C restatement (oracle/c/oracle_feat.c) — bit-exact vs the live reference, pinned by
                        tests/golden/mulaw_*.npz and the exhaustive log1pf sweep (oracle/validate_mulaw.py).
  * WaveNet / SampleRNN: pinned by tests/golden/{wavenet,samplernn}_*.npz generated from the live reference
                        (oracle/make_golden.py); structural KATs (rf, output lengths) follow
                        tests/test_wavenet.py:251-275 of the reference.
  * STFT magnitudes   : pinned by tests/golden/magspec_*.npz (live reference, torch.stft) and the frame-count
                        KATs of the reference's tests/test_fft_alignment.py:28-110.
  * mel               : PARITY UNPINNED at the librosa boundary — librosa (pyproject.toml:38 `librosa>=0.9.1`,
                        no upper pin, not vendored, not installed here) holds the arithmetic; restated from its
                        published Slaney-filterbank algorithm and cross-checked against
                        torchaudio.functional.melscale_fbanks(norm="slaney", mel_scale="slaney").
"""
import ctypes
import math
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
f32 = np.float32


def clib():
    """The C part of the oracle (oracle/c/oracle_feat.c), built by `make -C oracle` / __graft_entry__.build()."""
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "_build", "liboracle.so")
        if not os.path.exists(path):
            raise RuntimeError(f"{path} missing: run `make -C oracle` (or __graft_entry__.build())")
        _LIB = ctypes.CDLL(path)
        _LIB.orc_log1pf.restype = ctypes.c_float
        _LIB.orc_log1pf.argtypes = [ctypes.c_float]
        _LIB.orc_expf.restype = ctypes.c_float
        _LIB.orc_expf.argtypes = [ctypes.c_float]
    return _LIB


def _ptr(a):
    return ctypes.c_void_p(a.ctypes.data)


# ---------------------------------------------------------------------------------------------
# features
# ---------------------------------------------------------------------------------------------

def mulaw_compress(x, q_levels=256, compression=1.0):
    """mimikit/features/functionals.py:330-338 (MuLawCompress.torch_func, fp32 CPU)."""
    x = np.ascontiguousarray(x, dtype=f32)
    out = np.empty(x.shape, dtype=np.int64)
    clib().orc_mulaw_compress(_ptr(x), _ptr(out), ctypes.c_int64(x.size), ctypes.c_int(int(q_levels)),
                              ctypes.c_float(float(compression)))
    return out


def mulaw_expand(idx, q_levels=256, compression=1.0):
    """mimikit/features/functionals.py:361-369 (MuLawExpand.torch_func); exp is the u10 Sleef algorithm, which is
    within 1 ulp of (not bit-identical to) torch's AVX-512 exp — float output, tolerance 1e-6 abs."""
    idx = np.ascontiguousarray(idx, dtype=np.int64)
    out = np.empty(idx.shape, dtype=f32)
    clib().orc_mulaw_expand(_ptr(idx), _ptr(out), ctypes.c_int64(idx.size), ctypes.c_int(int(q_levels)),
                            ctypes.c_float(float(compression)))
    return out


def normalize_inf(x, eps=1e-12):
    """Normalize(p=inf, dim=-1).torch_func = F.normalize(x, p=inf, dim=-1) (features/functionals.py:236-253; torch:
    x / clamp_min(vector_norm(x, inf, dim, keepdim), eps)) in fp32: one IEEE division per sample."""
    x = np.asarray(x, dtype=f32)
    with np.errstate(invalid="ignore", divide="ignore"):
        n = np.abs(x).max(axis=-1, keepdims=True)
        d = np.where(np.isnan(n), n, np.maximum(n, f32(eps))).astype(f32)
        return (x / d).astype(f32)


def remove_dc(x):
    """RemoveDC.np_func (features/functionals.py:216-233): scipy.signal.lfilter([1, -1], [1, -0.99], x, axis=-1) —
    scipy's fp64 direct-form-II-transposed loop restated in C — cast back to fp32."""
    x = np.ascontiguousarray(x, dtype=f32)
    out = np.empty_like(x)
    L = x.shape[-1] if x.ndim else 1
    clib().orc_remove_dc(_ptr(x), _ptr(out), ctypes.c_int64(x.size // max(L, 1)), ctypes.c_int64(L))
    return out


def expf_portable(x):
    x = np.ascontiguousarray(x, dtype=f32)
    out = np.empty_like(x)
    clib().orc_expf_arr(_ptr(x), _ptr(out), ctypes.c_int64(x.size))
    return out


def stft_target_length(L, n_fft, hop, center):
    """STFT._fix_length (functionals.py:468-486) through item_spec.convert (item_spec.py:58-98)."""
    extra = 0 if center else (n_fft - hop)
    n = (L - extra) // hop + int(center)  # Sample -> Frame (as_length) + int(center)
    n -= int(center)                      # Frame -> Sample: x -= int(has_padding)
    return int(n * hop) + extra


def stft_n_frames(L, n_fft, hop, center):
    t = stft_target_length(L, n_fft, hop, center)
    return t // hop + 1 if center else (t - n_fft) // hop + 1


def stft_fix_length(x, n_fft, hop, center, alignment="end"):
    if alignment is None:
        return x
    t = stft_target_length(x.shape[-1], n_fft, hop, center)
    if alignment == "end":
        return x[..., -t:] if t != 0 else x  # python's x[-0:] keeps everything (reference quirk)
    if alignment == "start":
        return x[..., :t]
    return x


def hann_periodic(n_fft):
    """torch.hann_window(n_fft) (periodic) — functionals.py:513 always uses it on the torch path."""
    n = np.arange(n_fft, dtype=np.float64)
    return 0.5 * (1.0 - np.cos(2.0 * np.pi * n / n_fft))


def magspec(x, n_fft=2048, hop=512, center=True, alignment="end"):
    """MagSpec.torch_func -> STFT(coordinate='mag').torch_func (functionals.py:507-524, 576-606):
    length fix, zero ("constant") centre padding, periodic hann, rFFT, |.|, layout (..., frames, n_fft/2+1).
    Evaluated in fp64 and rounded to fp32 (the reference evaluates in fp32; tolerance 1e-4, see tests)."""
    x = np.asarray(x, dtype=np.float64)
    x = stft_fix_length(x, n_fft, hop, center, alignment)
    if center:
        pad = [(0, 0)] * (x.ndim - 1) + [(n_fft // 2, n_fft // 2)]
        x = np.pad(x, pad)
    n_frames = (x.shape[-1] - n_fft) // hop + 1
    idx = np.arange(n_fft)[None, :] + hop * np.arange(n_frames)[:, None]
    frames = x[..., idx] * hann_periodic(n_fft)
    return np.abs(np.fft.rfft(frames, axis=-1)).astype(f32)


def _hz_to_mel(f, htk=False):
    f = np.asarray(f, dtype=np.float64)
    if htk:
        return 2595.0 * np.log10(1.0 + f / 700.0)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-30) / min_log_hz) / logstep, mels)


def _mel_to_hz(m, htk=False):
    m = np.asarray(m, dtype=np.float64)
    if htk:
        return 700.0 * (10.0 ** (m / 2595.0) - 1.0)
    f_sp = 200.0 / 3
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)


def mel_filterbank(n_fft=2048, n_mels=128, fmin=0.0, fmax=None, htk=False, sr=22050):
    """librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax, htk, norm='slaney') as reached from
    MelSpec.np_func (functionals.py:665-668) -> librosa.feature.melspectrogram(S=...).  The reference never passes
    `sr`, so librosa's default 22050 applies whatever the audio's rate is.  Returns (n_mels, n_fft/2+1) fp32."""
    if fmax is None:
        fmax = sr / 2.0
    n_bins = 1 + n_fft // 2
    fftfreqs = np.linspace(0.0, sr / 2.0, n_bins)
    mel_pts = _mel_to_hz(np.linspace(_hz_to_mel(fmin, htk), _hz_to_mel(fmax, htk), n_mels + 2), htk)
    fdiff = np.diff(mel_pts)
    ramps = mel_pts[:, None] - fftfreqs[None, :]
    w = np.zeros((n_mels, n_bins), dtype=np.float64)
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        w[i] = np.maximum(0.0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_pts[2:n_mels + 2] - mel_pts[:n_mels])
    w *= enorm[:, None]
    return w.astype(f32)


def melspec(mag, n_mels=128, fmin=0.0, fmax=None, htk=False, n_fft=None):
    """MelSpec.np_func (functionals.py:665-668): mel_basis @ S with S = mag.T, transposed back.
    `mag` is (..., frames, n_bins) fp32; accumulation in fp64, rounded to fp32."""
    mag = np.asarray(mag)
    n_fft = 2 * (mag.shape[-1] - 1) if n_fft is None else n_fft
    fb = mel_filterbank(n_fft, n_mels, fmin, fmax, htk)
    return (mag.astype(np.float64) @ fb.T.astype(np.float64)).astype(f32)


# ---------------------------------------------------------------------------------------------
# sampling contract (ours; SURVEY.md App. A.3, blocked-scan order — see DESIGN.md §sampling)
# ---------------------------------------------------------------------------------------------

def sample_inverse_cdf(logits, temperature, u):
    """Replaces torch.multinomial in CategoricalSampler.forward (mimikit/modules/targets.py:40-52), which cannot
    be driven by external noise.

      l_k = logits_k / T_b (fp32);  m = max_k l_k;  e_k = expf_u10(l_k - m)
      the Q classes are dealt to 32 lanes in runs of n = ceil(Q/32) consecutive classes; each lane forms its
      sequential fp32 inclusive prefix; lane totals are scanned with the 5-stage Kogge-Stone pattern
      (offsets 1,2,4,8,16); c_k = exclusive_lane_prefix + in-lane prefix; total = inclusive prefix of lane 31
      q = min(Q-1, #{k : c_k <= u * total})

    logits (B,Q) fp32, temperature (1,) or (B,) fp32, u (B,) fp32 in [0,1).  Returns (B,) int64."""
    logits = np.asarray(logits, dtype=f32)
    B, Q = logits.shape
    T = np.broadcast_to(np.asarray(temperature, dtype=f32).reshape(-1), (B,)) if np.size(temperature) != B \
        else np.asarray(temperature, dtype=f32).reshape(B)
    u = np.asarray(u, dtype=f32).reshape(B)
    l = (logits / T[:, None]).astype(f32)
    m = l.max(axis=1, keepdims=True)
    e = expf_portable((l - m).astype(f32))
    n = (Q + 31) // 32
    pad = np.zeros((B, 32 * n), dtype=f32)
    pad[:, :Q] = e
    lanes = pad.reshape(B, 32, n)
    pref = np.empty_like(lanes)
    acc = np.zeros((B, 32), dtype=f32)
    for i in range(n):
        acc = (acc + lanes[:, :, i]).astype(f32)
        pref[:, :, i] = acc
    v = acc.copy()
    for off in (1, 2, 4, 8, 16):
        nv = v.copy()
        nv[:, off:] = (v[:, off:] + v[:, :-off]).astype(f32)
        v = nv
    excl = np.zeros_like(v)
    excl[:, 1:] = v[:, :-1]
    c = (excl[:, :, None] + pref).astype(f32).reshape(B, 32 * n)[:, :Q]
    thr = (u * v[:, 31]).astype(f32)
    cnt = (c <= thr[:, None]).sum(axis=1)
    return np.minimum(Q - 1, cnt).astype(np.int64)


def argmax_first(logits):
    """targets.py:43 — torch argmax: lowest index among equal maxima (numpy has the same rule)."""
    return np.argmax(logits, axis=-1).astype(np.int64)


def normalize_temperature(temperature, B):
    """targets.py:27-34 (as_tensor): None | float | 1-sequence | (B,) -> None or fp32 (1,)/(B,)."""
    if temperature is None:
        return None
    t = np.asarray(temperature, dtype=f32).reshape(-1)
    if t.size not in (1, B):
        raise ValueError(f"temperature must have 1 or {B} entries, got {t.size}")
    return t


# ---------------------------------------------------------------------------------------------
# shared pieces of the networks
# ---------------------------------------------------------------------------------------------

def _sigmoid(x):
    return (1.0 / (1.0 + np.exp(-x.astype(f32)))).astype(f32)


def _mish(x):
    x = x.astype(f32)
    sp = np.where(x > 20.0, x, np.log1p(np.exp(np.minimum(x, 20.0)))).astype(f32)  # F.softplus threshold 20
    return (x * np.tanh(sp)).astype(f32)


def _softplus(x):
    x = x.astype(f32)
    return np.where(x > 20.0, x, np.log1p(np.exp(np.minimum(x, 20.0)))).astype(f32)  # nn.Softplus(beta=1, threshold=20)


# the point-wise members of ActivationEnum (modules/activations.py:26-40, 70-88; torch.nn for the rest)
ACTIVATIONS = {
    "Tanh": lambda x: np.tanh(x.astype(f32)).astype(f32),
    "Sigmoid": _sigmoid,
    "Mish": _mish,
    "ReLU": lambda x: np.maximum(x.astype(f32), f32(0)),
    "Softplus": _softplus,
    "Identity": lambda x: x.astype(f32),
    "Abs": lambda x: np.abs(x.astype(f32)),
    "Sin": lambda x: np.sin(x.astype(f32)).astype(f32),
    "Cos": lambda x: np.cos(x.astype(f32)).astype(f32),
}


def mlp_head(x, W1, b1, W2, b2, min_temp, Q, Wh=None, bh=None, n_hidden=0):
    """networks/mlp.py:44-63: Linear, Mish, n_hidden x (the SAME Linear(Hh, Hh), Mish — the tuple repetition of :47-50 shares
    one module), Linear(+1), learned-temperature divide."""
    hid = _mish(x @ W1.T + b1)
    for _ in range(n_hidden):
        hid = _mish((hid @ Wh.T + bh).astype(f32))
    z = hid @ W2.T + b2
    if min_temp is None:                       # MLP(min_temperature=None): Q outputs, no temperature column (mlp.py:29, 54-62)
        return z.astype(f32)
    temp = np.maximum(_sigmoid(z[..., Q:Q + 1]), f32(min_temp))
    return (z[..., :Q] / temp).astype(f32)


def _np(sd, key):
    v = sd[key]
    return np.ascontiguousarray(v.detach().cpu().numpy() if hasattr(v, "detach") else v, dtype=f32)


def _np_or0(sd, key):
    """A bias of a module built with bias=False is absent from the state dict; adding 0.0f changes no value."""
    return _np(sd, key) if key in sd else f32(0)


def fold_weight_norm(sd):
    """nn.utils.weight_norm (sample_rnn_v2.py:67-81): w = g * v / ||v|| with the norm over all dims but 0."""
    out = {}
    for k, v in sd.items():
        if k.endswith("_g") or k.endswith("_v"):
            continue
        out[k] = v
    for k in sd:
        if k.endswith("_v"):
            base = k[:-2]
            v, g = _np(sd, k).astype(np.float64), _np(sd, base + "_g").astype(np.float64)
            norm = np.sqrt((v.reshape(v.shape[0], -1) ** 2).sum(1)).reshape((-1,) + (1,) * (v.ndim - 1))
            out[base] = (g * v / norm).astype(f32)
    return out


# ---------------------------------------------------------------------------------------------
# WaveNet
# ---------------------------------------------------------------------------------------------

def wavenet_kernels_and_dilations(kernel_sizes, blocks):
    """WaveNet.get_kernels_and_dilation (networks/wavenet_v2.py:295-327)."""
    kernel_sizes, blocks = tuple(kernel_sizes), tuple(blocks)
    if not blocks:
        dil, acc = [], 1
        for k in (1,) + kernel_sizes:
            acc *= k
            dil.append(acc)
        return list(kernel_sizes), dil  # NB: the reference yields one more dilation than kernels; zip() trims it
    if len(set(blocks)) == 1 and blocks[0] == len(kernel_sizes):
        dil = []
        for _ in blocks:
            acc = 1
            dil.append(acc)
            for k in kernel_sizes[:-1]:
                acc *= k
                dil.append(acc)
        return list(kernel_sizes) * len(blocks), dil
    if len(kernel_sizes) == sum(blocks):
        dil, start = [], 0
        for b in blocks:
            acc = 1
            dil.append(acc)
            for k in kernel_sizes[start:start + b - 1]:
                acc *= k
                dil.append(acc)
            start += b
        return list(kernel_sizes), dil
    if len(kernel_sizes) == 1:
        k = kernel_sizes[0]
        return [k] * sum(blocks), [k ** i for b in blocks for i in range(b)]
    raise ValueError(f"number of layers and number of kernel sizes not compatible."
                     f" Got kernel_sizes={kernel_sizes} ; blocks={blocks}")


def wavenet_rf(kernel_sizes, blocks):
    """WaveNet.rf (wavenet_v2.py:337-339): sum of (k-1)*d over layers + 1."""
    ks, ds = wavenet_kernels_and_dilations(kernel_sizes, blocks)
    return sum((k - 1) * d for k, d in zip(ks, ds)) + 1


class WaveNetOracle:
    """Cached-step restatement of WaveNet.generate_step == forward on the last rf samples
    (networks/wavenet_v2.py:447-452, 276-293; WNLayer.forward 131-176), for the mu-law embedding-input, gated configuration:
    any kernel sizes (tap j of a size-k layer reads the sample (k - 1 - j) dilations back), `layerwise_inputs` (:283-284:
    the embedded network input is added to every layer's output) and hidden MLP layers (mlp.py:47-50: one shared Linear).
    `pad_side` does not enter: the generation loop evaluates the LAST position of an rf-long window (eval_slice, :273), whose
    dependency cone never touches the left padding, so pad_side=1 and pad_side=0 give the same value there.
    `with_affine_residuals` (:121-122, 148-149, 164-165; parametrized.py:34-47): a layer's input first goes through
    aff_res, z = x_hat * a + b with (x_hat, a, b) the three chunks of one 1x1 conv; the dilated conv AND the residual add
    (:174 trims the re-bound `inputs_dilated`) then read z, so the per-layer history holds z."""

    def __init__(self, state_dict, blocks, kernel_sizes=(2,), q_levels=256, layerwise_inputs=False, n_mlp_hidden=0,
                 reverse_layer_order=False, act_f="Tanh", act_g="Sigmoid"):
        sd = state_dict
        self.act_f, self.act_g = ACTIVATIONS[str(act_f)], ACTIVATIONS[str(act_g or "Sigmoid")]   # wavenet_v2.py:198-199, 224-225
        ks, ds = wavenet_kernels_and_dilations(kernel_sizes, blocks)
        self.kernels = [int(k) for k, _ in zip(ks, ds)]
        self.dilations = [int(d) for _, d in zip(ks, ds)]
        if reverse_layer_order:                               # wavenet_v2.py:270: nn.ModuleList(reversed(layers))
            self.kernels.reverse()
            self.dilations.reverse()
        self.L = len(self.dilations)
        self.Q = q_levels
        self.layerwise_inputs = bool(layerwise_inputs)
        self.n_mlp_hidden = int(n_mlp_hidden)
        self.E = _np(sd, "input_modules.0.0.weight")
        self.C = self.E.shape[1]
        self.Wd, self.bd, self.Ws, self.bs, self.Wr, self.br = [], [], [], [], [], []
        self.Wa, self.ba = [], []
        self.has_skips = "layers.0.conv_skip.weight" in sd
        for l in range(self.L):
            self.gated = f"layers.{l}.conv_dil.0.0.weight" in sd     # act_g=None: a bare Conv1d, y = tanh(conv) (wavenet_v2.py:109-112, 160-163)
            pre = f"layers.{l}.conv_dil.0.0." if self.gated else f"layers.{l}.conv_dil.0."
            w = _np(sd, pre + "weight")  # (2C | C, C / groups, k): tap 0 = oldest sample (cross-correlation)
            assert w.shape[2] == self.kernels[l]
            if w.shape[1] != self.C:     # grouped conv (wavenet_v2.py:92): the block-diagonal weight written out densely
                G = self.C // w.shape[1]
                dense = np.zeros((w.shape[0], self.C, w.shape[2]), dtype=f32)
                og, cg = w.shape[0] // G, w.shape[1]
                for gi in range(G):
                    dense[gi * og:(gi + 1) * og, gi * cg:(gi + 1) * cg] = w[gi * og:(gi + 1) * og]
                w = dense
            self.Wd.append([np.ascontiguousarray(w[:, :, j]) for j in range(w.shape[2])])
            self.bd.append(_np_or0(sd, pre + "bias"))                # WaveNet.Config.bias=False (wavenet_v2.py:92-93): no layer biases
            if f"layers.{l}.aff_res.params.weight" in sd:
                self.Wa.append(_np(sd, f"layers.{l}.aff_res.params.weight")[:, :, 0])
                self.ba.append(_np_or0(sd, f"layers.{l}.aff_res.params.bias"))
            else:
                self.Wa.append(None)
                self.ba.append(None)
            if self.has_skips:
                self.Ws.append(_np(sd, f"layers.{l}.conv_skip.weight")[:, :, 0])
                self.bs.append(_np_or0(sd, f"layers.{l}.conv_skip.bias"))
            if f"layers.{l}.conv_res.weight" in sd:
                self.Wr.append(_np(sd, f"layers.{l}.conv_res.weight")[:, :, 0])
                self.br.append(_np_or0(sd, f"layers.{l}.conv_res.bias"))
            else:
                self.Wr.append(None)
                self.br.append(None)
        self.Wd0 = [w[0] for w in self.Wd]           # the two taps of a size-2 layer, as the bf16 oracle names them
        self.Wd1 = [w[-1] for w in self.Wd]
        p = "output_modules.0.estimator.0."
        last = 2 + 2 * self.n_mlp_hidden
        self.W1, self.b1 = _np(sd, p + "fc.0.weight"), _np(sd, p + "fc.0.bias")
        self.Wh = _np(sd, p + "fc.2.weight") if self.n_mlp_hidden else None
        self.bh = _np(sd, p + "fc.2.bias") if self.n_mlp_hidden else None
        self.W2, self.b2 = _np(sd, p + f"fc.{last}.weight"), _np(sd, p + f"fc.{last}.bias")
        self.min_temp = float(_np(sd, p + "min_temp").reshape(-1)[0]) if p + "min_temp" in sd else None
        self.rf = sum((k - 1) * d for k, d in zip(self.kernels, self.dilations)) + 1

    def _head(self, out):
        return mlp_head(out, self.W1, self.b1, self.W2, self.b2, self.min_temp, self.Q, self.Wh, self.bh, self.n_mlp_hidden)

    def _layer(self, l, taps, skips, e=None):
        """taps: the k inputs of the layer, oldest first (the last one is the current sample); e: the embedded network input
        aligned with the output (layerwise_inputs)."""
        C = self.C
        a = self.bd[l]
        for w, x in zip(self.Wd[l], taps):                               # wavenet_v2.py:101,150
            a = x @ w.T + a if a.ndim == 1 else a + x @ w.T
        a = a.astype(f32)
        if self.gated:
            y = (self.act_f(a[..., :C]) * self.act_g(a[..., C:])).astype(f32)   # :102,151
        else:
            y = self.act_f(a).astype(f32)                                    # :160-163 (act_g=None)
        if self.has_skips:
            s = y @ self.Ws[l].T + self.bs[l]                            # :165-171
            skips = s if skips is None else (s + skips).astype(f32)
        h = (taps[-1] + (y @ self.Wr[l].T + self.br[l])).astype(f32) if self.Wr[l] is not None else y  # :172-175
        if e is not None:
            h = (h + e).astype(f32)                                      # :283-284
        return h, skips

    def _aff(self, l, x):
        """aff_res of layer l on its input (parametrized.py:44-47: x_hat.mul(a).add(b), two roundings)."""
        if self.Wa[l] is None:
            return x
        C = self.C
        p = (x @ self.Wa[l].T + self.ba[l]).astype(f32)
        return ((p[..., :C] * p[..., C:2 * C]).astype(f32) + p[..., 2 * C:]).astype(f32)

    def _dense_taps(self, l, h):
        k, d = self.kernels[l], self.dilations[l]
        n = h.shape[1] - (k - 1) * d
        return [h[:, j * d:j * d + n] for j in range(k)]

    def logits_teacher_forced(self, x):
        """x (B,T>=rf) int64 -> (B, T-rf+1, Q): entry i uses x[:, i:i+rf] and predicts sample i+rf
        (== train-mode forward of the reference)."""
        h0 = h = self.E[np.asarray(x)]
        skips = None
        for l, (k, d) in enumerate(zip(self.kernels, self.dilations)):
            taps = self._dense_taps(l, self._aff(l, h))
            sk = None if skips is None else skips[:, (k - 1) * d:]
            e = h0[:, -taps[-1].shape[1]:] if self.layerwise_inputs else None
            h, skips = self._layer(l, taps, sk, e)
        out = skips if self.has_skips else h
        return self._head(out)

    def generate(self, prompts, n_steps, temperature=None, noise=None, forced=None):
        """GenerateLoopV2.run semantics (loops/generate.py:184-229) with cached per-layer histories.
        `forced` (B, P+n_steps) teacher-forces the fed-back samples (decisions are still returned).
        Returns (sequence (B,P+n) int64, logits (B,n,Q) fp32)."""
        prompts = np.asarray(prompts, dtype=np.int64)
        B, P = prompts.shape
        if P < self.rf:
            raise ValueError(f"prompt length {P} < receptive field {self.rf}")
        T = normalize_temperature(temperature, B)
        W = self.rf
        seq = np.concatenate([prompts, np.zeros((B, n_steps), dtype=np.int64)], 1)
        # hist[l][:, j] = h_l(P - W + j): inputs of layer l; prefill densely over the last rf prompt samples
        hist = [np.zeros((B, W + n_steps, self.C), dtype=f32) for _ in range(self.L)]
        h0 = h = self.E[prompts[:, P - W:]]
        off = 0
        for l, (k, d) in enumerate(zip(self.kernels, self.dilations)):
            h = self._aff(l, h)
            hist[l][:, off:W] = h
            taps = self._dense_taps(l, h)
            h, _ = self._layer(l, taps, None, h0[:, -taps[-1].shape[1]:] if self.layerwise_inputs else None)
            off += (k - 1) * d
        logits_out = np.zeros((B, n_steps, self.Q), dtype=f32)
        for i in range(n_steps):
            t = P + i
            j = W + i - 1                      # column of time t-1
            src = seq if forced is None else np.asarray(forced)
            e = x1 = self.E[src[:, t - 1]]
            skips = None
            for l, (k, d) in enumerate(zip(self.kernels, self.dilations)):
                hist[l][:, j] = self._aff(l, x1)
                taps = [hist[l][:, j - (k - 1 - a) * d] for a in range(k)]
                x1, skips = self._layer(l, taps, skips, e if self.layerwise_inputs else None)
            out = skips if self.has_skips else x1
            lg = self._head(out)
            logits_out[:, i] = lg
            seq[:, t] = argmax_first(lg) if T is None else sample_inverse_cdf(lg, T, noise[:, i])
        return seq, logits_out


def bf16_round(a):
    """fp32 -> nearest-even bf16, returned as fp32 (what __floats2bfloat162_rn / the host packer do)."""
    a = np.ascontiguousarray(a, dtype=f32)
    u = a.view(np.uint32)
    r = ((u >> 16) & 1) + np.uint32(0x7fff)
    return ((u + r) & np.uint32(0xffff0000)).view(f32)


class WaveNetBf16Oracle(WaveNetOracle):
    """The arithmetic of the bf16 tensor-core kernel (csrc/wavenet_tc.cu) restated on the CPU: every MMA operand (weights,
    layer inputs, gated outputs, skip sum, head hidden) rounded to bf16, fp32 accumulation, the residual stream kept in
    fp32 WITHOUT the residual-conv biases (their effect enters through gate biases W1 . cumsum(b_res), computed in fp64),
    exact tanh / sigmoid / mish instead of the kernel's MUFU approximations.  Used to hold the kernel to a much tighter
    tolerance than the north star's 5e-2 against the fp32 reference."""

    def __init__(self, state_dict, blocks, **kw):
        super().__init__(state_dict, blocks, **kw)
        assert self.has_skips and all(w is not None for w in self.Wr[:-1]) and self.Wr[-1] is None
        C = self.C
        cbr = np.zeros(C, dtype=np.float64)
        self.gate_bias = []
        for l in range(self.L):
            w = self.Wd0[l].astype(np.float64) + self.Wd1[l].astype(np.float64)
            self.gate_bias.append((self.bd[l].astype(np.float64) + w @ cbr).astype(f32))
            if self.br[l] is not None:
                cbr = cbr + self.br[l].astype(np.float64)
        self.cbs = np.sum(np.stack(self.bs).astype(f32), axis=0, dtype=f32)
        q = bf16_round
        self.qWd0, self.qWd1 = [q(w) for w in self.Wd0], [q(w) for w in self.Wd1]
        self.qWs = [q(w) for w in self.Ws]
        self.qWr = [None if w is None else q(w) for w in self.Wr]
        self.qW1, self.qW2 = q(self.W1), q(self.W2)

    def logits_for(self, seq, prompt_len):
        """Teacher-forced logits (B, T - P, Q) for every position >= P of `seq` (B, T), T - P >= 1, P >= rf."""
        seq = np.asarray(seq, dtype=np.int64)
        B, T = seq.shape
        P, C, q = int(prompt_len), self.C, bf16_round
        t0 = P - self.rf                                   # first time index the layers see
        n = T - 1 - t0                                     # inputs t0 .. T-2
        hist = [np.zeros((B, n, C), dtype=f32) for _ in range(self.L)]     # bf16 layer inputs, as the rings hold them
        out = np.zeros((B, T - P, self.Q), dtype=f32)
        for i in range(n):
            H = self.E[seq[:, t0 + i]].astype(f32)         # residual stream (no residual biases)
            SK = np.zeros((B, self.Ws[0].shape[0]), dtype=f32)
            for l, d in enumerate(self.dilations):
                x1 = q(H)
                hist[l][:, i] = x1
                x0 = hist[l][:, i - d] if i - d >= 0 else np.zeros_like(x1)
                a = (x0 @ self.qWd0[l].T + x1 @ self.qWd1[l].T + self.gate_bias[l]).astype(f32)
                y = q((np.tanh(a[:, :C]) * _sigmoid(a[:, C:])).astype(f32))
                SK = (SK + y @ self.qWs[l].T).astype(f32)
                if self.qWr[l] is not None:
                    H = (H + y @ self.qWr[l].T).astype(f32)
            t = t0 + i                                     # the head predicts sample t + 1
            if t + 1 >= P:
                A = q((SK + self.cbs).astype(f32))
                hid = q(_mish((A @ self.qW1.T + self.b1).astype(f32)))
                z = (hid @ self.qW2.T + self.b2).astype(f32)
                if self.min_temp is None:                  # MLP(min_temperature=None), mlp.py:29, 54-62
                    out[:, t + 1 - P] = z
                    continue
                temp = np.maximum(_sigmoid(z[:, self.Q]), f32(self.min_temp)).astype(f32)
                out[:, t + 1 - P] = (z[:, :self.Q] / temp[:, None]).astype(f32)
        return out


# ---------------------------------------------------------------------------------------------
# SampleRNN
# ---------------------------------------------------------------------------------------------

class SampleRNNOracle:
    """Restatement of SampleRNN.before_generate / generate_step
    (networks/sample_rnn_v2.py:226-260) with SampleRNNTier.forward (83-99), FramedLinearIO / FramedConv1dIO
    (modules/io.py:106-133,185-198), LinearResampler (modules/resamplers.py:13-23) and GRU cells (PyTorch gate
    order r,z,n), LSTM cells (i,f,g,o) or tanh RNN cells, n_rnn stacked layers (sample_rnn_v2.py:62-66), the initial state of
    _init_h0 (:113-119; `h0` = {(tier, layer, which): (B, H)}, which 1 = LSTM cell state; default zeros) and the MLP head
    with n_hidden_layers shared hidden layers (networks/mlp.py:47-50).  inputs_mode='sum', single mu-law input."""

    def __init__(self, state_dict, frame_sizes, q_levels=256, rnn_class="gru", n_rnn=1, n_mlp_hidden=0):
        sd = fold_weight_norm(state_dict) if any(k.endswith("_g") for k in state_dict) else state_dict
        self.fs = tuple(int(f) for f in frame_sizes)
        self.n_tiers = len(self.fs)
        self.Q = q_levels
        self.tiers = []
        for i in range(self.n_tiers - 1):
            p = f"tiers.{i}."
            self.tiers.append(dict(
                Win=_np(sd, p + "input_module.heads.0.2.weight"), bin=_np(sd, p + "input_module.heads.0.2.bias"),
                Wih=[_np(sd, p + f"rnn.weight_ih_l{k}") for k in range(n_rnn)],
                Whh=[_np(sd, p + f"rnn.weight_hh_l{k}") for k in range(n_rnn)],
                # rnn_bias=False (sample_rnn_v2.py:66): torch registers no bias parameters; adding 0.0f changes no value
                bih=[_np(sd, p + f"rnn.bias_ih_l{k}") if p + f"rnn.bias_ih_l{k}" in sd else f32(0) for k in range(n_rnn)],
                bhh=[_np(sd, p + f"rnn.bias_hh_l{k}") if p + f"rnn.bias_hh_l{k}" in sd else f32(0) for k in range(n_rnn)],
                Wup=_np(sd, p + "up_sampler.fc.weight"), bup=_np(sd, p + "up_sampler.fc.bias")))
        p = f"tiers.{self.n_tiers - 1}.input_module.heads.0.2.2.cv."
        self.Wc = _np(sd, p + "weight")[:, 0, :]  # (H, fs_last)
        self.bc = _np(sd, p + "bias")
        self.H = self.Wc.shape[0]
        p = "output_modules.0.estimator.0."
        self.rnn_class, self.n_rnn, self.n_mlp_hidden = str(rnn_class), int(n_rnn), int(n_mlp_hidden)
        self.W1, self.b1 = _np(sd, p + "fc.0.weight"), _np(sd, p + "fc.0.bias")
        last = 2 + 2 * self.n_mlp_hidden
        self.Wh = _np(sd, p + "fc.2.weight") if self.n_mlp_hidden else None      # ONE shared Linear (mlp.py:47-50)
        self.bh = _np(sd, p + "fc.2.bias") if self.n_mlp_hidden else None
        self.W2, self.b2 = _np(sd, p + f"fc.{last}.weight"), _np(sd, p + f"fc.{last}.bias")
        self.min_temp = float(_np(sd, p + "min_temp").reshape(-1)[0]) if p + "min_temp" in sd else None
        self.rf = self.fs[0]
        # sample_rnn_v2.py:155-158
        self.up = [self.fs[i] // (self.fs[i + 1] if i < self.n_tiers - 2 else 1) for i in range(self.n_tiers - 1)]

    def lin(self, q):
        """Linearizer (modules/io.py:111-112)."""
        return ((q.astype(f32) / f32(self.Q)) - f32(0.5)) * f32(2.0)

    def _cell(self, tier, k, x, state):
        """One step of layer k: state = h (GRU, RNN) or (h, c) (LSTM).  Returns the new state."""
        H = self.H
        h = state[0] if self.rnn_class == "lstm" else state
        gi = x @ tier["Wih"][k].T + tier["bih"][k]
        gh = h @ tier["Whh"][k].T + tier["bhh"][k]
        if self.rnn_class == "gru":
            r = _sigmoid(gi[:, :H] + gh[:, :H])
            z = _sigmoid(gi[:, H:2 * H] + gh[:, H:2 * H])
            n = np.tanh(gi[:, 2 * H:] + r * gh[:, 2 * H:]).astype(f32)
            return ((1.0 - z) * n + z * h).astype(f32)
        if self.rnn_class == "lstm":
            a = (gi + gh).astype(f32)
            i, f = _sigmoid(a[:, :H]), _sigmoid(a[:, H:2 * H])
            g, o = np.tanh(a[:, 2 * H:3 * H]).astype(f32), _sigmoid(a[:, 3 * H:])
            c = (f * state[1] + i * g).astype(f32)
            return (o * np.tanh(c).astype(f32)).astype(f32), c
        return np.tanh((gi + gh).astype(f32)).astype(f32)

    def _rnn(self, tier, x, states):
        """n_rnn stacked layers: layer k reads the new hidden state of layer k - 1.  `states` is updated in place."""
        for k in range(self.n_rnn):
            states[k] = self._cell(tier, k, x, states[k])
            x = states[k][0] if self.rnn_class == "lstm" else states[k]
        return x

    def _frame_tiers(self, window, t, hid, O):
        """the `for i in range(len(tiers) - 1)` part of generate_step (sample_rnn_v2.py:245-253);
        window = the rf samples preceding (logical) time t."""
        fs = self.fs
        for i, tier in enumerate(self.tiers):
            if t % fs[i] == 0:
                x = self.lin(window[:, -fs[i]:]) @ tier["Win"].T + tier["bin"]
                if i > 0:
                    x = x + O[i - 1][:, (t // fs[i]) % (fs[i - 1] // fs[i])]
                top = self._rnn(tier, x.astype(f32), hid[i])
                O[i] = (top @ tier["Wup"].T + tier["bup"]).reshape(-1, self.up[i], self.H)

    def _initial_states(self, B, h0):
        def get(i, k, which):
            v = None if h0 is None else h0.get((i, k, which))
            return np.zeros((B, self.H), dtype=f32) if v is None else np.asarray(v, dtype=f32).copy()
        if self.rnn_class == "lstm":
            return [[(get(i, k, 0), get(i, k, 1)) for k in range(self.n_rnn)] for i in range(len(self.tiers))]
        return [[get(i, k, 0) for k in range(self.n_rnn)] for i in range(len(self.tiers))]

    def generate(self, prompts, n_steps, temperature=None, noise=None, forced=None, h0=None):
        prompts = np.asarray(prompts, dtype=np.int64)
        B, P = prompts.shape
        fs, rf = self.fs, self.rf
        if P < rf:
            raise ValueError(f"prompt length {P} < frame size {rf}")
        T = normalize_temperature(temperature, B)
        hid = self._initial_states(B, h0)
        O = [None] * len(self.tiers)
        offset = P % rf                      # sample_rnn_v2.py:229-231
        plen = P - offset
        for t in range(rf, plen):            # warm-up, :232-234
            self._frame_tiers(prompts[:, t + offset - rf:t + offset], t, hid, O)
        seq = np.concatenate([prompts, np.zeros((B, n_steps), dtype=np.int64)], 1)
        logits_out = np.zeros((B, n_steps, self.Q), dtype=f32)
        for i in range(n_steps):
            t = P + i                        # loops/generate.py:207-211: absolute t, window seq[:, t-rf:t]
            src = seq if forced is None else np.asarray(forced)
            window = src[:, t - rf:t]
            self._frame_tiers(window, t, hid, O)
            x = self.lin(window[:, -fs[-1]:]) @ self.Wc.T + self.bc          # Conv1d(1,H,k=fs_last), one frame
            x = (x + O[-1][:, (t % fs[-2]) - fs[-2]]).astype(f32)            # :256-257
            lg = mlp_head(x, self.W1, self.b1, self.W2, self.b2, self.min_temp, self.Q, self.Wh, self.bh, self.n_mlp_hidden)
            logits_out[:, i] = lg
            seq[:, t] = argmax_first(lg) if T is None else sample_inverse_cdf(lg, T, noise[:, i])
        return seq, logits_out


# ---------------------------------------------------------------------------------------------
# synthetic inputs
# ---------------------------------------------------------------------------------------------

def synthetic_waveform(B, n, sr=16000, seed=1234):
    rng = np.random.default_rng(seed)
    t = np.arange(n, dtype=np.float64) / sr
    phi = 2 * np.pi * np.arange(B)[:, None] / max(B, 1)
    x = 0.6 * np.sin(2 * np.pi * 220 * t[None] + phi) + 0.3 * np.sin(2 * np.pi * 659 * t[None]) \
        + 0.05 * rng.standard_normal((B, n))
    x = x / np.abs(x).max(axis=1, keepdims=True)
    return x.astype(f32)


def synthetic_prompts(B, P, sr=16000, seed=1234, q_levels=256):
    return mulaw_compress(synthetic_waveform(B, P, sr, seed), q_levels, 1.0)
