"""WaveNet generation on the B200 — drop-in for the reference's `WaveNet` on the generation path
(mimikit/networks/wavenet_v2.py:185-469: Config 186-206, from_config 231-257, rf 337-339, generate_params
364-366, generate_step 447-452; layers WNLayer.forward 131-176).

Same `Config` fields, same state-dict key names, same `ARM` methods; the arithmetic runs in the persistent
sm_100a kernel behind `mmk_wavenet_*` (include/mmk_b200.h).  Configurations the kernel does not implement raise
at `from_config` — there is no fallback.
"""
import ctypes
import dataclasses as dtc
import math
from collections import OrderedDict
from itertools import accumulate
from operator import mul
from typing import Optional, Tuple

import torch

from . import _capi
from .arm import NativeARM, as_temperature, prepare_noise, prepare_sequence
from .io_spec import IOSpec

__all__ = ["WaveNet"]

# MMK_ACT_* (include/mmk_b200.h): the point-wise members of ActivationEnum (modules/activations.py:26-40)
ACT_CODES = {"Tanh": 1, "Sigmoid": 2, "Mish": 3, "ReLU": 4, "Softplus": 5, "Identity": 6, "Abs": 7, "Sin": 8, "Cos": 9}


class WaveNet(NativeARM):
    @dtc.dataclass
    class Config:
        """wavenet_v2.py:186-206 — field for field."""
        io_spec: IOSpec = None
        kernel_sizes: Tuple[int, ...] = (2,)
        blocks: Tuple[int, ...] = (4,)
        dims_dilated: Tuple[int, ...] = (128,)
        dims_1x1: Tuple[int, ...] = ()
        residuals_dim: Optional[int] = None
        apply_residuals: bool = False
        skips_dim: Optional[int] = None
        with_affine_residuals: bool = False
        groups: int = 1
        act_f: str = "Tanh"
        act_g: Optional[str] = "Sigmoid"
        pad_side: int = 0
        stride: int = 1
        bias: bool = True
        use_fast_generate: bool = False
        tie_io_weights: bool = False
        layerwise_inputs: bool = False
        reverse_layer_order: bool = False

    # ------------------------------------------------------------------------------------------
    @staticmethod
    def get_kernels_and_dilation(kernel_sizes, blocks):
        """Layer schedule of wavenet_v2.py:295-327: dilations restart at 1 in every block and grow by the product
        of the kernel sizes before them."""
        ks, blocks = tuple(kernel_sizes), tuple(blocks)

        def grow(sizes):  # 1, k0, k0*k1, ... one entry per layer of a block
            return list(accumulate((1,) + tuple(sizes[:-1]), mul))

        if not blocks:
            return list(ks), list(accumulate((1,) + ks, mul))[:len(ks)]
        if len(set(blocks)) == 1 and blocks[0] == len(ks):
            return list(ks) * len(blocks), grow(ks) * len(blocks)
        if len(ks) == sum(blocks):
            dil, at = [], 0
            for b in blocks:
                dil += grow(ks[at:at + b])
                at += b
            return list(ks), dil
        if len(ks) == 1:
            return [ks[0]] * sum(blocks), [ks[0] ** i for b in blocks for i in range(b)]
        raise ValueError(f"number of layers and number of kernel sizes not compatible."
                         f" Got kernel_sizes={ks} ; blocks={blocks}")

    @classmethod
    def _check_supported(cls, c: "WaveNet.Config"):
        def need(cond, what):
            if not cond:
                raise NotImplementedError(f"mimikit_b200 WaveNet kernel: {what} is not implemented (no fallback)")
        need(c.io_spec is not None and len(c.io_spec.inputs) == 1 and len(c.io_spec.targets) == 1,
             "more than one input/target")
        need(c.io_spec.inputs[0].module_type == "embedding", "input_module_type other than 'embedding'")
        need(len(c.dims_dilated) == 1 and not c.dims_1x1, "dims_1x1 conditioning inputs")
        # residuals_dim != dims_dilated[0]: WNLayer silently builds no residual path (wavenet_v2.py:78, has_residuals); so do we
        # apply_residuals is stored by WNLayer (wavenet_v2.py:63) and read nowhere in its forward: accepted, changes nothing
        # with_affine_residuals (wavenet_v2.py:121-122, 148-149): hosted by the general fp32 kernel (aff_res stage per layer)
        need(c.groups >= 1 and c.dims_dilated[0] % c.groups == 0, "groups that do not divide the channels")
        need(str(c.act_f) in ACT_CODES and (c.act_g is None or str(c.act_g) in ACT_CODES),
             "activations outside the point-wise members of ActivationEnum (PhaseA/B/C, GLU, Softmax)")
        # bias=False (wavenet_v2.py:92-93): conv_dil / conv_skip / conv_res carry no bias; the kernels get zeros, which change no value
        need(c.pad_side in (0, 1) and c.stride == 1, "pad_side < 0 or stride != 1")
        # tie_io_weights (wavenet_v2.py:247-255) re-ties nn.Linear weights of the input module to the output module; the
        # embedding input module holds no nn.Linear, so with the only supported input type it changes nothing: accepted as a no-op
        need(0 <= c.io_spec.targets[0].module.n_hidden_layers <= 8, "more than 8 hidden MLP layers")
        # blocks=() (wavenet_v2.py:304-307, 216): `n != sum(blocks) - 1` never holds, so EVERY layer keeps its conv_res; the last
        # one's output is read by nothing when the head takes the skip sum (it is then not handed to the kernel)
        # (without skips the head then reads the last layer's x + conv_res(y): the general kernel gathers it, wavenet.cu)
        ks, _ = cls.get_kernels_and_dilation(c.kernel_sizes, c.blocks)
        need(all(2 <= k <= 4 for k in ks), "kernel sizes outside [2, 4]")

    @classmethod
    def from_config(cls, config: "WaveNet.Config") -> "WaveNet":
        cls._check_supported(config)
        return cls(config)

    def __init__(self, config: "WaveNet.Config"):
        super().__init__()
        self._check_supported(config)
        self._config = config
        ks, dil = self.get_kernels_and_dilation(config.kernel_sizes, config.blocks)
        kd = [(int(k), int(d)) for k, d in zip(ks, dil)]
        if config.reverse_layer_order:       # wavenet_v2.py:270: the ModuleList is reversed, so layers.0 is the widest dilation
            kd.reverse()                     # and the layer built WITHOUT conv_res (wavenet_v2.py:216) runs first
        self.kernels = [k for k, _ in kd]
        self.dilations = [d for _, d in kd]
        self.has_skips = config.skips_dim is not None
        self.has_residuals = config.residuals_dim is not None and config.residuals_dim == config.dims_dilated[0]   # wavenet_v2.py:78
        self._sd = self._init_state_dict()
        self._gen = None  # state of the step-wise protocol
        self._cont = None  # where the last generate() stopped (generate_more continues from the rings as they are)
        self._compute_dtype = torch.float32

    # ---- arithmetic ------------------------------------------------------------------------------
    @property
    def compute_dtype(self):
        """torch.float32 (default): fp32 FFMA kernels, sequences bit-exact with the oracle.
        torch.bfloat16: the tensor-core kernel (tcgen05.mma, TMEM accumulators; csrc/wavenet_tc.cu) — bf16 operands,
        fp32 accumulation, logits within 5e-2 relative; one CTA per 128 prompts, meant for large batches."""
        return self._compute_dtype

    @compute_dtype.setter
    def compute_dtype(self, dtype):
        if dtype not in (torch.float32, torch.bfloat16):
            raise ValueError("compute_dtype must be torch.float32 or torch.bfloat16")
        if dtype != self._compute_dtype:
            self._release()
            self._compute_dtype = dtype

    def bfloat16(self):
        self.compute_dtype = torch.bfloat16
        return self

    def float(self):
        self.compute_dtype = torch.float32
        return self

    # ---- geometry -------------------------------------------------------------------------------
    @property
    def config(self):
        return self._config

    @property
    def rf(self) -> int:
        """wavenet_v2.py:337-339: sum of the layers' causes (kernel_size - 1) * dilation, + 1."""
        return sum((k - 1) * d for k, d in zip(self.kernels, self.dilations)) + 1

    @property
    def shift(self) -> int:
        """wavenet_v2.py:333-335."""
        return 1 if self._config.pad_side == 1 else self.rf

    def output_length(self, n_input_steps: int) -> int:
        """wavenet_v2.py:341-342."""
        return n_input_steps if self._config.pad_side != 0 else n_input_steps - self.shift + 1

    @property
    def _n_mlp_hidden(self):
        return int(self._config.io_spec.targets[0].module.n_hidden_layers)

    @property
    def generate_params(self):
        """The reference yields {} here (wavenet_v2.py:364-366 looks `sampling_params` up on an nn.ModuleList), so
        its loop can never pass a temperature to WaveNet; the evident intent — and SampleRNN's behaviour
        (sample_rnn_v2.py:309-311) — is {"temperature"}.  Documented deviation."""
        return {"temperature"}

    @property
    def _gated(self):
        return self._config.act_g is not None

    def _layer_has_res(self, l):
        """wavenet_v2.py:216 — the layer BUILT last has no conv_res; with reverse_layer_order it is executed first."""
        L = len(self.dilations)
        if not self._config.blocks:          # sum(()) - 1 == -1: no layer is the "last" one, all of them keep conv_res
            return self.has_residuals
        return self.has_residuals and l != (0 if self._config.reverse_layer_order else L - 1)

    def _dims(self):
        c = self._config
        C = c.dims_dilated[0]
        S = c.skips_dim if self.has_skips else 0
        head = c.io_spec.targets[0].module
        return C, S, head.hidden_dim, c.io_spec.targets[0].out_dim

    def _expected_shapes(self):
        e = self._all_shapes()
        if not self._config.bias:
            for k in self._absent_biases(e):
                del e[k]
        return e

    @staticmethod
    def _absent_biases(e):
        """bias=False reaches the three convs of a layer (wavenet_v2.py:92-93); aff_res keeps its own bias (:122)."""
        return [k for k in e if k.startswith("layers.") and k.endswith(".bias") and ".aff_res." not in k]

    def _all_shapes(self):
        C, S, Hh, Q = self._dims()
        L = len(self.dilations)
        e = OrderedDict()
        e["input_modules.0.0.weight"] = (self._config.io_spec.inputs[0].class_size, C)
        for l in range(L):
            Cg = C // self._config.groups               # grouped dilated convs (wavenet_v2.py:92): weight (out, in / groups, k)
            if self._gated:
                e[f"layers.{l}.conv_dil.0.0.weight"] = (2 * C, Cg, self.kernels[l])
                e[f"layers.{l}.conv_dil.0.0.bias"] = (2 * C,)
            else:                                       # wavenet_v2.py:109-112: a bare Conv1d (no Sequential / Chunk) without gated units
                e[f"layers.{l}.conv_dil.0.weight"] = (C, Cg, self.kernels[l])
                e[f"layers.{l}.conv_dil.0.bias"] = (C,)
            if self.has_skips:
                e[f"layers.{l}.conv_skip.weight"] = (S, C, 1)
                e[f"layers.{l}.conv_skip.bias"] = (S,)
            if self._layer_has_res(l):
                e[f"layers.{l}.conv_res.weight"] = (C, C, 1)
                e[f"layers.{l}.conv_res.bias"] = (C,)
            if self._config.with_affine_residuals:      # wavenet_v2.py:121-122: ParametrizedLinear(C, C, as_1x1_conv=True)
                e[f"layers.{l}.aff_res.params.weight"] = (3 * C, C, 1)
                e[f"layers.{l}.aff_res.params.bias"] = (3 * C,)
        p = "output_modules.0.estimator.0."
        if self._learns_temperature:                      # mlp.py:29, 54-57: without it, Q outputs and no buffer
            e[p + "min_temp"] = ()
        e[p + "fc.0.weight"] = (Hh, S if self.has_skips else C)
        e[p + "fc.0.bias"] = (Hh,)
        nh = self._n_mlp_hidden                           # mlp.py:47-50: fc.2, fc.4, ... are ONE shared Linear(Hh, Hh)
        for r in range(nh):
            e[p + f"fc.{2 + 2 * r}.weight"] = (Hh, Hh)
            e[p + f"fc.{2 + 2 * r}.bias"] = (Hh,)
        e[p + f"fc.{2 + 2 * nh}.weight"] = (Q + int(self._learns_temperature), Hh)
        e[p + f"fc.{2 + 2 * nh}.bias"] = (Q + int(self._learns_temperature),)
        return e

    def _init_state_dict(self):
        """Random init with torch's default distributions (Embedding N(0,1); Conv/Linear U(+-1/sqrt(fan_in)))."""
        sd = OrderedDict()
        for k, shape in self._expected_shapes().items():
            if k.endswith("min_temp"):
                mt = self._config.io_spec.targets[0].module.min_temperature
                sd[k] = torch.tensor(1e-4 if mt is None else float(mt), dtype=torch.float32)
            elif k == "input_modules.0.0.weight":
                sd[k] = torch.randn(shape)
            else:
                wshape = shape if k.endswith("weight") else self._expected_shapes()[k[:-4] + "weight"]
                bound = 1.0 / math.sqrt(max(1, int(torch.tensor(wshape[1:]).prod())))
                sd[k] = (torch.rand(shape) * 2 - 1) * bound
        p = "output_modules.0.estimator.0."
        for r in range(1, self._n_mlp_hidden):
            sd[p + f"fc.{2 + 2 * r}.weight"] = sd[p + "fc.2.weight"]
            sd[p + f"fc.{2 + 2 * r}.bias"] = sd[p + "fc.2.bias"]
        return sd

    # ---- native handle --------------------------------------------------------------------------
    def _create_handle(self, max_batch):
        C, S, Hh, Q = self._dims()
        L = len(self.dilations)
        e = self._all_shapes()
        self._absent = {} if self._config.bias else {k: torch.zeros(e[k], dtype=torch.float32) for k in self._absent_biases(e)}
        dx = _capi.WaveNetDescEx()
        d = dx.base
        d.n_layers, d.dilated_dim, d.skips_dim, d.head_hidden, d.q_levels = L, C, S, Hh, Q
        dil = (ctypes.c_int * L)(*self.dilations)
        d.dilations = dil
        d.embedding = self._w("input_modules.0.0.weight")
        keep = [dil]
        def arr(fmt, present=lambda l: True):
            a = self._warray([fmt.format(l) if present(l) else None for l in range(L)])
            keep.append(a)
            return a
        if self._gated and self._config.groups == 1:
            d.conv_dil_w = arr("layers.{}.conv_dil.0.0.weight")
            d.conv_dil_b = arr("layers.{}.conv_dil.0.0.bias")
        else:
            # Two forms are hosted by the dense gated kernel through their weights alone:
            #  * groups > 1 (wavenet_v2.py:92): the block-diagonal conv written out densely — the zero blocks add exact zeros;
            #  * act_g=None (wavenet_v2.py:160-163): y = tanh(conv(x)) as a gated unit whose gate is exactly one: zero gate weights
            #    and a gate bias of 40 — sigmoid(40) = 1 / (1 + exp(-40)) rounds to 1.0f, and tanh(a) * 1.0f = tanh(a).
            self._dense_pack = {}
            G = int(self._config.groups)
            for l in range(L):
                pre = f"layers.{l}.conv_dil.0.0." if self._gated else f"layers.{l}.conv_dil.0."
                w, b = self._sd[pre + "weight"], self._sd[pre + "bias"] if self._config.bias else self._absent[pre + "bias"]
                if G > 1:
                    O, Cg, K = w.shape
                    dense = torch.zeros((O, C, K), dtype=w.dtype)
                    for gi in range(G):
                        dense[gi * (O // G):(gi + 1) * (O // G), gi * Cg:(gi + 1) * Cg] = w[gi * (O // G):(gi + 1) * (O // G)]
                    w = dense
                if not self._gated:
                    w = torch.cat([w, torch.zeros_like(w)], 0)
                    b = torch.cat([b, torch.full_like(b, 40.0)], 0)
                self._dense_pack[f"w{l}"], self._dense_pack[f"b{l}"] = w.contiguous(), b.contiguous()
            wa = (ctypes.POINTER(ctypes.c_float) * L)(*[_capi.fptr(self._dense_pack[f"w{l}"]) for l in range(L)])
            ba = (ctypes.POINTER(ctypes.c_float) * L)(*[_capi.fptr(self._dense_pack[f"b{l}"]) for l in range(L)])
            keep += [wa, ba]
            d.conv_dil_w, d.conv_dil_b = wa, ba
        if self.has_skips:
            d.conv_skip_w = arr("layers.{}.conv_skip.weight")
            d.conv_skip_b = arr("layers.{}.conv_skip.bias")
        # reverse_layer_order: the layer executed last carries a conv_res whose result nothing reads when the head takes the
        # skip sum (wavenet_v2.py:286-292): it is not handed to the kernel
        has_res = lambda l: self._layer_has_res(l) and (l != L - 1 or not self.has_skips)
        d.conv_res_w = arr("layers.{}.conv_res.weight", has_res)
        d.conv_res_b = arr("layers.{}.conv_res.bias", has_res)
        p = "output_modules.0.estimator.0."
        nh = self._n_mlp_hidden
        d.head_w1, d.head_b1 = self._w(p + "fc.0.weight"), self._w(p + "fc.0.bias")
        d.head_w2, d.head_b2, d.min_temperature = self._head_last(p + f"fc.{2 + 2 * nh}.weight", p + f"fc.{2 + 2 * nh}.bias")
        dx.act_f = ACT_CODES[str(self._config.act_f)]
        dx.act_g = ACT_CODES[str(self._config.act_g)] if self._gated else ACT_CODES["Sigmoid"]   # not gated: the exact-one gate below
        if self._config.with_affine_residuals:
            dx.aff_res_w = arr("layers.{}.aff_res.params.weight")
            dx.aff_res_b = arr("layers.{}.aff_res.params.bias")
        ks = (ctypes.c_int * L)(*self.kernels)
        keep.append(ks)
        dx.kernel_sizes = ks
        dx.layerwise_inputs = int(bool(self._config.layerwise_inputs))
        dx.head_hidden_layers = nh
        if nh > 0:
            for r in range(1, nh):
                if not (torch.equal(self._sd[p + f"fc.{2 + 2 * r}.weight"], self._sd[p + "fc.2.weight"])
                        and torch.equal(self._sd[p + f"fc.{2 + 2 * r}.bias"], self._sd[p + "fc.2.bias"])):
                    raise RuntimeError("the hidden layers of the reference MLP share one Linear (networks/mlp.py:47-50): "
                                       "fc.2, fc.4, ... must hold the same tensors")
            dx.head_wh, dx.head_bh = self._w(p + "fc.2.weight"), self._w(p + "fc.2.bias")
        h = ctypes.c_void_p()
        mode = 1 if self._compute_dtype == torch.bfloat16 else 0      # MMK_COMPUTE_BF16_TC / MMK_COMPUTE_FP32
        _capi.check(_capi.lib().mmk_wavenet_create_cfg(ctypes.byref(dx), int(max_batch), mode, ctypes.byref(h)))
        return h

    def _destroy_handle(self, h):
        _capi.lib().mmk_wavenet_destroy(h)

    def launch_info(self, batch=1):
        info = _capi.LaunchInfo()
        _capi.check(_capi.lib().mmk_wavenet_launch_info(self._get_handle(batch), ctypes.byref(info)))
        return {f: getattr(info, f) for f, _ in info._fields_}

    def _run(self, seq, seq_t0, t_begin, t_head, t_end, teacher_forced, temperature, noise, noise_t0, want_logits,
             want_decisions, want_ts, check=True):
        B = seq.shape[0]
        h = self._get_handle(B)
        n_head = max(0, t_end - t_head)
        Q = self._dims()[3]
        logits = torch.empty((B, n_head, Q), dtype=torch.float32, device=seq.device) if want_logits else None
        decisions = torch.empty((B, n_head), dtype=torch.int64, device=seq.device) if want_decisions else None
        ts = torch.zeros((t_end - t_begin,), dtype=torch.int64, device=seq.device) if want_ts else None
        ptr = lambda x: x.data_ptr() if x is not None else None
        with torch.cuda.device(seq.device):
            _capi.check(_capi.lib().mmk_wavenet_run(
                h, seq.data_ptr(), B, seq.stride(0), int(seq_t0), int(t_begin), int(t_head), int(t_end), int(teacher_forced),
                ptr(temperature), 0 if temperature is None else temperature.numel(),
                ptr(noise), 0 if noise is None else noise.stride(0), int(noise_t0),
                ptr(logits), ptr(decisions), ptr(ts), _capi.stream_ptr()))
            if check:
                _capi.check(_capi.lib().mmk_wavenet_sync_check(h, _capi.stream_ptr()))
        return logits, decisions, ts

    # ---- whole-sequence fast path ---------------------------------------------------------------
    def generate(self, prompts, n_steps, temperature=None, noise=None, return_logits=False,
                 return_step_timestamps=False, generator=None):
        """GenerateLoopV2.run's inner loop for one batch (loops/generate.py:195-219) in ONE kernel launch.

        prompts (B, P >= rf) int64 mu-law indices (host or device); temperature None (argmax) | float | (1,) | (B,);
        noise (B, n_steps) uniform [0,1) fp32 consumed by the inverse-CDF sampler (drawn with torch.rand when
        omitted).  Returns the (B, P + n_steps) int64 sequence on the device (plus logits (B, n_steps, Q) and/or
        per-step device timestamps in ns when asked)."""
        seq = prepare_sequence(prompts, n_steps, self.device)
        B, total = seq.shape
        P = total - n_steps
        if P < self.rf:
            raise RuntimeError(f"prompt length {P} is shorter than the receptive field {self.rf}")
        T = as_temperature(temperature, B, self.device)
        U = prepare_noise(noise, T, B, n_steps, self.device, generator)
        logits, ts = None, None
        if n_steps > 0:
            # layers consume samples P-rf .. P+n-2; the head predicts sample t+1 for t >= P-1
            logits, _, ts = self._run(seq, 0, P - self.rf, P - 1, P + n_steps - 1, False, T, U, P, return_logits, False,
                                      return_step_timestamps)
            self._cont = dict(handle=self._handle, B=B, t=P + n_steps - 1, last=seq[:, -1].clone())
        out = (seq,)
        if return_logits:
            out += (logits,)
        if return_step_timestamps:
            out += (ts,)
        return out[0] if len(out) == 1 else out

    def generate_more(self, n_steps, temperature=None, noise=None, return_logits=False, generator=None):
        """Continue the last `generate` / `generate_more` of this network for the SAME batch without re-prompting: the
        dilation rings in the native handle still hold every layer's last `dilation` inputs, so the kernel is launched
        with an empty prefill and its clock simply runs on (chunked long-form generation, loops/generate_chunks.py:39-56,
        minus the window recompute of each re-prompt).  Returns the (B, n_steps) new samples (and their logits)."""
        c = self._cont
        if c is None or c["handle"] is not self._handle or self._handle is None:
            raise RuntimeError("generate_more() continues a previous generate() on the same network: nothing to continue "
                               "(the native handle was rebuilt or generate() was never called)")
        B, t = c["B"], c["t"]
        T = as_temperature(temperature, B, self.device)
        U = prepare_noise(noise, T, B, n_steps, self.device, generator)
        buf = torch.zeros((B, n_steps + 1), dtype=torch.int64, device=self.device)
        buf[:, 0] = c["last"]
        logits = None
        if n_steps > 0:
            # column 0 holds the newest known sample (time t): the layers consume t .. t+n-1, the head predicts t+1 .. t+n
            logits, _, _ = self._run(buf, t, t, t, t + n_steps, False, T, U, t + 1, return_logits, False, False)
            self._cont = dict(handle=self._handle, B=B, t=t + n_steps, last=buf[:, -1].clone())
        return (buf[:, 1:], logits) if return_logits else buf[:, 1:]

    def teacher_forced(self, sequence, prompt_len, temperature=None, noise=None):
        """Logits and decisions for every position >= prompt_len of a GIVEN sequence (nothing is fed back).
        Returns (logits (B, n, Q), decisions (B, n)) with n = T - prompt_len."""
        seq = prepare_sequence(sequence, 0, self.device)
        B, total = seq.shape
        P = int(prompt_len)
        if P < self.rf:
            raise RuntimeError(f"prompt length {P} is shorter than the receptive field {self.rf}")
        n = total - P
        T = as_temperature(temperature, B, self.device)
        U = prepare_noise(noise, T, B, n, self.device)
        logits, dec, _ = self._run(seq, 0, P - self.rf, P - 1, total - 1, True, T, U, P, True, True, False)
        return logits, dec

    # ---- step-wise ARM protocol (arm.py:56-75) --------------------------------------------------
    def before_generate(self, prompts, batch_index) -> None:
        self._gen = None

    def generate_step(self, inputs, *, t: int = 0, temperature=None, noise=None):
        """One sample from the window `inputs[0]` = the rf samples before t.  The first call (or any call whose t
        does not follow the previous one) refills the dilation rings from the whole window — the reference's
        window recompute; consecutive calls only push the newest sample through the cached step."""
        x = inputs[0]
        if x.dim() != 2 or x.shape[1] < self.rf:
            raise RuntimeError(f"expected a (B, >= {self.rf}) window, got {tuple(x.shape)}")
        B, rf = x.shape[0], self.rf
        x = x.to(self.device, torch.int64)
        T = as_temperature(temperature, B, self.device)
        U = None
        if T is not None:
            U = torch.rand((B, 1), device=self.device) if noise is None else \
                torch.as_tensor(noise, dtype=torch.float32).reshape(B, 1).to(self.device)
        g = self._gen
        if g is None or g["t"] + 1 != t or g["buf"].shape[0] != B:
            buf = torch.zeros((B, rf + 1), dtype=torch.int64, device=self.device)
            buf[:, :rf] = x[:, -rf:]
            self._run(buf, 0, 0, rf - 1, rf, False, T, U, rf, False, False, False)
            # the kernel's clock (which selects the ring slots, t mod dilation) started at 0 for sample t - rf
            self._gen = dict(t=t, origin=t - rf, buf=buf, step=torch.zeros((B, 2), dtype=torch.int64,
                                                                             device=self.device))
            return (buf[:, rf:rf + 1].clone(),)
        step, tl = g["step"], t - 1 - g["origin"]   # kernel time of the newest sample
        step[:, 0] = x[:, -1]
        self._run(step, tl, tl, tl, tl + 1, False, T, U, tl + 1, False, False, False)
        g["t"] = t
        return (step[:, 1:2].clone(),)

    def after_generate(self, final_outputs, batch_index) -> None:
        self._gen = None
