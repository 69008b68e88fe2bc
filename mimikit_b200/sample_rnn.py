"""SampleRNN generation on the B200 — drop-in for the reference's `SampleRNN` on the generation path
(mimikit/networks/sample_rnn_v2.py:122-317: Config 123-134, from_config 136-186, before_generate 226-234,
generate_step 236-260, reset_hidden 266-268, rf 274-276, generate_params 309-311).

Same `Config` fields, state-dict key names (weight_norm'ed checkpoints are folded on load) and `ARM` methods; the
arithmetic runs in the persistent sm_100a kernel behind `mmk_samplernn_*` (include/mmk_b200.h).
"""
import ctypes
import dataclasses as dtc
import math
from collections import OrderedDict
from typing import Tuple

import torch

from . import _capi
from .arm import NativeARM, as_temperature, prepare_noise, prepare_sequence
from .io_spec import IOSpec

__all__ = ["SampleRNN"]


class SampleRNN(NativeARM):
    @dtc.dataclass
    class Config:
        """sample_rnn_v2.py:123-134 — field for field."""
        frame_sizes: Tuple[int, ...] = (16, 8, 8)
        hidden_dim: int = 256
        rnn_class: str = "lstm"
        n_rnn: int = 1
        rnn_dropout: float = 0.
        rnn_bias: bool = True
        h0_init: str = "zeros"
        weight_norm: bool = False
        inputs_mode: str = "sum"
        io_spec: IOSpec = None

    @staticmethod
    def _check_supported(c: "SampleRNN.Config"):
        def need(cond, what):
            if not cond:
                raise NotImplementedError(f"mimikit_b200 SampleRNN kernel: {what} is not implemented (no fallback)")
        need(c.io_spec is not None and len(c.io_spec.inputs) == 1 and len(c.io_spec.targets) == 1,
             "more than one input/target")
        need(c.io_spec.inputs[0].module_type == "framed_linear", "input_module_type other than 'framed_linear'")
        need(str(c.rnn_class) == "gru", "rnn_class other than 'gru'")
        need(c.n_rnn == 1 and c.rnn_bias, "n_rnn > 1 or rnn_bias=False")
        need(str(c.h0_init) == "zeros", "h0_init other than 'zeros'")
        need(str(c.inputs_mode) == "sum", "inputs_mode other than 'sum'")
        need(len(c.frame_sizes) >= 2, "fewer than two tiers")
        need(c.io_spec.targets[0].module.n_hidden_layers == 0, "n_mlp_layers > 0")
        fs = c.frame_sizes
        for i in range(len(fs) - 2):
            need(fs[i] % fs[i + 1] == 0, "frame sizes that do not divide each other")

    @classmethod
    def from_config(cls, config: "SampleRNN.Config") -> "SampleRNN":
        return cls(config)

    def __init__(self, config: "SampleRNN.Config"):
        super().__init__()
        self._check_supported(config)
        self._config = config
        self.frame_sizes = tuple(int(f) for f in config.frame_sizes)
        self._sd = self._init_state_dict()
        self._prompt_len = None

    @property
    def config(self):
        return self._config

    @property
    def rf(self):
        """sample_rnn_v2.py:274-276."""
        return self.frame_sizes[0]

    @property
    def generate_params(self):
        """sample_rnn_v2.py:309-311."""
        return {"temperature"}

    def _dims(self):
        c = self._config
        head = c.io_spec.targets[0].module
        return c.hidden_dim, head.hidden_dim, c.io_spec.targets[0].out_dim

    def _up(self, i):
        """sample_rnn_v2.py:155-158: the last frame tier up-samples to the sample rate."""
        fs = self.frame_sizes
        return fs[i] // (fs[i + 1] if i < len(fs) - 2 else 1)

    def _expected_shapes(self):
        H, Hh, Q = self._dims()
        fs = self.frame_sizes
        e = OrderedDict()
        for i in range(len(fs) - 1):
            p = f"tiers.{i}."
            e[p + "input_module.heads.0.2.weight"] = (H, fs[i])
            e[p + "input_module.heads.0.2.bias"] = (H,)
            e[p + "rnn.weight_ih_l0"] = (3 * H, H)
            e[p + "rnn.weight_hh_l0"] = (3 * H, H)
            e[p + "rnn.bias_ih_l0"] = (3 * H,)
            e[p + "rnn.bias_hh_l0"] = (3 * H,)
            e[p + "up_sampler.fc.weight"] = (H * self._up(i), H)
            e[p + "up_sampler.fc.bias"] = (H * self._up(i),)
        p = f"tiers.{len(fs) - 1}.input_module.heads.0.2.2.cv."
        e[p + "weight"] = (H, 1, fs[-1])
        e[p + "bias"] = (H,)
        p = "output_modules.0.estimator.0."
        e[p + "min_temp"] = ()
        e[p + "fc.0.weight"] = (Hh, H)
        e[p + "fc.0.bias"] = (Hh,)
        e[p + "fc.2.weight"] = (Q + 1, Hh)
        e[p + "fc.2.bias"] = (Q + 1,)
        return e

    def _init_state_dict(self):
        """torch's default distributions: Linear/Conv U(+-1/sqrt(fan_in)), GRU U(+-1/sqrt(H))."""
        H = self._config.hidden_dim
        shapes = self._expected_shapes()
        sd = OrderedDict()
        for k, shape in shapes.items():
            if k.endswith("min_temp"):
                mt = self._config.io_spec.targets[0].module.min_temperature
                sd[k] = torch.tensor(1e-4 if mt is None else float(mt), dtype=torch.float32)
                continue
            if ".rnn." in k:
                bound = 1.0 / math.sqrt(H)
            else:
                wshape = shape if k.endswith("weight") else shapes[k[:-4] + "weight"]
                bound = 1.0 / math.sqrt(max(1, int(torch.tensor(wshape[1:]).prod())))
            sd[k] = (torch.rand(shape) * 2 - 1) * bound
        return sd

    # ---- native handle --------------------------------------------------------------------------
    def _create_handle(self, max_batch):
        H, Hh, Q = self._dims()
        fs = self.frame_sizes
        n = len(fs)
        d = _capi.SampleRNNDesc()
        d.n_tiers, d.hidden_dim, d.head_hidden, d.q_levels = n, H, Hh, Q
        fsa = (ctypes.c_int * n)(*fs)
        d.frame_sizes = fsa
        d.min_temperature = float(self._sd["output_modules.0.estimator.0.min_temp"])
        keep = [fsa]
        def arr(fmt):
            a = self._warray([fmt.format(i) for i in range(n - 1)])
            keep.append(a)
            return a
        d.in_w, d.in_b = arr("tiers.{}.input_module.heads.0.2.weight"), arr("tiers.{}.input_module.heads.0.2.bias")
        d.w_ih, d.w_hh = arr("tiers.{}.rnn.weight_ih_l0"), arr("tiers.{}.rnn.weight_hh_l0")
        d.b_ih, d.b_hh = arr("tiers.{}.rnn.bias_ih_l0"), arr("tiers.{}.rnn.bias_hh_l0")
        d.up_w, d.up_b = arr("tiers.{}.up_sampler.fc.weight"), arr("tiers.{}.up_sampler.fc.bias")
        p = f"tiers.{n - 1}.input_module.heads.0.2.2.cv."
        d.conv_w, d.conv_b = self._w(p + "weight"), self._w(p + "bias")
        p = "output_modules.0.estimator.0."
        d.head_w1, d.head_b1 = self._w(p + "fc.0.weight"), self._w(p + "fc.0.bias")
        d.head_w2, d.head_b2 = self._w(p + "fc.2.weight"), self._w(p + "fc.2.bias")
        h = ctypes.c_void_p()
        _capi.check(_capi.lib().mmk_samplernn_create(ctypes.byref(d), int(max_batch), ctypes.byref(h)))
        return h

    def _destroy_handle(self, h):
        _capi.lib().mmk_samplernn_destroy(h)

    def launch_info(self, batch=1):
        info = _capi.LaunchInfo()
        _capi.check(_capi.lib().mmk_samplernn_launch_info(self._get_handle(batch), ctypes.byref(info)))
        return {f: getattr(info, f) for f, _ in info._fields_}

    def _run(self, seq, seq_t0, warm, gen, reset, teacher_forced, temperature, noise, noise_t0, want_logits,
             want_decisions, want_ts):
        B = seq.shape[0]
        h = self._get_handle(B)
        n_gen = max(0, gen[1] - gen[0])
        Q = self._dims()[2]
        logits = torch.empty((B, n_gen, Q), dtype=torch.float32, device=seq.device) if want_logits else None
        decisions = torch.empty((B, n_gen), dtype=torch.int64, device=seq.device) if want_decisions else None
        ts = torch.zeros((n_gen,), dtype=torch.int64, device=seq.device) if want_ts else None
        ptr = lambda x: x.data_ptr() if x is not None else None
        with torch.cuda.device(seq.device):
            _capi.check(_capi.lib().mmk_samplernn_run(
                h, seq.data_ptr(), B, seq.stride(0), int(seq_t0), int(warm[0]), int(warm[1]), int(warm[2]),
                int(gen[0]), int(gen[1]), int(reset), int(teacher_forced),
                ptr(temperature), 0 if temperature is None else temperature.numel(),
                ptr(noise), 0 if noise is None else noise.stride(0), int(noise_t0),
                ptr(logits), ptr(decisions), ptr(ts), _capi.stream_ptr()))
            _capi.check(_capi.lib().mmk_samplernn_sync_check(h, _capi.stream_ptr()))
        return logits, decisions, ts

    def _warm_range(self, P):
        """before_generate (sample_rnn_v2.py:229-234): logical t in [rf, P - P % rf), data shifted by P % rf."""
        offset = P % self.rf
        return (self.rf, P - offset, offset)

    # ---- whole-sequence fast path ---------------------------------------------------------------
    def generate(self, prompts, n_steps, temperature=None, noise=None, return_logits=False,
                 return_step_timestamps=False, generator=None):
        """before_generate (hidden reset + warm-up over the prompt) followed by n_steps of generate_step as
        GenerateLoopV2.run drives them (loops/generate.py:193-219), in ONE kernel launch.  Arguments and returns as
        `WaveNet.generate`."""
        seq = prepare_sequence(prompts, n_steps, self.device)
        B, total = seq.shape
        P = total - n_steps
        if P < self.rf:
            raise RuntimeError(f"prompt length {P} is shorter than the top frame size {self.rf}")
        T = as_temperature(temperature, B, self.device)
        U = prepare_noise(noise, T, B, n_steps, self.device, generator)
        logits, _, ts = self._run(seq, 0, self._warm_range(P), (P, P + n_steps), True, False, T, U, P,
                                  return_logits, False, return_step_timestamps)
        self._cont = dict(handle=self._handle, B=B, t=P + n_steps, tail=seq[:, -self.rf:].clone())
        out = (seq,)
        if return_logits:
            out += (logits,)
        if return_step_timestamps:
            out += (ts,)
        return out[0] if len(out) == 1 else out

    def generate_more(self, n_steps, temperature=None, noise=None, return_logits=False, generator=None):
        """Continue the last `generate` / `generate_more` for the SAME batch with the GRU states as they are (no hidden
        reset, no warm-up over a re-prompt): chunked long-form generation (loops/generate_chunks.py:39-56) as one
        uninterrupted sequence.  Returns the (B, n_steps) new samples (and their logits)."""
        c = getattr(self, "_cont", None)
        if c is None or c["handle"] is not self._handle or self._handle is None:
            raise RuntimeError("generate_more() continues a previous generate() on the same network: nothing to continue")
        B, t, rf = c["B"], c["t"], self.rf
        T = as_temperature(temperature, B, self.device)
        U = prepare_noise(noise, T, B, n_steps, self.device, generator)
        buf = torch.zeros((B, rf + n_steps), dtype=torch.int64, device=self.device)
        buf[:, :rf] = c["tail"]
        logits = None
        if n_steps > 0:
            logits, _, _ = self._run(buf, t - rf, (0, 0, 0), (t, t + n_steps), False, False, T, U, t, return_logits, False, False)
            self._cont = dict(handle=self._handle, B=B, t=t + n_steps, tail=buf[:, -rf:].clone())
        return (buf[:, rf:], logits) if return_logits else buf[:, rf:]

    def teacher_forced(self, sequence, prompt_len, temperature=None, noise=None):
        """Step-wise generate_step logits/decisions on forced inputs (this, not SampleRNN.forward,
        is the teacher-forced definition for SampleRNN)."""
        seq = prepare_sequence(sequence, 0, self.device)
        B, total = seq.shape
        P = int(prompt_len)
        if P < self.rf:
            raise RuntimeError(f"prompt length {P} is shorter than the top frame size {self.rf}")
        n = total - P
        T = as_temperature(temperature, B, self.device)
        U = prepare_noise(noise, T, B, n, self.device)
        logits, dec, _ = self._run(seq, 0, self._warm_range(P), (P, total), True, True, T, U, P, True, True, False)
        return logits, dec

    # ---- step-wise ARM protocol -------------------------------------------------------------------
    def reset_hidden(self):
        self._prompt_len = None

    def before_generate(self, prompts, batch_index) -> None:
        """sample_rnn_v2.py:226-234: reset the hidden states and warm the frame tiers up over the prompt."""
        p = prepare_sequence(prompts[0], 0, self.device)
        P = p.shape[1]
        if P < self.rf:
            raise RuntimeError(f"prompt length {P} is shorter than the top frame size {self.rf}")
        self._run(p, 0, self._warm_range(P), (P, P), True, True, None, None, 0, False, False, False)
        self._prompt_len = P

    def generate_step(self, inputs, *, t: int = 0, temperature=None, noise=None):
        """sample_rnn_v2.py:236-260 for t >= prompt length (the warm-up calls are made by before_generate)."""
        x = inputs[0].to(self.device, torch.int64).contiguous()
        B, rf = x.shape[0], self.rf
        if x.shape[1] < rf:
            raise RuntimeError(f"expected a (B, >= {rf}) window, got {tuple(x.shape)}")
        T = as_temperature(temperature, B, self.device)
        U = None
        if T is not None:
            U = torch.rand((B, 1), device=self.device) if noise is None else \
                torch.as_tensor(noise, dtype=torch.float32).reshape(B, 1).to(self.device)
        buf = torch.zeros((B, rf + 1), dtype=torch.int64, device=self.device)
        buf[:, :rf] = x[:, -rf:]
        self._run(buf, t - rf, (0, 0, 0), (t, t + 1), False, False, T, U, t, False, False, False)
        return (buf[:, rf:rf + 1],)

    def after_generate(self, final_outputs, batch_index) -> None:
        self._prompt_len = None
