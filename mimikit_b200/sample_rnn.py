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
        need(str(c.rnn_class) in ("gru", "lstm", "rnn"), "rnn_class other than 'lstm', 'gru' or 'rnn'")
        need(1 <= c.n_rnn <= 4, "n_rnn outside [1, 4]")     # rnn_bias=False: no rnn.bias_* parameters; the kernels get zero biases
        need(str(c.h0_init) in ("zeros", "ones", "randn"), "h0_init other than 'zeros', 'ones' or 'randn'")
        # ZipReduceVariables (modules/io.py:289-313) over the ONE input of the mu-law path: "sum" and "mean" weigh it by 1 and 1 / 1,
        # "static_mix" by the softmax of a one-element parameter — exactly 1.0f in all three; static_mix adds `input_module.weights`
        need(str(c.inputs_mode) in ("sum", "mean", "static_mix"), "inputs_mode other than 'sum', 'mean' or 'static_mix'")
        need(len(c.frame_sizes) >= 2, "fewer than two tiers")
        need(0 <= c.io_spec.targets[0].module.n_hidden_layers <= 8, "more than 8 hidden MLP layers")
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
        self._compute_dtype = torch.float32

    @property
    def config(self):
        return self._config

    # ---- arithmetic ------------------------------------------------------------------------------
    @property
    def compute_dtype(self):
        """torch.float32 (default): fp32 FFMA kernels, sequences bit-exact with the oracle.
        torch.bfloat16: the frame tiers' GRU and up-sampler contractions run on the tensor cores (tcgen05.mma, accumulators in
        TMEM; csrc/samplernn2.cu) with bf16 operands and fp32 accumulation; the cell, the head and the sampler stay fp32.
        Logits within 5e-2 relative.  GRU tiers, one layer, zero initial state, hidden_dim in {128, 256, 512}, <= 128 prompts."""
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

    @property
    def _gates(self):
        """Rows of weight_ih / weight_hh per hidden unit: nn.GRU 3 (r, z, n), nn.LSTM 4 (i, f, g, o), nn.RNN 1."""
        return {"gru": 3, "lstm": 4, "rnn": 1}[str(self._config.rnn_class)]

    @property
    def _n_mlp_hidden(self):
        return int(self._config.io_spec.targets[0].module.n_hidden_layers)

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
            if str(self._config.inputs_mode) == "static_mix":
                e[p + "input_module.weights"] = (1,)
            e[p + "input_module.heads.0.2.weight"] = (H, fs[i])
            e[p + "input_module.heads.0.2.bias"] = (H,)
            for k in range(self._config.n_rnn):       # nn.GRU / nn.LSTM / nn.RNN parameter names, layer by layer
                e[p + f"rnn.weight_ih_l{k}"] = (self._gates * H, H)
                e[p + f"rnn.weight_hh_l{k}"] = (self._gates * H, H)
                if self._config.rnn_bias:             # sample_rnn_v2.py:66: nn.GRU / nn.LSTM / nn.RNN(bias=rnn_bias)
                    e[p + f"rnn.bias_ih_l{k}"] = (self._gates * H,)
                    e[p + f"rnn.bias_hh_l{k}"] = (self._gates * H,)
            e[p + "up_sampler.fc.weight"] = (H * self._up(i), H)
            e[p + "up_sampler.fc.bias"] = (H * self._up(i),)
        if str(self._config.inputs_mode) == "static_mix":
            e[f"tiers.{len(fs) - 1}.input_module.weights"] = (1,)
        p = f"tiers.{len(fs) - 1}.input_module.heads.0.2.2.cv."
        e[p + "weight"] = (H, 1, fs[-1])
        e[p + "bias"] = (H,)
        p = "output_modules.0.estimator.0."
        if self._learns_temperature:                      # mlp.py:29, 54-57: without it, Q outputs and no buffer
            e[p + "min_temp"] = ()
        e[p + "fc.0.weight"] = (Hh, H)
        e[p + "fc.0.bias"] = (Hh,)
        # networks/mlp.py:47-50 builds the hidden layers by repeating a TUPLE that holds one nn.Linear: fc.2, fc.4, ... are
        # the same module (the reference's state_dict lists it under every index); the last Linear follows them
        nh = self._n_mlp_hidden
        for r in range(nh):
            e[p + f"fc.{2 + 2 * r}.weight"] = (Hh, Hh)
            e[p + f"fc.{2 + 2 * r}.bias"] = (Hh,)
        e[p + f"fc.{2 + 2 * nh}.weight"] = (Q + int(self._learns_temperature), Hh)
        e[p + f"fc.{2 + 2 * nh}.bias"] = (Q + int(self._learns_temperature),)
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
            if k.endswith("input_module.weights"):        # modules/io.py:304
                sd[k] = -torch.rand(shape)
                continue
            if ".rnn." in k:
                bound = 1.0 / math.sqrt(H)
            else:
                wshape = shape if k.endswith("weight") else shapes[k[:-4] + "weight"]
                bound = 1.0 / math.sqrt(max(1, int(torch.tensor(wshape[1:]).prod())))
            sd[k] = (torch.rand(shape) * 2 - 1) * bound
        p = "output_modules.0.estimator.0."
        for r in range(1, self._n_mlp_hidden):            # one shared Linear under several names
            sd[p + f"fc.{2 + 2 * r}.weight"] = sd[p + "fc.2.weight"]
            sd[p + f"fc.{2 + 2 * r}.bias"] = sd[p + "fc.2.bias"]
        return sd

    # ---- native handle --------------------------------------------------------------------------
    def _create_handle(self, max_batch):
        H, Hh, Q = self._dims()
        fs = self.frame_sizes
        n = len(fs)
        c = self._config
        dx = _capi.SampleRNNDescEx()
        d = dx.base
        d.n_tiers, d.hidden_dim, d.head_hidden, d.q_levels = n, H, Hh, Q
        fsa = (ctypes.c_int * n)(*fs)
        d.frame_sizes = fsa
        keep = [fsa]
        def arr(fmt):
            a = self._warray([fmt.format(i) for i in range(n - 1)])
            keep.append(a)
            return a
        def rnn_arr(name):
            a = self._warray([f"tiers.{i}.rnn.{name}_l{k}" for i in range(n - 1) for k in range(c.n_rnn)])
            keep.append(a)
            return a
        d.in_w, d.in_b = arr("tiers.{}.input_module.heads.0.2.weight"), arr("tiers.{}.input_module.heads.0.2.bias")
        dx.rnn_type = {"gru": 0, "lstm": 1, "rnn": 2}[str(c.rnn_class)]
        dx.n_rnn = int(c.n_rnn)
        dx.w_ih, dx.w_hh = rnn_arr("weight_ih"), rnn_arr("weight_hh")
        if c.rnn_bias:
            dx.b_ih, dx.b_hh = rnn_arr("bias_ih"), rnn_arr("bias_hh")
        else:                                         # adding 0.0f changes no value: the bias-free cell, bit for bit
            self._zero_bias = torch.zeros(self._gates * H, dtype=torch.float32)
            za = (ctypes.POINTER(ctypes.c_float) * ((n - 1) * c.n_rnn))(*([_capi.fptr(self._zero_bias)] * ((n - 1) * c.n_rnn)))
            keep.append(za)
            dx.b_ih = dx.b_hh = za
        d.up_w, d.up_b = arr("tiers.{}.up_sampler.fc.weight"), arr("tiers.{}.up_sampler.fc.bias")
        p = f"tiers.{n - 1}.input_module.heads.0.2.2.cv."
        d.conv_w, d.conv_b = self._w(p + "weight"), self._w(p + "bias")
        p = "output_modules.0.estimator.0."
        nh = self._n_mlp_hidden
        d.head_w1, d.head_b1 = self._w(p + "fc.0.weight"), self._w(p + "fc.0.bias")
        d.head_w2, d.head_b2, d.min_temperature = self._head_last(p + f"fc.{2 + 2 * nh}.weight", p + f"fc.{2 + 2 * nh}.bias")
        dx.head_hidden_layers = nh
        if nh > 0:
            for r in range(1, nh):
                if not (torch.equal(self._sd[p + f"fc.{2 + 2 * r}.weight"], self._sd[p + "fc.2.weight"])
                        and torch.equal(self._sd[p + f"fc.{2 + 2 * r}.bias"], self._sd[p + "fc.2.bias"])):
                    raise RuntimeError("the hidden layers of the reference MLP share one Linear (networks/mlp.py:47-50): "
                                       "fc.2, fc.4, ... must hold the same tensors")
            dx.head_wh, dx.head_bh = self._w(p + "fc.2.weight"), self._w(p + "fc.2.bias")
        dx.need_set_hidden = int(str(c.h0_init) != "zeros")
        dx.compute_mode = 1 if self._compute_dtype == torch.bfloat16 else 0      # MMK_COMPUTE_BF16_TC / MMK_COMPUTE_FP32
        h = ctypes.c_void_p()
        _capi.check(_capi.lib().mmk_samplernn_create_ex(ctypes.byref(dx), int(max_batch), ctypes.byref(h)))
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

    def _initial_state(self, B, generator=None, h0=None):
        """SampleRNNTier._init_h0 (sample_rnn_v2.py:113-119): zeros (None here: the kernel's reset does it), ones, or randn.
        Returns {(tier, layer, which): (B, H) cuda tensor}; which = 1 is the LSTM cell state, drawn like the hidden state.
        `h0` overrides the draw with explicit tensors under the same keys (the reference draws randn from torch's global
        generator at its first forward; pass what it drew to reproduce it)."""
        c = self._config
        if h0 is not None:
            return {k: torch.as_tensor(v, dtype=torch.float32).to(self.device).contiguous() for k, v in h0.items()}
        if str(c.h0_init) == "zeros":
            return None
        H = c.hidden_dim
        out = {}
        for i in range(len(self.frame_sizes) - 1):
            for which in ((0, 1) if str(c.rnn_class) == "lstm" else (0,)):
                if str(c.h0_init) == "ones":
                    block = torch.ones((c.n_rnn, B, H), device=self.device)
                else:
                    block = torch.randn((c.n_rnn, B, H), generator=generator,
                                        device=generator.device if generator is not None else "cpu").to(self.device)
                for k in range(c.n_rnn):
                    out[(i, k, which)] = block[k].contiguous()
        return out

    def _install_state(self, state, B):
        """Zero every state (a run with reset_hidden = 1 over empty ranges), then write the non-zero initial values."""
        h = self._get_handle(B)
        dummy = torch.zeros((B, self.rf), dtype=torch.int64, device=self.device)
        self._run(dummy, 0, (0, 0, 0), (self.rf, self.rf), True, True, None, None, 0, False, False, False)
        with torch.cuda.device(self.device):
            for (tier, layer, which), v in state.items():
                if tuple(v.shape) != (B, self._config.hidden_dim):
                    raise ValueError(f"initial state {(tier, layer, which)} must be ({B}, {self._config.hidden_dim})")
                _capi.check(_capi.lib().mmk_samplernn_set_hidden(h, int(tier), int(layer), int(which), v.data_ptr(), B,
                                                                 _capi.stream_ptr()))

    def _warm_range(self, P):
        """before_generate (sample_rnn_v2.py:229-234): logical t in [rf, P - P % rf), data shifted by P % rf."""
        offset = P % self.rf
        return (self.rf, P - offset, offset)

    # ---- whole-sequence fast path ---------------------------------------------------------------
    def generate(self, prompts, n_steps, temperature=None, noise=None, return_logits=False,
                 return_step_timestamps=False, generator=None, h0=None):
        """before_generate (hidden reset + warm-up over the prompt) followed by n_steps of generate_step as
        GenerateLoopV2.run drives them (loops/generate.py:193-219), in ONE kernel launch.  Arguments and returns as
        `WaveNet.generate`; `h0` (optional) = explicit initial states, see `_initial_state`."""
        seq = prepare_sequence(prompts, n_steps, self.device)
        B, total = seq.shape
        P = total - n_steps
        if P < self.rf:
            raise RuntimeError(f"prompt length {P} is shorter than the top frame size {self.rf}")
        T = as_temperature(temperature, B, self.device)
        U = prepare_noise(noise, T, B, n_steps, self.device, generator)
        state = self._initial_state(B, generator, h0)
        if state is not None:
            self._install_state(state, B)
        logits, _, ts = self._run(seq, 0, self._warm_range(P), (P, P + n_steps), state is None, False, T, U, P,
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

    def teacher_forced(self, sequence, prompt_len, temperature=None, noise=None, h0=None):
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
        state = self._initial_state(B, None, h0)
        if state is not None:
            self._install_state(state, B)
        logits, dec, _ = self._run(seq, 0, self._warm_range(P), (P, total), state is None, True, T, U, P, True, True, False)
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
        state = self._initial_state(p.shape[0])
        if state is not None:
            self._install_state(state, p.shape[0])
        self._run(p, 0, self._warm_range(P), (P, P), state is None, True, None, None, 0, False, False, False)
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
