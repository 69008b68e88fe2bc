"""Host-side base of the B200 networks: the reference's `ARM` plug-in interface (mimikit/networks/arm.py:28-80)
over a native handle.

A network here is NOT an nn.Module: it owns a flat fp32 state dict under the reference's parameter names (so the
reference's checkpoints load with `load_state_dict`), and a native handle (C ABI, include/mmk_b200.h) that holds
the repacked device copy of the weights plus the generation state.  The methods the reference's loop calls on a
network (`eval/train/to/device/training`, `rf`, `generate_params`, `before_generate`, `generate_step`,
`after_generate`) exist with the same meaning; `generate(...)` is the whole-sequence fast path (one persistent
kernel launch) that `GenerateLoopV2` here uses instead of one Python call per sample.
"""
import ctypes
from collections import OrderedDict
from typing import Optional, Sequence, Union

import numpy as np
import torch

from . import _capi

Temperature = Union[None, float, Sequence[float], torch.Tensor, np.ndarray]


def as_temperature(temperature: Temperature, batch: int, device) -> Optional[torch.Tensor]:
    """CategoricalSampler's `as_tensor` (mimikit/modules/targets.py:27-34): None | float | 1-sequence | (B,)."""
    if temperature is None:
        return None
    t = torch.as_tensor(temperature, dtype=torch.float32).reshape(-1)
    if t.numel() not in (1, batch):
        raise ValueError(f"temperature must have 1 or {batch} entries, got {t.numel()}")
    return t.to(device).contiguous()


class NativeARM:
    """Common state-dict / handle plumbing.  Subclasses define `_expected_shapes()`, `_create_handle(max_batch)`,
    `_destroy_handle(h)` and the generation entry points."""

    def __init__(self):
        self._sd = OrderedDict()
        self._absent = {}        # zero stand-ins for the biases of modules the reference built with bias=False
        self._handle = None
        self._handle_batch = 0
        self._handle_device = None
        self._device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() \
            else torch.device("cpu")
        self.training = False

    # ---- nn.Module-like surface the reference's loop touches (loops/generate.py:170-182) ----
    @property
    def device(self):
        return self._device

    def eval(self):
        self.training = False
        return self

    def train(self, mode=True):
        if mode:
            raise NotImplementedError("mimikit_b200 networks are generation-only (training is out of scope)")
        self.training = False
        return self

    def to(self, device):
        device = torch.device(device)
        if device.type == "cuda" and device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        if device != self._device:
            self._release()
            self._device = device
        return self

    def parameters(self):
        return iter(self._sd.values())

    @property
    def q_levels(self) -> int:
        """Size of the output alphabet (TargetSpec.out_dim, io_spec.py:136-149) — what the multi-GPU gather sizes its
        wire dtype with."""
        return int(self.config.io_spec.targets[0].out_dim)

    def state_dict(self):
        return OrderedDict((k, v.clone()) for k, v in self._sd.items())

    def load_state_dict(self, state_dict, strict=True):
        """Same key names and shapes as the reference network's state_dict.  weight_norm'ed checkpoints
        (`*_g` / `*_v`, sample_rnn_v2.py:67-81) are folded on load."""
        sd = _fold_weight_norm(state_dict)
        expected = self._expected_shapes()
        missing = [k for k in expected if k not in sd]
        unexpected = [k for k in sd if k not in expected]
        if strict and (missing or unexpected):
            raise RuntimeError(f"Error(s) in loading state_dict: missing keys {missing}, unexpected keys {unexpected}")
        new = OrderedDict()
        for k, shape in expected.items():
            if k not in sd:
                new[k] = self._sd[k]
                continue
            v = torch.as_tensor(np.asarray(sd[k]) if not isinstance(sd[k], torch.Tensor) else sd[k])
            v = v.detach().to("cpu", torch.float32).contiguous()
            if tuple(v.shape) != tuple(shape):
                raise RuntimeError(f"size mismatch for {k}: got {tuple(v.shape)}, expected {tuple(shape)}")
            new[k] = v
        self._sd = new
        self._release()
        return self

    # ---- native handle ----
    def _release(self):
        if self._handle is not None:
            self._destroy_handle(self._handle)
            self._handle = None
            self._handle_batch = 0

    def __del__(self):
        try:
            self._release()
        except Exception:
            pass

    def _get_handle(self, batch):
        _capi.require_cuda()
        if self._device.type != "cuda":
            raise _capi.MmkError("network is on the CPU: call .to('cuda') first (there is no CPU fallback)")
        if self._handle is None or batch > self._handle_batch or self._handle_device != self._device:
            self._release()
            with torch.cuda.device(self._device):
                self._handle = self._create_handle(batch)
            self._handle_batch = batch
            self._handle_device = self._device
        return self._handle

    def _w(self, key):
        t = self._sd.get(key)
        if t is None:                                  # a bias the reference did not build (bias=False): zeros stand in for it
            t = self._absent[key]
        return _capi.fptr(t)

    @property
    def _learns_temperature(self):
        """networks/mlp.py:29, 54-62: MLP(min_temperature=None) has Q outputs, no `min_temp` buffer and no division."""
        return self._config.io_spec.targets[0].module.min_temperature is not None

    def _head_last(self, wkey, bkey):
        """(weight pointer, bias pointer, min_temperature) of the head's last Linear as the kernels expect it: Q + 1 rows, the
        last one the temperature logit.  A head WITHOUT the learned temperature is hosted through its weights alone: a zero
        row with bias 40 is appended and min_temperature is 0, so the kernels divide by max(sigmoid(40), 0) = 1 / (1 + exp(-40)),
        which rounds to exactly 1.0f — logits / 1.0f are the logits, bit for bit."""
        if self._learns_temperature:
            return self._w(wkey), self._w(bkey), float(self._sd["output_modules.0.estimator.0.min_temp"])
        w, b = self._sd[wkey], self._sd[bkey]
        self._head_pad = (torch.cat([w, torch.zeros_like(w[:1])], 0).contiguous(),
                          torch.cat([b, torch.full_like(b[:1], 40.0)], 0).contiguous())
        return _capi.fptr(self._head_pad[0]), _capi.fptr(self._head_pad[1]), 0.0

    def _warray(self, keys):
        """const float* const* over per-layer tensors; a None key gives a NULL entry."""
        arr = (ctypes.POINTER(ctypes.c_float) * len(keys))()
        for i, k in enumerate(keys):
            arr[i] = self._w(k) if k is not None else ctypes.POINTER(ctypes.c_float)()
        return arr


def _fold_weight_norm(sd):
    if not any(k.endswith("_g") for k in sd):
        return sd
    out = OrderedDict((k, v) for k, v in sd.items() if not (k.endswith("_g") or k.endswith("_v")))
    for k in sd:
        if k.endswith("_v"):
            base = k[:-2]
            v = torch.as_tensor(sd[k]).double()
            g = torch.as_tensor(sd[base + "_g"]).double()
            norm = v.reshape(v.shape[0], -1).norm(dim=1).reshape((-1,) + (1,) * (v.dim() - 1))
            out[base] = (g * v / norm).float()
    return out


def prepare_sequence(prompts, n_steps, device):
    """loops/generate.py:197-200 (`fill`): the prompt followed by n_steps zeros, int64, on the device."""
    if isinstance(prompts, np.ndarray):
        prompts = torch.from_numpy(prompts)
    if prompts.dim() == 1:
        prompts = prompts.unsqueeze(0)
    if prompts.dtype != torch.int64:
        prompts = prompts.to(torch.int64)
    B, P = prompts.shape
    seq = torch.zeros((B, P + n_steps), dtype=torch.int64, device=device)
    seq[:, :P].copy_(prompts, non_blocking=True)
    return seq


def prepare_noise(noise, temperature, B, n_steps, device, generator=None):
    if temperature is None:
        return None
    if noise is None:
        return torch.rand((B, n_steps), device=device, dtype=torch.float32, generator=generator)
    noise = torch.as_tensor(noise, dtype=torch.float32).to(device).contiguous()
    if tuple(noise.shape) != (B, n_steps):
        raise ValueError(f"noise must be ({B}, {n_steps}) uniform [0,1) fp32, got {tuple(noise.shape)}")
    return noise
