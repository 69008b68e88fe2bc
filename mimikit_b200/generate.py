"""`GenerateLoopV2` — the reference's generation driver (mimikit/loops/generate.py:85-252) over the B200 networks.

Same constructor, `Config` fields and generator contract: `GenerateLoopV2(config, network, n_steps, dataloader,
logger=None).run()` yields, per dataloader batch `[prompt_idx, *prompt_tensors]`, a tuple with one
(B, prior_t + n_steps) tensor per network input — expanded to a float waveform by `feature.inv` when
`yield_inversed_outputs` (the default), else the mu-law indices.

A network that offers the whole-sequence fast path (`network.generate`, one persistent kernel launch) is driven
through it; any other `ARM` is driven step by step exactly as the reference does (:207-219), so third-party ARMs
keep working.  Device placement: the reference moves the network to `default_device()` (utils.py:27-35); here that
is the current CUDA device and there is no CPU alternative.
"""
import dataclasses as dtc
from typing import Any, Callable, Dict, Optional, Tuple

import numpy as np
import torch

from . import _capi

__all__ = ["GenerateLoopV2", "fill", "prepare_prompt"]


def prepare_prompt(device, prompt, n_blanks, at_least_nd=2):
    """generate.py:26-39 for a single array / tensor or a (nested) tuple of them."""
    if isinstance(prompt, (tuple, list)):
        return type(prompt)(prepare_prompt(device, p, n_blanks, at_least_nd) for p in prompt)
    if isinstance(prompt, np.ndarray):
        prompt = torch.from_numpy(prompt)
    while prompt.dim() < at_least_nd:
        prompt = prompt.unsqueeze(0)
    prompt = prompt.to(device)
    if n_blanks > 0:
        blank = torch.zeros((prompt.size(0), n_blanks, *prompt.shape[2:]), dtype=prompt.dtype, device=prompt.device)
        prompt = torch.cat((prompt, blank), dim=1)
    return prompt


def fill(x, prior_t, n_steps):
    """generate.py:50-73: `x` followed by blanks (zeros of x's dtype) or per-example constants."""
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(x)
    parts = [x] if x is not None else []
    dt, dev = (x.dtype, x.device) if x is not None else (torch.float32, "cpu")
    B, D = (x.size(0), tuple(x.shape[2:])) if x is not None else (1, (1,))
    for kind, n in (prior_t, n_steps):
        if isinstance(kind, torch.Tensor):
            assert kind.shape == (B,)
            parts.append(kind.expand(B, n, 1))
        elif kind == "blank":
            parts.append(torch.zeros(B, n, *D, dtype=dt, device=dev))
    return torch.cat(parts, dim=1)


class GenerateLoopV2:
    @dtc.dataclass
    class Config:
        """generate.py:86-99 — field for field."""
        output_duration_sec: float = 1.
        prompts_length_sec: float = 1.
        prompts_position_sec: Tuple[Optional[float], ...] = (None,)
        parameters: Optional[Dict[str, Any]] = None
        batch_size: int = 1
        downsampling: int = 1
        output_name_template: Optional[str] = None
        display_waveform: bool = True
        write_waveform: bool = False
        yield_inversed_outputs: bool = True
        callback: Optional[Callable[[Tuple[torch.Tensor, ...]], None]] = None

    @classmethod
    def get_n_steps(cls, config, network):
        """generate.py:101-111 for a Sample-unit io_spec (mu-law): int(sr * output_duration_sec)."""
        return int(network.config.io_spec.sr * config.output_duration_sec)

    def __init__(self, config, network, n_steps: int, dataloader, logger=None):
        self.config = config
        self.network = network
        self.n_steps = n_steps
        self.dataloader = dataloader
        self.logger = logger
        self.device = None
        self.template_vars = {}
        self._initial_device = None
        self._was_training = False
        self._grad_was_enabled = True

    def setup(self):
        _capi.require_cuda()
        net = self.network
        self._initial_device = net.device
        self._was_training = bool(getattr(net, "training", False))
        net.eval()
        self.device = torch.device("cuda", torch.cuda.current_device())
        net.to(self.device)
        self._grad_was_enabled = torch.is_grad_enabled()
        torch.set_grad_enabled(False)

    def teardown(self):
        self.network.to(self._initial_device)
        if self._was_training:
            self.network.train()
        torch.set_grad_enabled(self._grad_was_enabled)

    def run(self):
        self.setup()
        try:
            for batch in self.dataloader:
                prompt_idx, batch = batch[0], batch[1:]
                batch = tuple((torch.from_numpy(x) if isinstance(x, np.ndarray) else x).to(self.device)
                              for x in batch)
                params = self.config.parameters or {}
                params = {k: v for k, v in params.items() if k in self.network.generate_params}
                self.network.before_generate(batch, prompt_idx)
                if hasattr(self.network, "generate") and len(batch) == 1:
                    final_outputs = (self.network.generate(batch[0], self.n_steps, **params),)
                else:
                    final_outputs = self._run_stepwise(batch, params)
                self.network.after_generate(final_outputs, prompt_idx)
                final_outputs = self.process_outputs(tuple(final_outputs), prompt_idx, **self.template_vars)
                yield final_outputs
                if self.config.callback is not None:
                    self.config.callback(final_outputs)
        finally:
            self.teardown()

    def _run_stepwise(self, batch, params):
        """generate.py:195-219."""
        rf, prior_t, n_steps = self.network.rf, batch[0].size(1), self.n_steps
        tensors = tuple(fill(x, ("data", prior_t), ("blank", n_steps)) for x in batch)
        until = 0
        for t in range(prior_t, prior_t + n_steps):
            if t < until:
                continue
            inputs = tuple(tensor[:, t - rf:t] for tensor in tensors)
            outputs = self.network.generate_step(inputs, t=t, **params)
            if not isinstance(outputs, tuple):
                outputs = outputs,
            for tensor, out in zip(tensors, outputs):
                if out is not None:
                    n_out = min(out.size(1), tensor.size(1) - t)
                    tensor[:, t:t + n_out] = out[:, :n_out]
                    until = t + n_out
        return tuple(tensors)

    def process_outputs(self, final_outputs, prompt_idx, **template_vars):
        """generate.py:231-252."""
        cfg = self.config
        if (self.logger is None or (not cfg.write_waveform and not cfg.display_waveform)) \
                and not cfg.yield_inversed_outputs:
            return final_outputs
        features = self.network.config.io_spec.targets
        outputs = tuple(feature.inv(out) for feature, out in zip(features, final_outputs))
        if self.logger is not None:
            for output in outputs:
                for example, idx in zip(output, prompt_idx):
                    idx = idx.item() if hasattr(idx, "item") else idx
                    if cfg.write_waveform:
                        self.logger.write(example, prompt_idx=idx, **template_vars)
                    if cfg.display_waveform:
                        self.logger.display(example, prompt_idx=idx, **template_vars)
        return outputs if cfg.yield_inversed_outputs else final_outputs
