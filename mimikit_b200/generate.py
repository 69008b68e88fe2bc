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

__all__ = ["GenerateLoopV2", "extend_with_blanks", "prepare_prompt"]


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


def extend_with_blanks(x, n_steps):
    """The working buffer of one network input: the prompt followed by `n_steps` zero positions of the same dtype
    and trailing shape (what the reference's loop builds before stepping, loops/generate.py:197-200)."""
    x = torch.from_numpy(x) if isinstance(x, np.ndarray) else x
    out = x.new_zeros((x.shape[0], x.shape[1] + int(n_steps)) + tuple(x.shape[2:]))
    out[:, :x.shape[1]] = x
    return out


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
        """Foreign ARMs (no whole-sequence fast path): one `generate_step` per write position, the contract of
        loops/generate.py:207-219 — the step sees the `rf` positions before the cursor of every buffer, may return a
        single tensor or a tuple, `None` entries leave their buffer untouched, and an output of several positions
        moves the cursor past all of them."""
        rf, start = self.network.rf, batch[0].size(1)
        stop = start + self.n_steps
        buffers = tuple(extend_with_blanks(x, self.n_steps) for x in batch)
        cursor = start
        while cursor < stop:
            produced = self.network.generate_step(tuple(buf[:, cursor - rf:cursor] for buf in buffers), t=cursor, **params)
            produced = produced if isinstance(produced, tuple) else (produced,)
            advance = 1
            for buf, new in zip(buffers, produced):
                if new is None:
                    continue
                width = min(new.size(1), stop - cursor)
                buf[:, cursor:cursor + width] = new[:, :width]
                advance = max(1, width)     # the reference resumes after the last buffer written
            cursor += advance
        return buffers

    def _invert(self, sequences):
        """`feature.inv` of every target (mu-law expand for the networks here; loops/generate.py:245, io_spec.py:77-79)."""
        return tuple(spec.inv(seq) for spec, seq in zip(self.network.config.io_spec.targets, sequences))

    def process_outputs(self, final_outputs, prompt_idx, **template_vars):
        """Same observable behaviour as loops/generate.py:231-252: the inverse transform runs only when somebody
        consumes it (a logger that writes/displays, or the caller via `yield_inversed_outputs`); the logger sees every
        example of every output with its prompt index; the yielded value is the waveform or the raw sequences."""
        cfg = self.config
        logging = self.logger is not None and (cfg.write_waveform or cfg.display_waveform)
        if not logging and not cfg.yield_inversed_outputs:
            return final_outputs
        waveforms = self._invert(final_outputs)
        if self.logger is not None:
            ids = [i.item() if hasattr(i, "item") else i for i in prompt_idx]
            for wave in waveforms:
                for example, idx in zip(wave, ids):
                    if cfg.write_waveform:
                        self.logger.write(example, prompt_idx=idx, **template_vars)
                    if cfg.display_waveform:
                        self.logger.display(example, prompt_idx=idx, **template_vars)
        return waveforms if cfg.yield_inversed_outputs else final_outputs
