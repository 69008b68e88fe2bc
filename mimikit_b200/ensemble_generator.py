"""EnsembleGenerator — mimikit/models/ensemble_generator.py:54-163: generate from a prompt by CHAINING networks.  A stream of
events {generator, seconds, temperature} is consumed one at a time; every event resamples the tail of what exists so far
to its network's rate, turns it into the network's input feature, runs one GenerateLoopV2 batch and splices the inverse-
transformed, resampled-back continuation onto the output.  The loop body of every event is one persistent-kernel launch
here.  The reference's NearestNextNeighbor generator (models/nnn.py) is not part of the B200 path."""
import dataclasses as dtc
from pprint import pprint
from typing import Any, Iterator, Optional

import torch

from .features import Resample
from .generate import GenerateLoopV2

__all__ = ["EnsembleGenerator", "Event"]


@dtc.dataclass
class Event:
    """ensemble_generator.py:54-58.  `generator` is a network (ARM) or any object with a `.network` (a checkpoint)."""
    generator: Any
    seconds: float
    temperature: Optional[float] = None


class EnsembleGenerator:
    """Same constructor, `run`, `generate_step`, `run_event` and `next_event` as the reference (:61-163)."""

    def __init__(self, prompt: torch.Tensor, max_seconds: float = 10., base_sr: int = 22050, stream: Iterator = (),
                 print_events: bool = False, device=None):
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.prompt = prompt.to(self.device)
        self.max_seconds = max_seconds
        self.base_sr = base_sr
        self.stream = iter(stream)
        self.print_events = print_events

    @property
    def _n_samples(self):
        return int(self.max_seconds * self.base_sr)

    def run(self):
        """:80-93 — a waveform buffer of max_seconds at base_sr, filled event by event; every event sees the last
        `prompt_length` samples."""
        window = cursor = self.prompt.size(-1)
        output = torch.zeros(self.prompt.size(0), self._n_samples, dtype=self.prompt.dtype, device=self.device)
        output[:, :cursor] = self.prompt
        while cursor < self._n_samples:
            piece = self.generate_step(cursor, output[:, cursor - window:cursor])
            if piece is None:
                break
            width = min(piece.size(1), self._n_samples - cursor)
            output[:, cursor:cursor + width] = piece[:, :width]
            cursor += piece.size(1)
        return output

    def generate_step(self, t, inputs):
        """:95-110 — the next event if it still fits before max_seconds, else silence up to the end."""
        if t >= self._n_samples:
            return None
        event, net, n_steps, params = self.next_event()
        if hasattr(net, "to"):
            net = net.to(self.device)
        if (t / self.base_sr + event.seconds) < self.max_seconds:
            if self.print_events:
                pprint({**dtc.asdict(event), "start": t / self.base_sr})
            return self.run_event(inputs, net, n_steps, params)
        return torch.zeros(inputs.size(0), self._n_samples - t, device=self.device)

    def run_event(self, inputs: torch.Tensor, net, n_steps: int, params: dict):
        """:112-144 — base_sr waveform -> network rate -> input features -> one generation batch -> waveform at base_sr
        (the generated part only)."""
        io = net.config.io_spec
        at_net_rate = Resample(self.base_sr, io.sr)(inputs)
        prompt = tuple(spec.transform(at_net_rate) for spec in io.inputs)
        n_prompt = prompt[0].shape[1]          # sample-level features: one feature step per waveform sample
        cfg = GenerateLoopV2.Config(parameters=params or None, display_waveform=False, write_waveform=False,
                                    yield_inversed_outputs=True)
        loop = GenerateLoopV2(cfg, network=net, n_steps=n_steps, dataloader=[[torch.ones(1), *prompt]], logger=None)
        for outputs in loop.run():
            return Resample(io.sr, self.base_sr)(outputs[0][:, n_prompt:])

    def next_event(self):
        """:146-163."""
        event = Event(**next(self.stream))
        net = getattr(event.generator, "network", event.generator)
        if not hasattr(net, "generate_step"):
            raise TypeError(f"event generator type '{type(event.generator)}' not supported")
        n_steps = int(event.seconds * net.config.io_spec.sr)      # GenerateLoopV2.get_n_steps for a sample-level target
        params = dict(temperature=event.temperature) if event.temperature is not None else {}
        return event, net, n_steps, params
