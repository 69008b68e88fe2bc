"""Chunked long-form generation — the function behind the reference's script mimikit/loops/generate_chunks.py:39-56.

The script generates a long track as a chain of GenerateLoopV2 runs: chunk i is prompted with the last
`prompts_length_sec` of chunk i-1, the per-prompt temperature vector drifts between chunks (:42-43), and only the newly
generated part of every run is kept (:52-55).  Two modes here:

* `carry_state=False` — exactly that: every chunk is a fresh `network.generate` on the tail of the previous one (WaveNet
  prefills its rings over the last rf samples again; SampleRNN resets its hidden states and warms up over the re-prompt).
* `carry_state=True`  — the B200 way: the first chunk prompts, every later chunk is `network.generate_more`, which relaunches
  the persistent kernel on the rings / GRU states still sitting in the native handle.  For WaveNet the samples are the
  same bit for bit (a re-prompt of >= rf samples rebuilds exactly the state that was kept); for SampleRNN it is one
  uninterrupted generation instead of a chain of cold restarts.
"""
from typing import Callable, Optional

import torch

from .arm import Temperature

__all__ = ["generate_chunks"]


def generate_chunks(network, prompts: torch.Tensor, n_chunks: int, chunk_steps: int, prompt_length: Optional[int] = None,
                    temperature: Temperature = None, temperature_update: Optional[Callable] = None,
                    noise: Optional[torch.Tensor] = None, carry_state: bool = False, generator=None) -> torch.Tensor:
    """prompts (B, P) int64 -> (B, P + n_chunks * chunk_steps) int64 on the network's device.

    prompt_length       samples of the previous output every chunk after the first is prompted with (default: P);
                        ignored with carry_state=True
    temperature         None | float | (1,) | (B,) for the first chunk; `temperature_update(i, temperature)` gives chunk i's
                        (the reference adds clipped noise to a per-prompt vector between chunks)
    noise               optional (B, n_chunks * chunk_steps) uniform noise, consumed chunk by chunk"""
    if prompts.dim() == 1:
        prompts = prompts.unsqueeze(0)
    B, P = prompts.shape
    L = P if prompt_length is None else int(prompt_length)
    if not carry_state and (L < network.rf or L > P):
        raise ValueError(f"prompt_length must lie in [{network.rf}, {P}], got {L}")
    out = torch.zeros((B, P + n_chunks * chunk_steps), dtype=torch.int64, device=network.device)
    out[:, :P] = prompts.to(network.device)
    T = temperature
    for i in range(n_chunks):
        if i > 0 and temperature_update is not None:
            T = temperature_update(i, T)
        lo = P + i * chunk_steps
        U = None if noise is None else noise[:, i * chunk_steps:(i + 1) * chunk_steps]
        if carry_state and i > 0:
            new = network.generate_more(chunk_steps, temperature=T, noise=U, generator=generator)
        else:
            src = out[:, :P] if i == 0 else out[:, lo - L:lo]
            new = network.generate(src, chunk_steps, temperature=T, noise=U, generator=generator)[:, -chunk_steps:]
        out[:, lo:lo + chunk_steps] = new
    return out
