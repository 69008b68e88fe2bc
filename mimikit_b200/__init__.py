"""mimikit_b200 — B200-native (sm_100a) drop-in for ktonal/mimikit's batched autoregressive generation path
(WaveNet, SampleRNN; mu-law q-level audio) and the mu-law / STFT / mel feature preprocessing that feeds it.

Python/PyTorch host code (this package) calls hand-written CUDA through the C ABI in include/mmk_b200.h.
No Triton, no multi-backend dispatch, no CPU fallback.
"""
from .features import *  # noqa: F401,F403
from . import features  # noqa: F401
from .io_spec import IOSpec  # noqa: F401
from .wavenet import WaveNet  # noqa: F401
from .sample_rnn import SampleRNN  # noqa: F401
from .generate import GenerateLoopV2  # noqa: F401
from .checkpoint import export_network, load_exported, save_exported  # noqa: F401
from .chunks import generate_chunks  # noqa: F401
from .ensemble_generator import EnsembleGenerator, Event  # noqa: F401

__version__ = "0.1.0"
