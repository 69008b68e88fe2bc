"""The part of the reference's IO-spec surface the generation path reads (mimikit/io_spec.py:27-253).

Only the mu-law wiring exists here: `IOSpec.mulaw_io(IOSpec.MuLawIOConfig(...))` gives one mu-law input (embedding
or framed-linear module) and one mu-law target with an MLP head + categorical sampler — the objects the networks'
`from_config` and `GenerateLoopV2.process_outputs` look at (`sr`, `unit`, `inputs[i].transform`, `targets[i].inv`).
Spectral IO (`magspec_io`), loss terms and dataset binding are training-side and out of scope.
"""
import dataclasses as dtc
from typing import Optional, Tuple

from .features import MuLawCompress, Sample

__all__ = ["IOSpec", "InputSpec", "TargetSpec", "MLPIOConfig"]


@dtc.dataclass
class MLPIOConfig:
    """The fields of modules/io.py:201-220 (MLPIO) that shape the output head."""
    hidden_dim: int = 128
    n_hidden_layers: int = 0
    min_temperature: Optional[float] = 1e-4


@dtc.dataclass
class _FeatureSpec:
    """io_spec.py:27-80."""
    extractor_name: str
    transform: MuLawCompress
    sample_rate: int

    @property
    def sr(self):
        return self.sample_rate

    @property
    def unit(self):
        return Sample(self.sample_rate)

    @property
    def elem_type(self):
        return self.transform.elem_type

    @property
    def inv(self):
        return self.transform.inv


@dtc.dataclass
class InputSpec(_FeatureSpec):
    """io_spec.py:83-92; module_type is 'embedding' (EmbeddingIO) or 'framed_linear' (FramedLinearIO)."""
    module_type: str = "framed_linear"

    @property
    def class_size(self):
        return self.elem_type.size


@dtc.dataclass
class TargetSpec(_FeatureSpec):
    """io_spec.py:132-149; objective is always 'categorical_dist' on this path."""
    module: MLPIOConfig = dtc.field(default_factory=MLPIOConfig)
    objective_type: str = "categorical_dist"

    @property
    def out_dim(self):
        return self.elem_type.size


@dtc.dataclass
class IOSpec:
    inputs: Tuple[InputSpec, ...]
    targets: Tuple[TargetSpec, ...]

    @dtc.dataclass
    class MuLawIOConfig:
        """io_spec.py:210-218."""
        sr: int = 16000
        q_levels: int = 256
        compression: float = 1.
        input_module_type: str = 'framed_linear'
        mlp_dim: int = 128
        n_mlp_layers: int = 0
        min_temperature: float = 1e-4

    @staticmethod
    def mulaw_io(config: "IOSpec.MuLawIOConfig", extractor=None):
        """io_spec.py:220-253."""
        c = config
        if c.input_module_type not in ("framed_linear", "embedding"):
            raise ValueError(f"Unimplemented input_module_type: '{c.input_module_type}'")
        mu_law = MuLawCompress(c.q_levels, c.compression)
        name = getattr(extractor, "name", "signal")
        return IOSpec(
            inputs=(InputSpec(name, mu_law, c.sr, module_type=c.input_module_type),),
            targets=(TargetSpec(name, mu_law, c.sr,
                                module=MLPIOConfig(c.mlp_dim, c.n_mlp_layers, c.min_temperature)),))

    @property
    def sr(self):
        srs = {i.sr for i in [*self.inputs, *self.targets]}
        if len(srs) > 1:
            raise RuntimeError(f"Expected to find a single sample_rate but found several: '{srs}'")
        return srs.pop()

    @property
    def unit(self):
        units = {i.unit for i in [*self.inputs, *self.targets]}
        if len(units) > 1:
            raise RuntimeError(f"Expected to find a single time unit but found several: '{units}'")
        return units.pop()

    @property
    def hop_length(self):
        return None
