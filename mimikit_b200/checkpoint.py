"""Checkpoint ingestion for the generation path.

The reference stores checkpoints as HDF5 banks written through h5mapper with an OmegaConf yaml config
(mimikit/checkpoint.py:51-93, 144-152).  Neither library is a dependency of this package, so ingestion is a two-step
hand-over that touches nothing but plain torch objects:

    # where mimikit is installed (scripts/export_mimikit_checkpoint.py does exactly this):
    net = mmk.Checkpoint.from_path(".../epoch=12.ckpt").network       # checkpoint.py:144-152
    torch.save(mimikit_b200.checkpoint.export_network(net), "model.b200.pt")
    # on the B200 box:
    net = mimikit_b200.load_exported("model.b200.pt").to("cuda")      # WaveNet / SampleRNN drop-in, weights loaded

`export_network` is duck-typed on the reference's network objects (it reads `net.config`, `net.config.io_spec` and
`net.state_dict()`; nothing is imported from mimikit) and works on this package's networks as well.  weight_norm'ed
SampleRNN tiers are folded when the state dict is loaded (sample_rnn_v2.py:67-81).
"""
import dataclasses as dtc
from collections import OrderedDict

import torch

from .io_spec import IOSpec
from .sample_rnn import SampleRNN
from .wavenet import WaveNet

__all__ = ["export_network", "save_exported", "load_exported", "FORMAT"]

FORMAT = "mimikit_b200.exported_network.v1"
_INPUT_TYPES = {"EmbeddingIO": "embedding", "FramedLinearIO": "framed_linear"}


def _plain(v):
    if isinstance(v, (list, tuple)):
        return [_plain(x) for x in v]
    if isinstance(v, (bool, int, float, str)) or v is None:
        return v
    return str(getattr(v, "act", v))      # ActivationConfig -> its name; anything else -> its repr


def _io_config(io_spec):
    """The MuLawIOConfig (io_spec.py:210-218) that `IOSpec.mulaw_io` would need to rebuild this spec."""
    if len(io_spec.inputs) != 1 or len(io_spec.targets) != 1:
        raise ValueError("only single-input / single-target mu-law IO specs can be exported")
    i0, t0 = io_spec.inputs[0], io_spec.targets[0]
    tr = i0.transform
    if type(tr).__name__ != "MuLawCompress":
        raise ValueError(f"input transform {type(tr).__name__} is not MuLawCompress")
    module_type = getattr(i0, "module_type", None) or _INPUT_TYPES.get(type(i0.module).__name__)
    if module_type not in ("embedding", "framed_linear"):
        raise ValueError(f"input module {type(i0.module).__name__} is not EmbeddingIO / FramedLinearIO")
    head = t0.module
    sr = getattr(i0, "sr", None)
    if sr is None:                                    # reference: the sample rate lives in the extractor's FileToSignal
        try:
            sr = i0.extractor.functional.functionals[0].sr
        except Exception:
            sr = 16000
    return dict(sr=int(sr), q_levels=int(tr.q_levels), compression=float(tr.compression), input_module_type=module_type,
                mlp_dim=int(head.hidden_dim), n_mlp_layers=int(head.n_hidden_layers),
                min_temperature=None if head.min_temperature is None else float(head.min_temperature))


def export_network(net):
    """-> a dict of plain python / torch objects: class name, config fields, mu-law IO config, state_dict (CPU)."""
    cls = type(net).__name__
    if cls not in ("WaveNet", "SampleRNN"):
        raise ValueError(f"cannot export a {cls}: the B200 path hosts WaveNet and SampleRNN")
    cfg = net.config
    fields = {f.name: _plain(getattr(cfg, f.name)) for f in dtc.fields(cfg) if f.name not in ("io_spec", "type")}
    sd = OrderedDict((k, v.detach().to("cpu").clone()) for k, v in net.state_dict().items())
    return {"format": FORMAT, "class": cls, "config": fields, "io": _io_config(cfg.io_spec), "state_dict": sd}


def save_exported(net, path):
    torch.save(export_network(net), path)
    return path


def load_exported(path_or_dict, device=None):
    """Rebuilds the network (WaveNet / SampleRNN of this package) from `export_network`'s dict or a file holding it and
    loads the weights.  Unsupported configurations raise at `from_config`, as everywhere in this package."""
    d = path_or_dict if isinstance(path_or_dict, dict) else torch.load(path_or_dict, map_location="cpu", weights_only=True)
    if d.get("format") != FORMAT:
        raise ValueError(f"not a {FORMAT} file")
    io_spec = IOSpec.mulaw_io(IOSpec.MuLawIOConfig(**d["io"]))
    cls = {"WaveNet": WaveNet, "SampleRNN": SampleRNN}[d["class"]]
    known = {f.name for f in dtc.fields(cls.Config)}
    unknown = sorted(set(d["config"]) - known)
    if unknown:
        raise ValueError(f"config fields {unknown} are not part of {d['class']}.Config")
    kw = {k: (tuple(v) if isinstance(v, list) else v) for k, v in d["config"].items()}
    net = cls.from_config(cls.Config(io_spec=io_spec, **kw))
    net.load_state_dict(d["state_dict"])
    return net.to(device) if device is not None else net
