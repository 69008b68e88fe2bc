#!/usr/bin/env python
"""Run WHERE MIMIKIT IS INSTALLED: turns a mimikit checkpoint into the plain-torch file mimikit_b200.load_exported reads.

    python export_mimikit_checkpoint.py path/to/<id>/epoch=<n>.ckpt model.b200.pt

Uses only mimikit's public API (Checkpoint.from_path(...).network, mimikit/checkpoint.py:118-152); the export itself is
mimikit_b200.checkpoint.export_network if this package is importable there, else the same 30 lines inlined below."""
import dataclasses as dtc
import sys
from collections import OrderedDict

import torch


def export_network(net):
    try:
        from mimikit_b200.checkpoint import export_network as impl
        return impl(net)
    except ImportError:
        pass
    plain = lambda v: [plain(x) for x in v] if isinstance(v, (list, tuple)) else \
        (v if isinstance(v, (bool, int, float, str)) or v is None else str(getattr(v, "act", v)))
    cfg = net.config
    i0, t0 = cfg.io_spec.inputs[0], cfg.io_spec.targets[0]
    try:
        sr = i0.extractor.functional.functionals[0].sr
    except Exception:
        sr = 16000
    io = dict(sr=int(sr), q_levels=int(i0.transform.q_levels), compression=float(i0.transform.compression),
              input_module_type={"EmbeddingIO": "embedding", "FramedLinearIO": "framed_linear"}[type(i0.module).__name__],
              mlp_dim=int(t0.module.hidden_dim), n_mlp_layers=int(t0.module.n_hidden_layers),
              min_temperature=t0.module.min_temperature)
    fields = {f.name: plain(getattr(cfg, f.name)) for f in dtc.fields(cfg) if f.name not in ("io_spec", "type")}
    sd = OrderedDict((k, v.detach().cpu().clone()) for k, v in net.state_dict().items())
    return {"format": "mimikit_b200.exported_network.v1", "class": type(net).__name__, "config": fields, "io": io,
            "state_dict": sd}


if __name__ == "__main__":
    import mimikit as mmk
    src, dst = sys.argv[1], sys.argv[2]
    torch.save(export_network(mmk.Checkpoint.from_path(src).network), dst)
    print("wrote", dst)
