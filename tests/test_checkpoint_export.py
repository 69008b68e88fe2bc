"""CPU tests of checkpoint ingestion (mimikit_b200/checkpoint.py): networks built by the LIVE reference are exported with
the duck-typed `export_network`, reloaded as this package's networks, and must carry the same weights, geometry and IO
wiring.  (Skipped where /root/reference is absent, e.g. on the GPU box; generation parity from reference state dicts is
what the golden GPU tests check.)"""
import numpy as np
import pytest
import torch

from oracle import ref_loader

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="the reference tree is not present")


def _roundtrip(ref_net, tmp_path):
    from mimikit_b200 import export_network, load_exported, save_exported
    path = save_exported(ref_net, str(tmp_path / "model.b200.pt"))
    ours = load_exported(path)
    sd_ref = {k: v for k, v in ref_net.state_dict().items()}
    sd = ours.state_dict()
    return ours, sd_ref, sd, export_network(ref_net)


def test_wavenet_export_roundtrip(tmp_path):
    ref_net = ref_loader.make_wavenet(blocks=(3, 2), dims=64, residuals_dim=64, skips_dim=48, mlp_dim=32, seed=3)
    ours, sd_ref, sd, d = _roundtrip(ref_net, tmp_path)
    assert type(ours).__name__ == "WaveNet" and ours.rf == ref_net.rf and list(sd) == [k for k in sd_ref]
    for k in sd_ref:
        assert torch.equal(sd[k], sd_ref[k].float().cpu()), k
    assert d["io"] == dict(sr=16000, q_levels=256, compression=1.0, input_module_type="embedding", mlp_dim=32,
                           n_mlp_layers=0, min_temperature=1e-4)
    assert ours.config.blocks == (3, 2) and ours.config.skips_dim == 48 and ours.config.act_g == "Sigmoid"
    # an export of OUR network reloads too (same format)
    from mimikit_b200 import export_network, load_exported
    again = load_exported(export_network(ours))
    assert all(torch.equal(again.state_dict()[k], sd[k]) for k in sd)


def test_wavenet_config_surface_export_roundtrip(tmp_path):
    """A reference WaveNet off the W-30 form — affine residuals, Mish / Softplus gate units, kernel size 3, an MLP head WITHOUT
    the learned temperature (no min_temp buffer, Q-row last Linear) — exported and reloaded key for key."""
    ref_net = ref_loader.make_wavenet(blocks=(2, 2), dims=32, residuals_dim=32, skips_dim=32, mlp_dim=32, seed=4,
                                      with_affine_residuals=True, act_f="Mish", act_g="Softplus", kernel_sizes=(3,),
                                      min_temperature=None)
    ours, sd_ref, sd, d = _roundtrip(ref_net, tmp_path)
    assert ours.rf == ref_net.rf and list(sd) == [k for k in sd_ref]
    assert "layers.0.aff_res.params.weight" in sd and "output_modules.0.estimator.0.min_temp" not in sd
    assert tuple(sd["output_modules.0.estimator.0.fc.2.weight"].shape) == (256, 32)
    for k in sd_ref:
        assert torch.equal(sd[k], sd_ref[k].float().cpu()), k
    c = ours.config
    assert c.with_affine_residuals and str(c.act_f) == "Mish" and str(c.act_g) == "Softplus" and d["io"]["min_temperature"] is None


def test_samplernn_export_roundtrip_and_errors(tmp_path):
    ref_net = ref_loader.make_samplernn(frame_sizes=(8, 2, 1), hidden_dim=64, mlp_dim=32, seed=5)
    ours, sd_ref, sd, d = _roundtrip(ref_net, tmp_path)
    assert type(ours).__name__ == "SampleRNN" and tuple(ours.frame_sizes) == (8, 2, 1)
    assert set(sd) == set(sd_ref)
    for k in sd_ref:
        assert torch.equal(sd[k], sd_ref[k].float().cpu()), k
    assert d["io"]["input_module_type"] == "framed_linear" and d["config"]["rnn_class"] == "gru"
    from mimikit_b200 import load_exported
    with pytest.raises(ValueError):
        load_exported({"format": "something else"})
    bad = dict(d, config=dict(d["config"], no_such_field=1))
    with pytest.raises(ValueError):
        load_exported(bad)
    none = dict(d, config=dict(d["config"], rnn_class="none"))
    with pytest.raises(NotImplementedError):
        load_exported(none)                      # unsupported configurations fail loudly, never silently
    lstm = dict(d, config=dict(d["config"], rnn_class="lstm"))
    with pytest.raises(RuntimeError):
        load_exported(lstm)                      # ... and a GRU state dict does not fit an LSTM network


def test_samplernn_reference_defaults_export_roundtrip(tmp_path):
    """The reference's DEFAULT tier (rnn_class 'lstm', sample_rnn_v2.py:127), two stacked layers and two hidden MLP layers —
    whose shared Linear the reference's state_dict lists under fc.2 AND fc.4 (mlp.py:47-50) — survive the hand-over."""
    ref_net = ref_loader.make_samplernn(frame_sizes=(8, 2, 1), hidden_dim=32, mlp_dim=32, seed=5, rnn_class="lstm", n_rnn=2,
                                        n_mlp_layers=2, h0_init="ones")
    ours, sd_ref, sd, d = _roundtrip(ref_net, tmp_path)
    assert set(sd) == set(sd_ref) and d["config"]["rnn_class"] == "lstm" and d["config"]["n_rnn"] == 2
    assert d["io"]["n_mlp_layers"] == 2 and d["config"]["h0_init"] == "ones"
    for k in sd_ref:
        assert torch.equal(sd[k], sd_ref[k].float().cpu()), k
    p = "output_modules.0.estimator.0."
    assert torch.equal(sd[p + "fc.2.weight"], sd[p + "fc.4.weight"]) and tuple(sd[p + "fc.6.weight"].shape) == (257, 32)
    assert tuple(sd["tiers.0.rnn.weight_ih_l1"].shape) == (128, 32)


def test_samplernn_config_surface_export_roundtrip(tmp_path):
    """rnn_bias=False (no rnn.bias_* parameters), inputs_mode 'static_mix' (an `input_module.weights` parameter per tier) and
    a head without the learned temperature: the exported network reloads with the reference's keys, in the reference's order."""
    ref_net = ref_loader.make_samplernn(frame_sizes=(4, 2, 1), hidden_dim=32, mlp_dim=32, seed=6, rnn_class="gru", rnn_bias=False,
                                        inputs_mode="static_mix", min_temperature=None)
    ours, sd_ref, sd, d = _roundtrip(ref_net, tmp_path)
    assert list(sd) == [k for k in sd_ref]
    assert "tiers.0.rnn.bias_ih_l0" not in sd and tuple(sd["tiers.2.input_module.weights"].shape) == (1,)
    assert "output_modules.0.estimator.0.min_temp" not in sd
    for k in sd_ref:
        assert torch.equal(sd[k], sd_ref[k].float().cpu()), k
    assert d["config"]["inputs_mode"] == "static_mix" and d["config"]["rnn_bias"] is False
