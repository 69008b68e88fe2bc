"""CPU tests: the oracle (oracle/restate.py + oracle/c) against the golden vectors generated from the live
reference (oracle/make_golden.py) and against the reference's own structural KATs."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, golden_state_dict, load_golden
from oracle import restate


def test_mulaw_compress_bit_exact_vs_reference():
    d = load_golden("mulaw")
    for q, C in [(256, 1.), (256, .5), (64, 2.), (1024, 1.)]:
        got = restate.mulaw_compress(d["x"], q, C)
        assert got.dtype == np.int64
        assert np.array_equal(got, d[f"idx_q{q}_c{C}"]), (q, C)
        exp = restate.mulaw_expand(np.arange(q), q, C)
        np.testing.assert_allclose(exp, d[f"expand_q{q}_c{C}"], rtol=0, atol=1e-6)
    got = restate.mulaw_compress(d["x_int"].astype(np.float32), 256, 1.)
    assert np.array_equal(got, d["idx_int_q256_c1.0"])


def test_mulaw_roundtrip_is_idempotent():
    # compress(expand(i)) == i for every class: size-independent property used at full size on the GPU
    for q, C in [(256, 1.), (256, .5), (512, 1.)]:
        idx = np.arange(q)
        assert np.array_equal(restate.mulaw_compress(restate.mulaw_expand(idx, q, C), q, C), idx)


def test_magspec_vs_reference_within_1e4():
    d = load_golden("magspec")
    for tag in "abc":
        n_fft, hop = (int(v) for v in d[f"cfg_{tag}"])
        for center in (True, False):
            ref = d[f"mag_{tag}_center{int(center)}"]
            got = restate.magspec(d[f"x_{tag}"], n_fft, hop, center)
            assert got.shape == ref.shape
            # tolerance: 1e-4 relative to the clip's peak magnitude (north star: "STFT/mel within 1e-4")
            assert np.abs(got - ref).max() <= 1e-4 * max(1.0, ref.max())


def test_stft_frame_count_kats():
    d = load_golden("magspec")
    for L, center, n in d["frame_kat"]:
        assert restate.stft_n_frames(int(L), 2048, 512, bool(center)) == int(n), (L, center)
    # reference tests/test_fft_alignment.py: centered -> L//hop + 1 frames
    assert restate.stft_n_frames(220500, 2048, 512, True) == 431
    assert restate.stft_n_frames(220500, 2048, 512, False) == 427


def test_mel_filterbank_vs_torchaudio_slaney():
    d = np.load(os.path.join(GOLDEN, "mel_fb_torchaudio.npz"))
    fb = np.zeros(tuple(d["shape"]), dtype=np.float32)
    fb[d["rows"].astype(int), d["cols"].astype(int)] = d["vals"]
    mine = restate.mel_filterbank(2048, 128, 0., None, False)
    assert mine.shape == (128, 1025)
    assert np.abs(mine - fb).max() < 1e-6


def test_wavenet_structure_kats():
    with open(os.path.join(GOLDEN, "wavenet_structure_kat.json")) as f:
        kats = json.load(f)
    for k in kats:
        ks, ds = restate.wavenet_kernels_and_dilations(k["kernel_sizes"], k["blocks"])
        pairs = list(zip(ks, ds))
        assert [d for _, d in pairs] == k["dilations"], k
        assert [kk for kk, _ in pairs] == k["kernels"], k
        assert sum((kk - 1) * d for kk, d in pairs) + 1 == k["rf"]
    # reference tests/test_wavenet.py:251-275: rf == 8 for these layouts
    for blocks, ks in [((3,), (2,)), ((3,), (2, 2, 2)), ((), (2, 2, 2))]:
        assert restate.wavenet_rf(ks, blocks) == 8
    assert restate.wavenet_rf((2,), (8, 8, 7, 7)) == 765
    assert restate.wavenet_rf((2,), (8, 8, 8, 8)) == 1021


@pytest.mark.parametrize("name", ["wavenet_default_small", "wavenet_res_skip_small", "wavenet_res_skip_mid"])
def test_wavenet_oracle_vs_reference(name):
    d = load_golden(name)
    orc = restate.WaveNetOracle(golden_state_dict(d), tuple(int(b) for b in d["meta/blocks"]))
    n = d["noise"].shape[1]
    for tag, T in [("argmax", None), ("t1", 1.0), ("tvec", d["tvec"])]:
        seq, lg = orc.generate(d["prompts"], n, T, d["noise"])
        assert np.array_equal(seq, d["seq_" + tag]), tag
        np.testing.assert_allclose(lg, d["logits_" + tag], rtol=1e-3, atol=1e-5)
    assert np.array_equal(d["seq_argmax"], d["seq_argmax_real_loop"])
    # window recompute (the reference's algorithm) == cached step: teacher-forced logits on the reference's own sequence
    tf = orc.logits_teacher_forced(d["seq_argmax"])
    P = d["prompts"].shape[1]
    np.testing.assert_allclose(tf[:, P - orc.rf:P - orc.rf + n], d["logits_argmax"], rtol=1e-3, atol=1e-5)
    # prompt shorter than the receptive field is an error (reference: RuntimeError, tests/test_wavenet.py:271-275)
    with pytest.raises(ValueError):
        orc.generate(d["prompts"][:, :orc.rf - 1], 2)


@pytest.mark.parametrize("name", ["samplernn_821_small", "samplernn_821_small_ragged", "samplernn_1642_small",
                                  "samplernn_41_small"])
def test_samplernn_oracle_vs_reference(name):
    d = load_golden(name)
    orc = restate.SampleRNNOracle(golden_state_dict(d), tuple(int(b) for b in d["meta/frame_sizes"]))
    n = d["noise"].shape[1]
    for tag, T in [("argmax", None), ("t1", 1.0), ("tvec", d["tvec"])]:
        seq, lg = orc.generate(d["prompts"], n, T, d["noise"])
        assert np.array_equal(seq, d["seq_" + tag]), tag
        np.testing.assert_allclose(lg, d["logits_" + tag], rtol=1e-3, atol=1e-5)


def test_sampling_contract_properties():
    rng = np.random.default_rng(0)
    lg = rng.standard_normal((64, 256)).astype(np.float32) * 3
    # u -> 0 picks the first class with mass, u -> 1 the last; monotone in u
    lo = restate.sample_inverse_cdf(lg, [1.0], np.zeros(64, np.float32))
    hi = restate.sample_inverse_cdf(lg, [1.0], np.full(64, np.float32(1) - np.float32(2 ** -24)))
    assert (lo == 0).all() and (hi >= 250).all()
    us = np.sort(rng.random((64, 50)).astype(np.float32), axis=1)
    prev = np.zeros(64, dtype=np.int64)
    for j in range(50):
        cur = restate.sample_inverse_cdf(lg, [0.9], us[:, j])
        assert (cur >= prev).all()
        prev = cur
    # low temperature converges to argmax
    cold = restate.sample_inverse_cdf(lg, [1e-3], np.full(64, .5, np.float32))
    assert np.array_equal(cold, restate.argmax_first(lg))
    # empirical frequencies follow softmax
    l1 = np.tile(lg[:1], (20000, 1))
    s = restate.sample_inverse_cdf(l1, [1.0], rng.random(20000).astype(np.float32))
    p = np.exp(lg[0] - lg[0].max()); p /= p.sum()
    freq = np.bincount(s, minlength=256) / 20000
    assert np.abs(freq - p).max() < 0.02


def test_normalize_and_compose_oracle_vs_reference():
    """Normalize(p=inf) and Compose(Normalize(), MuLawCompress()) of the live reference (tests/golden/normalize.npz)."""
    d = load_golden("normalize")
    got = restate.normalize_inf(d["x"])
    assert got.dtype == np.float32 and np.array_equal(got.view(np.int32), d["norm"].view(np.int32))   # bit-exact
    assert np.array_equal(restate.normalize_inf(d["x1"]).view(np.int32), d["norm1"].view(np.int32))
    assert np.abs(got[6]).max() == 1.0 and got[6, 17] == -1.0 and not got[5].any()
    for q, C in [(256, 1.), (64, 2.)]:
        assert np.array_equal(restate.mulaw_compress(got, q, C), d[f"compose_q{q}_c{C}"])


def test_remove_dc_oracle_vs_reference():
    """RemoveDC.np_func of the live reference (scipy fp64 IIR): the C restatement is bit-exact, alone and after Normalize."""
    d = load_golden("normalize")
    assert np.array_equal(restate.remove_dc(d["dc_x"]).view(np.int32), d["dc_y"].view(np.int32))
    assert np.array_equal(restate.remove_dc(restate.normalize_inf(d["dc_x"])).view(np.int32), d["dc_norm_y"].view(np.int32))
    y = restate.remove_dc(d["dc_x"])
    assert not y[3].any() and abs(float(y[4, -1])) < 1e-6 and abs(float(y[0, 2000:].mean())) < 1e-3   # silence, constant, DC gone


def test_bf16_faithful_wavenet_oracle_tracks_the_fp32_oracle():
    """WaveNetBf16Oracle (the tensor-core kernel's arithmetic on the CPU) stays within bf16 noise of the fp32 oracle on a
    golden network, and its rounding helper is round-to-nearest-even."""
    assert restate.bf16_round(np.float32(1.00390625)) == np.float32(1.0)            # tie -> even mantissa
    assert restate.bf16_round(np.float32(1.01171875)) == np.float32(1.015625)       # tie -> even mantissa (up)
    assert restate.bf16_round(np.float32(-3.1415927)) == np.float32(-3.140625)
    d = load_golden("wavenet_res_skip_small")
    sd = golden_state_dict(d)
    blocks = tuple(int(b) for b in d["meta/blocks"])
    P, seq = d["prompts"].shape[1], d["seq_argmax"]
    _, lg32 = restate.WaveNetOracle(sd, blocks).generate(d["prompts"], seq.shape[1] - P, None, None, forced=seq)
    lg16 = restate.WaveNetBf16Oracle(sd, blocks).logits_for(seq, P)
    err = np.abs(lg16 - lg32).max() / np.abs(lg32).max()
    assert 1e-5 < err < 2e-2, err


@pytest.mark.parametrize("name", ["wavenet_default_small", "wavenet_res_skip_small"])
def test_wavenet_torch_port_vs_reference(name):
    """oracle/torch_port.py is what bench.py times as `cpu_baseline` / `--impl reference`: the window-recompute port must
    produce the live reference's sequences (argmax and noise-sampled) bit for bit."""
    import torch
    from oracle import torch_port
    d = load_golden(name)
    sd = {k: torch.from_numpy(v) for k, v in golden_state_dict(d).items()}
    blocks = tuple(int(b) for b in d["meta/blocks"])
    _, dil = restate.wavenet_kernels_and_dilations((2,), blocks)
    port = torch_port.WaveNetPort(sd, dil)
    prompts, noise = torch.from_numpy(d["prompts"]), torch.from_numpy(d["noise"])
    n = noise.shape[1]
    assert np.array_equal(port.generate(prompts, n).numpy(), d["seq_argmax"])
    assert np.array_equal(port.generate(prompts, n, 1.0, noise).numpy(), d["seq_t1"])
    P = prompts.shape[1]
    lg = port.window_logits(torch.from_numpy(d["seq_argmax"]))[:, P - port.rf:P - port.rf + n].numpy()
    np.testing.assert_allclose(lg, d["logits_argmax"], rtol=1e-3, atol=1e-5)


@pytest.mark.parametrize("name", ["samplernn_821_small", "samplernn_821_small_ragged", "samplernn_41_small"])
def test_samplernn_torch_port_vs_reference(name):
    import torch
    from oracle import torch_port
    d = load_golden(name)
    sd = {k: torch.from_numpy(v) for k, v in golden_state_dict(d).items()}
    port = torch_port.SampleRNNPort(sd, tuple(int(b) for b in d["meta/frame_sizes"]))
    prompts, noise = torch.from_numpy(d["prompts"]), torch.from_numpy(d["noise"])
    n = noise.shape[1]
    assert np.array_equal(port.generate(prompts, n).numpy(), d["seq_argmax"])
    assert np.array_equal(port.generate(prompts, n, 1.0, noise).numpy(), d["seq_t1"])


def test_product_wavenet_structure_matches_the_reference_kats():
    """The PRODUCT's WaveNet.get_kernels_and_dilation / rf (mimikit_b200/wavenet.py) against the KATs taken from the live
    reference (wavenet_v2.py:295-342) — the same vectors the oracle's function is held to above."""
    from mimikit_b200 import WaveNet
    with open(os.path.join(GOLDEN, "wavenet_structure_kat.json")) as f:
        kats = json.load(f)
    for k in kats:
        ks, ds = WaveNet.get_kernels_and_dilation(tuple(k["kernel_sizes"]), tuple(k["blocks"]))
        assert [int(v) for v in ds] == k["dilations"], k
        assert [int(v) for v in ks] == k["kernels"], k
        assert sum((int(a) - 1) * int(b) for a, b in zip(ks, ds)) + 1 == k["rf"]


VARIANTS = ["samplernn_lstm_default", "samplernn_lstm_2layers_ones", "samplernn_gru_3layers_randn_mlp2", "samplernn_rnn_tanh_mlp1",
            "samplernn_no_temperature", "samplernn_lstm_nobias", "samplernn_gru_nobias_2layers",
            "samplernn_static_mix", "samplernn_mean"]


def variant_setup(d):
    """(oracle kwargs, h0 dict) of a samplernn_* variant golden (oracle/make_golden.py: gen_samplernn_variant)."""
    m = {k[5:]: v for k, v in d.items() if k.startswith("meta/")}
    kw = dict(rnn_class=str(m["rnn_class"]), n_rnn=int(m["n_rnn"]), n_mlp_hidden=int(m["n_mlp_layers"]))
    B, H = d["prompts"].shape[0], int(m["hidden_dim"])
    h0 = None
    if str(m["h0_init"]) == "ones":
        n_ft = len(m["frame_sizes"]) - 1
        h0 = {(i, k, w): np.ones((B, H), np.float32) for i in range(n_ft) for k in range(kw["n_rnn"])
              for w in ((0, 1) if kw["rnn_class"] == "lstm" else (0,))}
    elif str(m["h0_init"]) == "randn":
        h0 = {tuple(int(v) for v in k[3:].split("_")): d[k] for k in d if k.startswith("h0/")}
    return m, kw, h0


@pytest.mark.parametrize("name", VARIANTS)
def test_samplernn_variant_oracle_vs_reference(name):
    """LSTM (the reference default), stacked layers, h0_init ones / randn, tanh RNN, hidden MLP layers (sample_rnn_v2.py:
    62-66, 101-119; mlp.py:47-50): the oracle against sequences and logits of the live reference."""
    d = load_golden(name)
    m, kw, h0 = variant_setup(d)
    orc = restate.SampleRNNOracle(golden_state_dict(d), tuple(int(b) for b in m["frame_sizes"]), **kw)
    n = d["noise"].shape[1]
    for tag, T in [("argmax", None), ("t1", 1.0)]:
        seq, lg = orc.generate(d["prompts"], n, T, d["noise"], h0=h0)
        assert np.array_equal(seq, d["seq_" + tag]), (name, tag)
        np.testing.assert_allclose(lg, d["logits_" + tag], rtol=1e-3, atol=1e-5)


WN_VARIANTS = ["wavenet_pad_side1", "wavenet_layerwise_inputs", "wavenet_layerwise_noskip_mlp2", "wavenet_kernel3", "wavenet_reversed", "wavenet_nongated", "wavenet_groups4",
               "wavenet_affine_res", "wavenet_affine_plain", "wavenet_act_mish_softplus", "wavenet_act_sin_cos",
               "wavenet_act_relu_identity", "wavenet_act_abs_tanh", "wavenet_act_sigmoid_none", "wavenet_no_temperature", "wavenet_noblocks", "wavenet_nobias_affine", "wavenet_dropped_res",
               "wavenet_reversed_noskip", "wavenet_noblocks_noskip"]


def wavenet_variant_kwargs(d):
    m = {k[5:]: v for k, v in d.items() if k.startswith("meta/")}
    ks = tuple(int(k) for k in m["kernel_sizes"]) if "kernel_sizes" in m else (int(m.get("kernel_size", 2)),)
    return m, dict(kernel_sizes=ks, layerwise_inputs=bool(int(m.get("layerwise_inputs", 0))),
                   n_mlp_hidden=int(m.get("n_mlp_layers", 0)), reverse_layer_order=bool(int(m.get("reverse_layer_order", 0))),
                   act_f=str(m.get("act_f", "Tanh")), act_g=str(m.get("act_g", "Sigmoid")))


@pytest.mark.parametrize("name", WN_VARIANTS)
def test_wavenet_variant_oracle_vs_reference(name):
    """pad_side=1 (same value at the evaluated position), layerwise_inputs, hidden MLP layers and kernel_size 3
    (wavenet_v2.py:137-138, 273, 283-284, 295-327; mlp.py:47-50): the oracle against the live reference, incl. its real loop."""
    d = load_golden(name)
    m, kw = wavenet_variant_kwargs(d)
    orc = restate.WaveNetOracle(golden_state_dict(d), tuple(int(b) for b in m["blocks"]), **kw)
    n = d["noise"].shape[1]
    for tag, T in [("argmax", None), ("t1", 1.0), ("tvec", d["tvec"])]:
        seq, lg = orc.generate(d["prompts"], n, T, d["noise"])
        assert np.array_equal(seq, d["seq_" + tag]), (name, tag)
        np.testing.assert_allclose(lg, d["logits_" + tag], rtol=1e-3, atol=1e-5)
    assert np.array_equal(d["seq_argmax"], d["seq_argmax_real_loop"])
    tf = orc.logits_teacher_forced(d["seq_argmax"])
    P = d["prompts"].shape[1]
    np.testing.assert_allclose(tf[:, P - orc.rf:P - orc.rf + n], d["logits_argmax"], rtol=1e-3, atol=1e-5)


# ---------------------------------------------------------------------------------------------------------------
# host logic of the drop-in networks (no GPU): every golden state dict of the live reference loads into the product
# network built from the same configuration — same keys, same order, same shapes — and unsupported ones raise
# ---------------------------------------------------------------------------------------------------------------
def product_wavenet_config(d):
    from mimikit_b200 import IOSpec, WaveNet
    m, kw = wavenet_variant_kwargs(d)
    head = dict(min_temperature=None) if int(m.get("no_temperature", 0)) else {}
    return WaveNet.Config(
        io_spec=IOSpec.mulaw_io(IOSpec.MuLawIOConfig(input_module_type="embedding", mlp_dim=int(m["mlp_dim"]),
                                                     n_mlp_layers=kw["n_mlp_hidden"], **head)),
        blocks=tuple(int(b) for b in m["blocks"]), dims_dilated=(int(m["dims"]),),
        residuals_dim=int(m["residuals_dim"]) if "residuals_dim" in m else None,
        skips_dim=int(m["skips_dim"]) if "skips_dim" in m else None, kernel_sizes=kw["kernel_sizes"],
        layerwise_inputs=kw["layerwise_inputs"], pad_side=int(m.get("pad_side", 0)), reverse_layer_order=kw["reverse_layer_order"],
        groups=int(m.get("groups", 1)), with_affine_residuals=bool(int(m.get("affine", 0))), bias=not int(m.get("nobias", 0)),
        act_f=kw["act_f"], act_g=None if int(m.get("nongated", 0)) else kw["act_g"])


@pytest.mark.parametrize("name", WN_VARIANTS + ["wavenet_default_small", "wavenet_res_skip_small", "wavenet_res_skip_mid"])
def test_product_wavenet_takes_the_reference_state_dict(name):
    from mimikit_b200 import WaveNet
    d = load_golden(name)
    sd = golden_state_dict(d)
    net = WaveNet.from_config(product_wavenet_config(d))
    assert list(net.state_dict()) == list(sd), name                       # the reference's keys in the reference's order
    net.load_state_dict(sd)
    for k, v in net.state_dict().items():
        assert np.array_equal(v.numpy(), np.asarray(sd[k], dtype=np.float32)), (name, k)
    orc = restate.WaveNetOracle(sd, tuple(int(b) for b in d["meta/blocks"]), **wavenet_variant_kwargs(d)[1])
    assert net.rf == orc.rf and net.dilations == orc.dilations and net.kernels == orc.kernels


@pytest.mark.parametrize("name", VARIANTS + ["samplernn_821_small", "samplernn_1642_small", "samplernn_41_small"])
def test_product_samplernn_takes_the_reference_state_dict(name):
    from mimikit_b200 import IOSpec, SampleRNN
    d = load_golden(name)
    sd = golden_state_dict(d)
    m = {k[5:]: v for k, v in d.items() if k.startswith("meta/")}
    head = dict(min_temperature=None) if int(m.get("no_temperature", 0)) else {}
    cfg = SampleRNN.Config(
        io_spec=IOSpec.mulaw_io(IOSpec.MuLawIOConfig(mlp_dim=int(m["mlp_dim"]), n_mlp_layers=int(m.get("n_mlp_layers", 0)), **head)),
        frame_sizes=tuple(int(f) for f in m["frame_sizes"]), hidden_dim=int(m["hidden_dim"]), rnn_class=str(m.get("rnn_class", "gru")),
        n_rnn=int(m.get("n_rnn", 1)), h0_init=str(m.get("h0_init", "zeros")), rnn_bias=bool(int(m.get("rnn_bias", 1))),
        inputs_mode=str(m.get("inputs_mode", "sum")))
    net = SampleRNN.from_config(cfg)
    assert list(net.state_dict()) == list(sd), name
    net.load_state_dict(sd)
    for k, v in net.state_dict().items():
        assert np.array_equal(v.numpy(), np.asarray(sd[k], dtype=np.float32)), (name, k)


def test_product_config_surface_rejections():
    """What the kernels do not host raises NotImplementedError at from_config — never a silent change of the computation."""
    from mimikit_b200 import IOSpec, SampleRNN, WaveNet
    io = IOSpec.mulaw_io(IOSpec.MuLawIOConfig(input_module_type="embedding"))
    for bad in (dict(pad_side=-1), dict(stride=2), dict(kernel_sizes=(5,)), dict(dims_1x1=(16,)), dict(act_f="GLU"),
                dict(act_g="PhaseA"), dict(act_f="Softmax")):
        with pytest.raises(NotImplementedError):
            WaveNet.from_config(WaveNet.Config(io_spec=io, **bad))
    with pytest.raises(NotImplementedError):      # the reference's own default input type raises in the reference too
        WaveNet.from_config(WaveNet.Config(io_spec=IOSpec.mulaw_io(IOSpec.MuLawIOConfig())))
    fio = IOSpec.mulaw_io(IOSpec.MuLawIOConfig())
    for bad in (dict(rnn_class="none"), dict(n_rnn=5), dict(inputs_mode="prod"), dict(h0_init="uniform"), dict(frame_sizes=(4,)),
                dict(frame_sizes=(8, 3, 1))):
        with pytest.raises(NotImplementedError):
            SampleRNN.from_config(SampleRNN.Config(io_spec=fio, **bad))
    for ok in (dict(blocks=(), kernel_sizes=(2, 3), residuals_dim=128), dict(reverse_layer_order=True, residuals_dim=128),
               dict(residuals_dim=64), dict(bias=False), dict(with_affine_residuals=True), dict(act_f="Sin", act_g="Cos")):
        WaveNet.from_config(WaveNet.Config(io_spec=io, **ok))
