"""TEST INFRASTRUCTURE — generates tests/golden/*.npz by running the UNMODIFIED reference (via oracle/ref_loader.py)
in this container.  The reference cannot travel to the GPU box, the fixtures can.  Run from the repo root:

    python -m oracle.make_golden

Every fixture stores the reference state_dict (reference key names), the seeded inputs and the reference's outputs.
Sampled sequences use the noise-driven sampler shim of ref_loader.NoiseSampler (the reference's torch.multinomial
cannot take external noise); argmax sequences additionally come from the real GenerateLoopV2.run.
"""
import os
import sys

import numpy as np
import torch

from . import ref_loader

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def sd_arrays(sd):
    return {"sd/" + k: v.detach().cpu().numpy() for k, v in sd.items()}


def gen_samplernn_variant(name, prompts, n_steps, h0_seed=None, **kw):
    """A SampleRNN of the live reference outside the GRU / one layer / zero state / plain head form: LSTM (the reference
    default), stacked layers, h0_init ones / randn, hidden MLP layers.  For randn the reference draws its initial states
    from torch's global generator at the first forward of every tier (sample_rnn_v2.py:101-119): the generator is seeded
    right before each run and the same draws, in the same order, are stored as `h0/<tier>_<layer>_<which>`."""
    kw = dict(kw)
    no_temp = kw.pop("no_temperature", False)
    if no_temp:
        kw["min_temperature"] = None
    net = ref_loader.make_samplernn(**kw)
    B = prompts.shape[0]
    noise = torch.rand(B, n_steps, generator=torch.Generator().manual_seed(4321))
    meta = dict(frame_sizes=kw["frame_sizes"], hidden_dim=kw["hidden_dim"], mlp_dim=kw["mlp_dim"], no_temperature=int(no_temp),
                rnn_class=kw.get("rnn_class", "gru"), n_rnn=kw.get("n_rnn", 1), h0_init=kw.get("h0_init", "zeros"),
                n_mlp_layers=kw.get("n_mlp_layers", 0), rnn_bias=int(kw.get("rnn_bias", True)), inputs_mode=kw.get("inputs_mode", "sum"))
    out = dict(sd_arrays(net.state_dict()), prompts=prompts.numpy(), noise=noise.numpy(),
               **{"meta/" + k: np.asarray(v) for k, v in meta.items()})
    H, n_rnn, lstm = kw["hidden_dim"], meta["n_rnn"], meta["rnn_class"] == "lstm"

    def run(temperature):
        for tier in net.tiers:
            tier.hidden = None                      # a fresh draw per run, as after_generate / reset_hidden leave it
        if h0_seed is not None:
            torch.manual_seed(h0_seed)
        return ref_loader.run_generate_loop(net, prompts, n_steps, temperature, noise)
    if meta["h0_init"] == "randn":
        torch.manual_seed(h0_seed)
        for i in range(len(kw["frame_sizes"]) - 1):           # every tier fires at the first warm-up step, top tier first
            for which in ((0, 1) if lstm else (0,)):
                block = torch.randn(n_rnn, B, H)
                for k in range(n_rnn):
                    out[f"h0/{i}_{k}_{which}"] = block[k].numpy()
    seq, lg = run(None)
    out["seq_argmax"], out["logits_argmax"] = seq.numpy(), lg.numpy()
    seq, lg = run(1.0)
    out["seq_t1"], out["logits_t1"] = seq.numpy(), lg.numpy()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, {k: v.shape for k, v in out.items() if not k.startswith("sd/")})


def gen_network(name, net, prompts, n_steps, meta):
    B = prompts.shape[0]
    noise = torch.rand(B, n_steps, generator=torch.Generator().manual_seed(4321))
    tvec = torch.linspace(.85, .999, B)
    out = dict(sd_arrays(net.state_dict()), prompts=prompts.numpy(), noise=noise.numpy(), tvec=tvec.numpy(),
               **{"meta/" + k: np.asarray(v) for k, v in meta.items()})
    seq, lg = ref_loader.run_generate_loop(net, prompts, n_steps, None, noise)
    out["seq_argmax"], out["logits_argmax"] = seq.numpy(), lg.numpy()
    real = ref_loader.run_real_generate_loop(net, prompts, n_steps)
    assert torch.equal(real, seq), "restated driver != GenerateLoopV2.run"
    out["seq_argmax_real_loop"] = real.numpy()
    seq, lg = ref_loader.run_generate_loop(net, prompts, n_steps, 1.0, noise)
    out["seq_t1"], out["logits_t1"] = seq.numpy(), lg.numpy()
    seq, lg = ref_loader.run_generate_loop(net, prompts, n_steps, tvec, noise)
    out["seq_tvec"], out["logits_tvec"] = seq.numpy(), lg.numpy()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, {k: v.shape for k, v in out.items() if not k.startswith("sd/")})


def restate_like_signal(g):
    """(5, 6007) fp32: sines with a DC offset + noise, one silent row, one constant row."""
    t = torch.arange(6007, dtype=torch.float64) / 16000
    x = torch.stack([0.4 * torch.sin(2 * np.pi * f * t) + dc for f, dc in [(220., .3), (659., -.2), (50., .05)]])
    x = x + 0.01 * torch.randn(3, 6007, generator=g, dtype=torch.float64)
    x = torch.cat([x, torch.zeros(1, 6007, dtype=torch.float64), torch.full((1, 6007), 0.25, dtype=torch.float64)])
    return x.float().numpy()


def gen_normalize(ref):
    """Normalize (functionals.py:236-253) and Compose(Normalize(), MuLawCompress()) (:196-213) of the live reference."""
    g = torch.Generator().manual_seed(99)
    x = (torch.randn(7, 4099, generator=g) * torch.tensor([1e-3, .1, .5, 1., 3., 1e-20, 1.])[:, None]).float()
    x[5] = 0.                                 # an all-zero clip: x / eps
    x[6, 17] = -7.5                           # the peak is a negative sample
    x1 = torch.rand(5, generator=g) * 2 - 1   # 1-D input
    F = ref.functionals
    d = dict(x=x.numpy(), x1=x1.numpy(), norm=F.Normalize().torch_func(x).numpy(), norm1=F.Normalize().torch_func(x1).numpy())
    for q, C in [(256, 1.), (64, 2.)]:
        d[f"compose_q{q}_c{C}"] = F.Compose(F.Normalize(), F.MuLawCompress(q, C))(x).numpy()
    # RemoveDC.np_func (functionals.py:216-233; the path used at extraction — torch_func passes lfilter's arguments in
    # the wrong order and cannot run): scipy fp64 IIR, cast back to fp32
    xs = restate_like_signal(g)
    d["dc_x"], d["dc_y"] = xs, F.RemoveDC().np_func(xs)
    d["dc_norm_y"] = F.RemoveDC().np_func(F.Normalize().torch_func(torch.from_numpy(xs)).numpy())
    np.savez_compressed(os.path.join(OUT, "normalize.npz"), **d)
    print("normalize", {k: v.shape for k, v in d.items()})


def main():
    os.makedirs(OUT, exist_ok=True)
    ref = ref_loader.load()
    if len(sys.argv) > 1 and sys.argv[1] == "normalize":   # only this fixture (the others are unchanged)
        return gen_normalize(ref)
    if len(sys.argv) > 1 and sys.argv[1] == "wavenet_variants":
        g = torch.Generator().manual_seed(88)
        kw = dict(blocks=(3, 2), dims=32, residuals_dim=32, skips_dim=32, mlp_dim=32)
        net = ref_loader.make_wavenet(seed=11, pad_side=1, **kw)
        gen_network("wavenet_pad_side1", net, torch.randint(0, 256, (3, 24), generator=g), 32, dict(kw, pad_side=1))
        net = ref_loader.make_wavenet(seed=12, layerwise_inputs=True, **kw)
        gen_network("wavenet_layerwise_inputs", net, torch.randint(0, 256, (3, 24), generator=g), 32,
                    dict(kw, layerwise_inputs=1))
        kw2 = dict(blocks=(4,), dims=32, mlp_dim=32)
        net = ref_loader.make_wavenet(seed=13, layerwise_inputs=True, n_mlp_layers=2, **kw2)
        gen_network("wavenet_layerwise_noskip_mlp2", net, torch.randint(0, 256, (2, 30), generator=g), 28,
                    dict(kw2, layerwise_inputs=1, n_mlp_layers=2))
        kw3 = dict(blocks=(2, 2), dims=32, residuals_dim=32, skips_dim=32, mlp_dim=32)
        net = ref_loader.make_wavenet(seed=14, kernel_sizes=(3,), **kw3)
        gen_network("wavenet_kernel3", net, torch.randint(0, 256, (3, 30), generator=g), 28, dict(kw3, kernel_size=3))
        return
    if len(sys.argv) > 1 and sys.argv[1] == "wavenet_reversed":      # reverse_layer_order (wavenet_v2.py:206, 270)
        g = torch.Generator().manual_seed(89)
        kw = dict(blocks=(3, 2), dims=32, residuals_dim=32, skips_dim=32, mlp_dim=32)
        net = ref_loader.make_wavenet(seed=15, reverse_layer_order=True, **kw)
        gen_network("wavenet_reversed", net, torch.randint(0, 256, (3, 24), generator=g), 32, dict(kw, reverse_layer_order=1))
        return
    if len(sys.argv) > 1 and sys.argv[1] == "wavenet_nongated":      # act_g=None (wavenet_v2.py:109-112, 160-163)
        g = torch.Generator().manual_seed(90)
        kw = dict(blocks=(3, 2), dims=32, residuals_dim=32, skips_dim=32, mlp_dim=32)
        net = ref_loader.make_wavenet(seed=16, gated=False, **kw)
        gen_network("wavenet_nongated", net, torch.randint(0, 256, (3, 24), generator=g), 32, dict(kw, nongated=1))
        return
    if len(sys.argv) > 1 and sys.argv[1] == "wavenet_groups":        # grouped dilated convs (wavenet_v2.py:92; FreqNet uses groups=8)
        g = torch.Generator().manual_seed(91)
        kw = dict(blocks=(3, 2), dims=32, residuals_dim=32, skips_dim=32, mlp_dim=32)
        net = ref_loader.make_wavenet(seed=17, groups=4, **kw)
        gen_network("wavenet_groups4", net, torch.randint(0, 256, (3, 24), generator=g), 32, dict(kw, groups=4))
        return
    if len(sys.argv) > 1 and sys.argv[1] == "wavenet_affine":        # with_affine_residuals (wavenet_v2.py:121-122, 148-149; parametrized.py:34-47)
        g = torch.Generator().manual_seed(92)
        kw = dict(blocks=(3, 2), dims=32, residuals_dim=32, skips_dim=32, mlp_dim=32)
        net = ref_loader.make_wavenet(seed=18, with_affine_residuals=True, **kw)
        gen_network("wavenet_affine_res", net, torch.randint(0, 256, (3, 24), generator=g), 32, dict(kw, affine=1))
        kw2 = dict(blocks=(4,), dims=32, mlp_dim=32)                 # no residual / skip convs, not gated, pad_side=1
        net = ref_loader.make_wavenet(seed=19, with_affine_residuals=True, gated=False, pad_side=1, **kw2)
        gen_network("wavenet_affine_plain", net, torch.randint(0, 256, (2, 20), generator=g), 28,
                    dict(kw2, affine=1, nongated=1, pad_side=1))
        return
    if len(sys.argv) > 1 and sys.argv[1] == "wavenet_acts":          # act_f / act_g other than Tanh / Sigmoid (wavenet_v2.py:198-199, 224-225;
        g = torch.Generator().manual_seed(93)                        # modules/activations.py:26-67): every point-wise member of ActivationEnum
        kw = dict(blocks=(3, 2), dims=32, residuals_dim=32, skips_dim=32, mlp_dim=32)
        for i, (f, gt) in enumerate([("Mish", "Softplus"), ("Sin", "Cos"), ("ReLU", "Identity"), ("Abs", "Tanh"), ("Sigmoid", None)]):
            net = ref_loader.make_wavenet(seed=20 + i, act_f=f, act_g=gt or "Sigmoid", gated=gt is not None, **kw)
            meta = dict(kw, act_f=f, **(dict(act_g=gt) if gt else dict(nongated=1)))
            gen_network(f"wavenet_act_{f.lower()}_{(gt or 'none').lower()}", net, torch.randint(0, 256, (2, 24), generator=g), 16, meta)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "no_temperature":        # MLP(min_temperature=None): Q outputs, no learned temperature (mlp.py:29, 54-62)
        g = torch.Generator().manual_seed(94)
        kw = dict(blocks=(3, 2), dims=32, residuals_dim=32, skips_dim=32, mlp_dim=32)
        net = ref_loader.make_wavenet(seed=30, min_temperature=None, **kw)
        gen_network("wavenet_no_temperature", net, torch.randint(0, 256, (2, 24), generator=g), 16, dict(kw, no_temperature=1))
        gen_samplernn_variant("samplernn_no_temperature", torch.randint(0, 256, (2, 24), generator=g), 16, no_temperature=True,
                              frame_sizes=(4, 2, 1), hidden_dim=32, mlp_dim=32, seed=31)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "wavenet_noblocks":      # blocks=() (wavenet_v2.py:304-307: one layer per kernel size, dilations their
        g = torch.Generator().manual_seed(95)                        # running product; :216 `n != sum(blocks) - 1` then never drops a conv_res)
        kw = dict(blocks=(), dims=32, residuals_dim=32, skips_dim=32, mlp_dim=32)
        net = ref_loader.make_wavenet(seed=32, kernel_sizes=(2, 3, 2, 2), **kw)
        assert "layers.3.conv_res.weight" in net.state_dict()
        gen_network("wavenet_noblocks", net, torch.randint(0, 256, (2, 30), generator=g), 16, dict(kw, kernel_sizes=(2, 3, 2, 2)))
        return
    if len(sys.argv) > 1 and sys.argv[1] == "wavenet_nobias":        # bias=False (wavenet_v2.py:92-93): no bias on conv_dil / conv_skip / conv_res (aff_res keeps its own, :122)
        g = torch.Generator().manual_seed(97)
        kw = dict(blocks=(3, 2), dims=32, residuals_dim=32, skips_dim=32, mlp_dim=32)
        net = ref_loader.make_wavenet(seed=35, bias=False, with_affine_residuals=True, **kw)
        assert not any(k.startswith("layers.") and k.endswith(".bias") and "aff_res" not in k for k in net.state_dict())   # :122 keeps its own
        gen_network("wavenet_nobias_affine", net, torch.randint(0, 256, (2, 24), generator=g), 16, dict(kw, affine=1, nobias=1))
        return
    if len(sys.argv) > 1 and sys.argv[1] == "wavenet_dropped_res":   # residuals_dim != dims_dilated[0]: WNLayer drops the residual path (wavenet_v2.py:78)
        g = torch.Generator().manual_seed(98)
        kw = dict(blocks=(3, 2), dims=32, residuals_dim=16, skips_dim=32, mlp_dim=32)
        net = ref_loader.make_wavenet(seed=36, **kw)
        assert not any("conv_res" in k for k in net.state_dict())
        gen_network("wavenet_dropped_res", net, torch.randint(0, 256, (2, 24), generator=g), 16, kw)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "wavenet_last_res":      # a conv_res on the layer executed LAST and no skips: the head reads x + conv_res(y)
        g = torch.Generator().manual_seed(99)                        # (wavenet_v2.py:216, 270, 286-292): reverse_layer_order, and blocks=()
        kw = dict(blocks=(3, 2), dims=32, residuals_dim=32, mlp_dim=32)
        net = ref_loader.make_wavenet(seed=37, reverse_layer_order=True, **kw)
        assert "layers.4.conv_res.weight" in net.state_dict() and "layers.0.conv_res.weight" not in net.state_dict()
        gen_network("wavenet_reversed_noskip", net, torch.randint(0, 256, (2, 24), generator=g), 16, dict(kw, reverse_layer_order=1))
        kw = dict(blocks=(), dims=32, residuals_dim=32, mlp_dim=32)
        net = ref_loader.make_wavenet(seed=38, kernel_sizes=(2, 2, 3), **kw)
        assert "layers.2.conv_res.weight" in net.state_dict()
        gen_network("wavenet_noblocks_noskip", net, torch.randint(0, 256, (2, 24), generator=g), 16, dict(kw, kernel_sizes=(2, 2, 3)))
        return
    if len(sys.argv) > 1 and sys.argv[1] == "samplernn_zipmodes":    # inputs_mode "mean" / "static_mix" (modules/io.py:283-313) with the ONE input of
        g = torch.Generator().manual_seed(100)                       # the mu-law path: weights 1 / 1 and softmax of one element, both exactly 1.0
        gen_samplernn_variant("samplernn_static_mix", torch.randint(0, 256, (2, 24), generator=g), 16, frame_sizes=(4, 2, 1),
                              hidden_dim=32, mlp_dim=32, seed=39, rnn_class="lstm", inputs_mode="static_mix")
        gen_samplernn_variant("samplernn_mean", torch.randint(0, 256, (2, 24), generator=g), 16, frame_sizes=(4, 2, 1),
                              hidden_dim=32, mlp_dim=32, seed=40, rnn_class="gru", inputs_mode="mean")
        return
    if len(sys.argv) > 1 and sys.argv[1] == "samplernn_nobias":      # rnn_bias=False (sample_rnn_v2.py:66, 130): no rnn.bias_* parameters
        g = torch.Generator().manual_seed(96)
        gen_samplernn_variant("samplernn_lstm_nobias", torch.randint(0, 256, (2, 24), generator=g), 16, frame_sizes=(4, 2, 1),
                              hidden_dim=32, mlp_dim=32, seed=33, rnn_class="lstm", rnn_bias=False)
        gen_samplernn_variant("samplernn_gru_nobias_2layers", torch.randint(0, 256, (2, 24), generator=g), 16, frame_sizes=(4, 2, 1),
                              hidden_dim=32, mlp_dim=32, seed=34, rnn_class="gru", n_rnn=2, rnn_bias=False)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "samplernn_variants":
        g = torch.Generator().manual_seed(77)
        gen_samplernn_variant("samplernn_lstm_default", torch.randint(0, 256, (3, 40), generator=g), 36,
                              frame_sizes=(8, 2, 1), hidden_dim=32, mlp_dim=32, rnn_class="lstm", seed=5)
        gen_samplernn_variant("samplernn_lstm_2layers_ones", torch.randint(0, 256, (2, 35), generator=g), 30,
                              frame_sizes=(4, 2), hidden_dim=32, mlp_dim=32, rnn_class="lstm", n_rnn=2, h0_init="ones", seed=6)
        gen_samplernn_variant("samplernn_gru_3layers_randn_mlp2", torch.randint(0, 256, (3, 32), generator=g), 30,
                              h0_seed=99, frame_sizes=(8, 2, 1), hidden_dim=32, mlp_dim=32, rnn_class="gru", n_rnn=3,
                              h0_init="randn", n_mlp_layers=2, seed=7)
        gen_samplernn_variant("samplernn_rnn_tanh_mlp1", torch.randint(0, 256, (2, 24), generator=g), 24,
                              frame_sizes=(4, 1), hidden_dim=32, mlp_dim=32, rnn_class="rnn", n_mlp_layers=1, seed=8)
        return
    gen_normalize(ref)
    g = torch.Generator().manual_seed(1234)

    # ---- mu-law (functionals.py:313-373)
    x = torch.rand(1 << 16, generator=g) * 2 - 1
    x[:12] = torch.tensor([0., -0., 1., -1., 1e-30, -1e-30, .5, -.5, 1e-3, -1e-3, 0.999999, -0.999999])
    d = dict(x=x.numpy())
    for q, C in [(256, 1.), (256, .5), (64, 2.), (1024, 1.)]:
        d[f"idx_q{q}_c{C}"] = ref.MuLawCompress(q, C).torch_func(x).numpy()
        d[f"expand_q{q}_c{C}"] = ref.MuLawExpand(q, C).torch_func(torch.arange(q)).numpy()
    xi = torch.randint(-300, 300, (64,), generator=g)  # non-float input is cast to fp32 first (:332-333)
    d["x_int"], d["idx_int_q256_c1.0"] = xi.numpy(), ref.MuLawCompress(256, 1.).torch_func(xi).numpy()
    np.savez_compressed(os.path.join(OUT, "mulaw.npz"), **d)
    print("mulaw", len(d))

    # ---- STFT magnitudes (functionals.py:450-528, 576-606)
    d = {}
    for tag, L, n_fft, hop in [("a", 6000, 512, 128), ("b", 9000, 2048, 512), ("c", 2100, 256, 64)]:
        xs = torch.rand(2, L, generator=g) * 2 - 1
        d[f"x_{tag}"] = xs.numpy()
        d[f"cfg_{tag}"] = np.asarray([n_fft, hop])
        for center in (True, False):
            d[f"mag_{tag}_center{int(center)}"] = ref.MagSpec(n_fft, hop, center=center).torch_func(xs).numpy()
    # frame-count KATs over many lengths (reference tests/test_fft_alignment.py:28-110 style)
    kat = []
    for L in list(range(2048, 2048 + 1100, 37)) + [22050, 220500, 32000]:
        for center in (True, False):
            n = ref.MagSpec(2048, 512, center=center).torch_func(torch.zeros(L)).shape[0]
            kat.append((L, int(center), n))
    d["frame_kat"] = np.asarray(kat)
    np.savez_compressed(os.path.join(OUT, "magspec.npz"), **d)
    print("magspec", {k: v.shape for k, v in d.items()})

    # ---- mel filterbank cross-check (torchaudio; librosa itself is absent => "parity unpinned")
    import torchaudio.functional as AF
    fb = AF.melscale_fbanks(1025, 0., 11025., 128, 22050, norm="slaney", mel_scale="slaney").T.numpy()
    nz = np.nonzero(fb)
    np.savez_compressed(os.path.join(OUT, "mel_fb_torchaudio.npz"), rows=nz[0].astype(np.int16),
                        cols=nz[1].astype(np.int16), vals=fb[nz], shape=np.asarray(fb.shape))

    # ---- structural KATs of WaveNet (reference tests/test_wavenet.py:251-275)
    kat = []
    for blocks, ks in [((3,), (2,)), ((4,), (2,)), ((2, 2), (2,)), ((8, 8, 7, 7), (2,)), ((8, 8, 8, 8), (2,)),
                       ((3,), (2, 2, 2)), ((2, 2), (2, 2)), ((2, 3), (2, 3, 2, 2, 2)), ((), (2, 2, 2))]:
        cfg = ref.WaveNet.Config(io_spec=ref.IOSpec.mulaw_io(ref.IOSpec.MuLawIOConfig(input_module_type="embedding")),
                                 blocks=blocks, kernel_sizes=ks, dims_dilated=(8,))
        net = ref.WaveNet.from_config(cfg)
        kat.append(dict(blocks=list(blocks), kernel_sizes=list(ks), rf=int(net.rf),
                        dilations=[int(l.dilation) for l in net.layers],
                        kernels=[int(l.kernel_size) for l in net.layers]))
    import json
    with open(os.path.join(OUT, "wavenet_structure_kat.json"), "w") as f:
        json.dump(kat, f, indent=1)

    # ---- networks
    net = ref_loader.make_wavenet(blocks=(4,), dims=32, seed=0, mlp_dim=32)
    gen_network("wavenet_default_small", net, torch.randint(0, 256, (3, 40), generator=g), 48,
                dict(blocks=(4,), dims=32, mlp_dim=32))
    net = ref_loader.make_wavenet(blocks=(3, 2), dims=32, residuals_dim=32, skips_dim=32, seed=1, mlp_dim=32)
    gen_network("wavenet_res_skip_small", net, torch.randint(0, 256, (4, 24), generator=g), 40,
                dict(blocks=(3, 2), dims=32, residuals_dim=32, skips_dim=32, mlp_dim=32))
    net = ref_loader.make_wavenet(blocks=(4, 4), dims=64, residuals_dim=64, skips_dim=48, seed=2, mlp_dim=64)
    gen_network("wavenet_res_skip_mid", net, torch.randint(0, 256, (5, 64), generator=g), 64,
                dict(blocks=(4, 4), dims=64, residuals_dim=64, skips_dim=48, mlp_dim=64))
    net = ref_loader.make_samplernn(frame_sizes=(8, 2, 1), hidden_dim=32, seed=0, mlp_dim=32)
    gen_network("samplernn_821_small", net, torch.randint(0, 256, (3, 64), generator=g), 48,
                dict(frame_sizes=(8, 2, 1), hidden_dim=32, mlp_dim=32))
    gen_network("samplernn_821_small_ragged", net, torch.randint(0, 256, (2, 67), generator=g), 40,
                dict(frame_sizes=(8, 2, 1), hidden_dim=32, mlp_dim=32))
    net = ref_loader.make_samplernn(frame_sizes=(16, 4, 2), hidden_dim=64, seed=3, mlp_dim=32)
    gen_network("samplernn_1642_small", net, torch.randint(0, 256, (4, 64), generator=g), 48,
                dict(frame_sizes=(16, 4, 2), hidden_dim=64, mlp_dim=32))
    net = ref_loader.make_samplernn(frame_sizes=(4, 1), hidden_dim=32, seed=4, mlp_dim=32)
    gen_network("samplernn_41_small", net, torch.randint(0, 256, (2, 32), generator=g), 32,
                dict(frame_sizes=(4, 1), hidden_dim=32, mlp_dim=32))


if __name__ == "__main__":
    sys.exit(main())
