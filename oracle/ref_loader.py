"""TEST INFRASTRUCTURE ONLY — loader for the UNMODIFIED reference (ktonal/mimikit 0.4.3).

Nothing in the product package (`mimikit_b200/`) may import this module.  It is used by
`oracle/make_golden.py` (to produce the committed fixtures under `tests/golden/`) and by
`tests/test_checkpoint_export.py` (networks of the live reference exported and reloaded), which skips when
`/root/reference` is absent (it does not exist on the GPU box).

The reference cannot be imported with a plain `import mimikit` in this image:
third-party packages are missing (h5mapper, omegaconf, librosa, pytorch_lightning, IPython,
matplotlib, pydub) and `mimikit/modules/io.py:205` trips the Python >= 3.11 dataclass
mutable-default check.  The recipe below registers inert stub modules for the
missing imports, bypasses `mimikit/__init__.py` (which pulls ui/views/demos) and patches
`ActivationConfig.__hash__`.  No file under /root/reference is modified or copied.
"""
import importlib
import os
import sys
import types

REF_ROOT = os.environ.get("MIMIKIT_REFERENCE_ROOT", "/root/reference")
REF_PKG = os.path.join(REF_ROOT, "mimikit")

_loaded = None


def available() -> bool:
    return os.path.isdir(REF_PKG)


def _stub(name, **attrs):
    parts = name.split(".")
    for i in range(1, len(parts) + 1):
        n = ".".join(parts[:i])
        if n not in sys.modules:
            m = types.ModuleType(n)
            m.__path__ = []
            sys.modules[n] = m
            if i > 1:
                setattr(sys.modules[".".join(parts[:i - 1])], parts[i - 1], m)
    sys.modules[name].__dict__.update(attrs)
    return sys.modules[name]


class _Any:
    def __init__(self, *a, **k):
        self.a, self.k = a, k

    def __call__(self, *a, **k):
        return None


def _process_batch(batch, test, func):
    # semantics inferred from its use at mimikit/loops/generate.py:39,197-200
    if isinstance(batch, (tuple, list)):
        return type(batch)(_process_batch(b, test, func) for b in batch)
    if test(batch):
        return func(batch)
    return batch


def load():
    """Returns a namespace with the reference's hot-path classes."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError(f"reference not found at {REF_PKG}")
    import torch

    class _Feature:
        pass

    class _TypedFile:
        pass

    _stub("h5mapper", Feature=_Feature, TypedFile=_TypedFile, Input=_Any, Getter=_Any,
          AsSlice=_Any, TensorDict=_Any, FileWalker=_Any, process_batch=_process_batch,
          ProgrammableDataset=object)
    _stub("omegaconf", OmegaConf=_Any, ListConfig=list, DictConfig=dict)
    if "librosa" not in sys.modules:
        try:
            import librosa  # noqa: F401
        except Exception:
            _stub("librosa")

    class _LM(torch.nn.Module):
        pass

    class _CB:
        pass

    _stub("pytorch_lightning", LightningModule=_LM, Trainer=object, Callback=_CB)
    _stub("pytorch_lightning.callbacks", Callback=_CB, TQDMProgressBar=_CB)
    _stub("pytorch_lightning.trainer.states", TrainerState=object)
    _stub("pytorch_lightning.utilities", rank_zero_only=lambda f: f)
    _stub("pytorch_lightning.loggers", Logger=object)
    _stub("lightning_fabric.loggers.logger", rank_zero_experiment=lambda f: f)
    _stub("IPython", get_ipython=lambda: None)
    _stub("IPython.display")
    _stub("matplotlib.pyplot")
    _stub("pydub")

    pkg = types.ModuleType("mimikit")
    pkg.__path__ = [REF_PKG]
    sys.modules["mimikit"] = pkg
    for sub in ["features", "modules", "networks", "loops"]:
        p = types.ModuleType("mimikit." + sub)
        p.__path__ = [os.path.join(REF_PKG, sub)]
        sys.modules["mimikit." + sub] = p
        setattr(pkg, sub, p)

    act = importlib.import_module("mimikit.modules.activations")
    act.ActivationConfig.__hash__ = lambda self: id(self)  # py>=3.11 dataclass default check

    ns = types.SimpleNamespace()
    ns.functionals = importlib.import_module("mimikit.features.functionals")
    ns.item_spec = importlib.import_module("mimikit.features.item_spec")
    ns.io_spec = importlib.import_module("mimikit.io_spec")
    ns.wavenet_v2 = importlib.import_module("mimikit.networks.wavenet_v2")
    ns.sample_rnn_v2 = importlib.import_module("mimikit.networks.sample_rnn_v2")
    ns.targets = importlib.import_module("mimikit.modules.targets")
    ns.generate = importlib.import_module("mimikit.loops.generate")
    ns.IOSpec = ns.io_spec.IOSpec
    ns.WaveNet = ns.wavenet_v2.WaveNet
    ns.SampleRNN = ns.sample_rnn_v2.SampleRNN
    ns.GenerateLoopV2 = ns.generate.GenerateLoopV2
    ns.MuLawCompress = ns.functionals.MuLawCompress
    ns.MuLawExpand = ns.functionals.MuLawExpand
    ns.MagSpec = ns.functionals.MagSpec
    ns.STFT = ns.functionals.STFT
    ns.convert = ns.item_spec.convert
    _loaded = ns
    return ns


# ---------------------------------------------------------------------------------------------
# Reference-side drivers used to generate golden vectors
# ---------------------------------------------------------------------------------------------

def make_wavenet(blocks=(4,), dims=128, residuals_dim=None, skips_dim=None, seed=0, pad_side=0,
                 mlp_dim=128, sr=16000, layerwise_inputs=False, n_mlp_layers=0, kernel_sizes=(2,), reverse_layer_order=False,
                 gated=True, groups=1, with_affine_residuals=False, act_f="Tanh", act_g="Sigmoid",
                 min_temperature=1e-4, bias=True):
    import torch
    ref = load()
    torch.manual_seed(seed)
    cfg = ref.WaveNet.Config(
        io_spec=ref.IOSpec.mulaw_io(ref.IOSpec.MuLawIOConfig(sr=sr, input_module_type="embedding",
                                                             mlp_dim=mlp_dim, n_mlp_layers=n_mlp_layers,
                                                             min_temperature=min_temperature)),
        blocks=tuple(blocks), dims_dilated=(dims,), residuals_dim=residuals_dim, skips_dim=skips_dim,
        pad_side=pad_side, layerwise_inputs=layerwise_inputs, kernel_sizes=tuple(kernel_sizes),
        reverse_layer_order=reverse_layer_order, groups=groups, with_affine_residuals=with_affine_residuals,
        act_f=act_f, act_g=act_g if gated else None, bias=bias)
    return ref.WaveNet.from_config(cfg)


def make_samplernn(frame_sizes=(8, 2, 1), hidden_dim=512, seed=0, mlp_dim=128, sr=16000,
                   rnn_class="gru", h0_init="zeros", n_rnn=1, n_mlp_layers=0, min_temperature=1e-4, rnn_bias=True, inputs_mode="sum"):
    import torch
    ref = load()
    torch.manual_seed(seed)
    cfg = ref.SampleRNN.Config(
        io_spec=ref.IOSpec.mulaw_io(ref.IOSpec.MuLawIOConfig(sr=sr, mlp_dim=mlp_dim, n_mlp_layers=n_mlp_layers,
                                                             min_temperature=min_temperature)),
        frame_sizes=tuple(frame_sizes), hidden_dim=hidden_dim, rnn_class=rnn_class, h0_init=h0_init, n_rnn=n_rnn, rnn_bias=rnn_bias, inputs_mode=inputs_mode)
    return ref.SampleRNN.from_config(cfg)


class NoiseSampler:
    """Replacement for CategoricalSampler.forward (mimikit/modules/targets.py:40-52) that draws with
    externally supplied uniform noise via the inverse-CDF contract of SURVEY.md App. A.3, because
    torch.multinomial cannot be driven by external noise.  `noise` is (B, n_steps) fp32; `step` is
    advanced by the caller.  Also records the logits it was given."""

    def __init__(self, noise=None):
        self.noise = noise
        self.step = 0
        self.logits = []

    def __call__(self, logits, *, temperature=None):
        import torch
        from oracle import restate
        self.logits.append(logits.detach().clone().reshape(logits.shape[0], -1))
        if temperature is None:
            return logits.argmax(dim=-1)
        l2 = logits.reshape(logits.shape[0], logits.shape[-1])
        T = ref_as_tensor(temperature, l2)
        u = self.noise[:, self.step]
        idx = restate.sample_inverse_cdf(l2.numpy(), T.numpy().reshape(-1), u.numpy())
        return torch.from_numpy(idx).reshape(*logits.shape[:-1])


def ref_as_tensor(temperature, tensor):
    return load().targets.as_tensor(temperature, tensor)


def run_generate_loop(net, prompts, n_steps, temperature=None, noise=None):
    """Drives the reference exactly as GenerateLoopV2.run does (mimikit/loops/generate.py:184-229)
    but passes `temperature` straight to generate_step (the real loop drops it for WaveNet,
    SURVEY.md §0.10) and swaps the sampler's forward for NoiseSampler.  Returns (sequence (B,P+n)
    int64, logits (B, n_steps, Q) fp32)."""
    import torch
    net.eval()
    sampler = NoiseSampler(noise)
    om = net.output_modules[0]
    orig = om.sampler.forward
    om.sampler.forward = sampler
    try:
        with torch.no_grad():
            net.before_generate((prompts,), 0)
            sampler.logits.clear()
            rf, prior_t = net.rf, prompts.size(1)
            x = torch.cat([prompts, torch.zeros(prompts.size(0), n_steps, dtype=prompts.dtype)], 1)
            params = {} if temperature is None else {"temperature": temperature}
            for t in range(prior_t, prior_t + n_steps):
                sampler.step = t - prior_t
                out = net.generate_step((x[:, t - rf:t],), t=t, **params)
                x[:, t:t + 1] = out[0][:, :1]
            net.after_generate((x,), 0)
    finally:
        om.sampler.forward = orig
    logits = torch.stack(sampler.logits, 1)
    return x, logits


def run_real_generate_loop(net, prompts, n_steps, parameters=None):
    """The unmodified GenerateLoopV2.run (argmax oracle for WaveNet, multinomial for SampleRNN)."""
    import torch
    ref = load()
    cfg = ref.GenerateLoopV2.Config(parameters=parameters, display_waveform=False, write_waveform=False,
                                    yield_inversed_outputs=False)
    loop = ref.GenerateLoopV2(cfg, network=net, n_steps=n_steps,
                              dataloader=[[torch.ones(prompts.size(0)), prompts]], logger=None)
    outs = [o for o in loop.run()]
    return outs[0][0]
