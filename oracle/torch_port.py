"""TEST INFRASTRUCTURE ONLY (oracle) — torch-CPU port of the reference's OWN algorithm for the hot path.

This is the CPU arm that bench.py times (`cpu_baseline`, `--impl reference`; kind "port") because the Python
reference cannot travel to the GPU box.  It executes the same torch ops, on the same shapes, as the reference
does per generated sample — in particular WaveNet re-runs the whole network on the last `rf` samples at every
step (mimikit/loops/generate.py:210-211 -> networks/wavenet_v2.py:447-452,276-293; the "fast generate" hooks are
dead code) and SampleRNN runs one Python `generate_step` per sample incl. the prompt warm-up
(networks/sample_rnn_v2.py:226-260).  It is written functionally over a reference state_dict, not as nn.Modules.
The product package never imports it.
"""
import torch
import torch.nn.functional as F


def _head(y, sd, Q):
    """MLP (networks/mlp.py:44-63), n_hidden_layers=0."""
    p = "output_modules.0.estimator.0."
    z = F.linear(F.mish(F.linear(y, sd[p + "fc.0.weight"], sd[p + "fc.0.bias"])),
                 sd[p + "fc.2.weight"], sd[p + "fc.2.bias"])
    temp = torch.sigmoid(z[..., -1:])
    return z[..., :-1] / torch.maximum(temp, sd[p + "min_temp"])


def _sample(logits, temperature, u=None):
    """CategoricalSampler.forward (modules/targets.py:40-52).  With `u` the draw follows the noise-driven
    inverse-CDF contract (oracle/restate.py); otherwise torch.multinomial exactly as the reference."""
    if temperature is None:
        return logits.argmax(dim=-1)
    if u is not None:
        from . import restate
        idx = restate.sample_inverse_cdf(logits.reshape(logits.shape[0], -1).numpy(),
                                         restate.normalize_temperature(temperature, logits.shape[0]), u.numpy())
        return torch.from_numpy(idx).reshape(logits.shape[:-1])
    T = torch.as_tensor(temperature, dtype=logits.dtype).reshape(-1, *([1] * (logits.ndim - 1)))
    l = logits / T
    l = l - l.logsumexp(-1, keepdim=True)
    return torch.multinomial(l.reshape(-1, l.shape[-1]).exp_(), 1).reshape(*logits.shape[:-1])


class WaveNetPort:
    def __init__(self, state_dict, dilations, q_levels=256):
        self.sd = {k: v.detach().clone().float() for k, v in state_dict.items()}
        self.dil = [int(d) for d in dilations]
        self.Q = q_levels
        self.rf = sum(self.dil) + 1
        self.has_skips = "layers.0.conv_skip.weight" in self.sd

    def window_logits(self, x):
        """x (B, T>=rf) int64 -> logits (B, T-rf+1, Q); WaveNet.forward in train mode (wavenet_v2.py:276-293)."""
        sd = self.sd
        h = F.embedding(x, sd["input_modules.0.0.weight"]).transpose(1, 2).contiguous()
        skips = None
        for l, d in enumerate(self.dil):
            p = f"layers.{l}."
            a = F.conv1d(h, sd[p + "conv_dil.0.0.weight"], sd[p + "conv_dil.0.0.bias"], dilation=d)
            f, g = torch.chunk(a, 2, dim=1)
            y = torch.tanh(f) * torch.sigmoid(g)
            if self.has_skips:
                s = F.conv1d(y, sd[p + "conv_skip.weight"], sd[p + "conv_skip.bias"])
                skips = s if skips is None else s + skips[:, :, d:]
            if p + "conv_res.weight" in sd:
                h = h[:, :, d:] + F.conv1d(y, sd[p + "conv_res.weight"], sd[p + "conv_res.bias"])
            else:
                h = y
        y = (skips if self.has_skips else h).transpose(1, 2).contiguous()
        return _head(y, sd, self.Q)

    @torch.no_grad()
    def generate(self, prompts, n_steps, temperature=None, noise=None):
        B, P = prompts.shape
        x = torch.cat([prompts, torch.zeros(B, n_steps, dtype=prompts.dtype)], 1)
        rf = self.rf
        for t in range(P, P + n_steps):
            logits = self.window_logits(x[:, t - rf:t])[:, :1]
            u = None if noise is None else noise[:, t - P]
            x[:, t:t + 1] = _sample(logits, temperature, u)
        return x


class SampleRNNPort:
    def __init__(self, state_dict, frame_sizes, q_levels=256):
        self.sd = {k: v.detach().clone().float() for k, v in state_dict.items()}
        self.fs = tuple(frame_sizes)
        self.Q = q_levels
        self.rf = self.fs[0]
        n = len(self.fs)
        self.up = [self.fs[i] // (self.fs[i + 1] if i < n - 2 else 1) for i in range(n - 1)]
        self.H = self.sd["tiers.0.rnn.weight_hh_l0"].shape[1]

    def _lin(self, q):
        return ((q.float() / self.Q) - .5) * 2

    def _tier(self, i, inpt, prev, hid):
        sd, p = self.sd, f"tiers.{i}."
        x = F.linear(self._lin(inpt), sd[p + "input_module.heads.0.2.weight"], sd[p + "input_module.heads.0.2.bias"])
        if prev is not None:
            x = x + prev
        h = torch.gru_cell(x, hid[i], sd[p + "rnn.weight_ih_l0"], sd[p + "rnn.weight_hh_l0"],
                           sd[p + "rnn.bias_ih_l0"], sd[p + "rnn.bias_hh_l0"])
        hid[i] = h
        return F.linear(h, sd[p + "up_sampler.fc.weight"], sd[p + "up_sampler.fc.bias"]).reshape(-1, self.up[i], self.H)

    def _frame_tiers(self, window, t, hid, O):
        fs = self.fs
        for i in range(len(fs) - 1):
            if t % fs[i] == 0:
                prev = None if i == 0 else O[i - 1][:, (t // fs[i]) % (fs[i - 1] // fs[i])]
                O[i] = self._tier(i, window[:, -fs[i]:], prev, hid)

    @torch.no_grad()
    def generate(self, prompts, n_steps, temperature=None, noise=None):
        sd, fs, rf = self.sd, self.fs, self.rf
        B, P = prompts.shape
        hid = [torch.zeros(B, self.H) for _ in range(len(fs) - 1)]
        O = [None] * (len(fs) - 1)
        offset = P % rf
        plen = P - offset
        for t in range(rf, plen):
            self._frame_tiers(prompts[:, t + offset - rf:t + offset], t, hid, O)
        x = torch.cat([prompts, torch.zeros(B, n_steps, dtype=prompts.dtype)], 1)
        pc = f"tiers.{len(fs) - 1}.input_module.heads.0.2.2.cv."
        Wc, bc = sd[pc + "weight"][:, 0, :], sd[pc + "bias"]
        for t in range(P, P + n_steps):
            window = x[:, t - rf:t]
            self._frame_tiers(window, t, hid, O)
            h = F.linear(self._lin(window[:, -fs[-1]:]), Wc, bc) + O[-1][:, (t % fs[-2]) - fs[-2]]
            logits = _head(h, sd, self.Q)
            u = None if noise is None else noise[:, t - P]
            x[:, t] = _sample(logits, temperature, u)
        return x


def mulaw_compress(x, q_levels=256, compression=1.0):
    """MuLawCompress.torch_func (features/functionals.py:330-338), same torch ops in the same order."""
    mu = torch.tensor(q_levels - 1.0, dtype=x.dtype)
    C = torch.tensor(compression, dtype=x.dtype)
    x_mu = torch.sign(x) * torch.log1p(mu * torch.abs(x) * C) / torch.log1p(mu * C)
    return ((x_mu + 1) / 2 * mu + 0.5).to(torch.int64)


def magspec(x, n_fft=2048, hop=512, center=True):
    """MagSpec.torch_func (features/functionals.py:507-524) for alignment='end'."""
    from . import restate
    t = restate.stft_target_length(x.shape[-1], n_fft, hop, center)
    x = x[..., -t:] if t else x
    S = torch.stft(x, n_fft, hop_length=hop, return_complex=True, center=center,
                   window=torch.hann_window(n_fft), pad_mode="constant")
    return S.transpose(-1, -2).contiguous().abs()


def melspec(mag, fb):
    """mel_basis @ S (features/functionals.py:665-668 via librosa), torch matmul on CPU."""
    return mag @ fb.T
