"""GPU tests of the orchestration above the loop: chunked long-form generation
(mimikit/loops/generate_chunks.py:39-56) and EnsembleGenerator (mimikit/models/ensemble_generator.py:61-163)."""
import numpy as np
import pytest
import torch

from oracle import restate

pytestmark = pytest.mark.gpu


def _wavenet(seed=0, sr=16000):
    from mimikit_b200 import IOSpec, WaveNet
    torch.manual_seed(seed)
    cfg = WaveNet.Config(io_spec=IOSpec.mulaw_io(IOSpec.MuLawIOConfig(sr=sr, input_module_type="embedding", mlp_dim=64)),
                         blocks=(4, 3), dims_dilated=(64,), residuals_dim=64, skips_dim=64)
    return WaveNet.from_config(cfg).to("cuda"), (4, 3)


def _samplernn(seed=0, sr=16000):
    from mimikit_b200 import IOSpec, SampleRNN
    torch.manual_seed(seed)
    fs = (8, 2, 1)
    cfg = SampleRNN.Config(io_spec=IOSpec.mulaw_io(IOSpec.MuLawIOConfig(sr=sr, mlp_dim=32)), frame_sizes=fs, hidden_dim=64,
                           rnn_class="gru")
    return SampleRNN.from_config(cfg).to("cuda"), fs


@pytest.mark.parametrize("temperature", [None, 0.9])
def test_wavenet_chunks_equal_one_long_generation(temperature):
    """Both modes of generate_chunks: re-prompting with >= rf samples (the reference script) and carrying the rings over
    are the same sequence as ONE long generate — bit for bit, with argmax and with the supplied noise — and the oracle's."""
    from mimikit_b200 import generate_chunks
    net, blocks = _wavenet()
    g = torch.Generator().manual_seed(3)
    B, P, chunks, steps = 9, net.rf + 11, 4, 23
    prompts = torch.randint(0, 256, (B, P), generator=g)
    noise = torch.rand(B, chunks * steps, generator=g)
    long = net.generate(prompts, chunks * steps, temperature=temperature, noise=noise)
    carried = generate_chunks(net, prompts, chunks, steps, temperature=temperature, noise=noise, carry_state=True)
    reprompted = generate_chunks(net, prompts, chunks, steps, prompt_length=net.rf + 3, temperature=temperature, noise=noise)
    assert torch.equal(carried, long) and torch.equal(reprompted, long)
    orc = restate.WaveNetOracle({k: v.numpy() for k, v in net.state_dict().items()}, blocks)
    ref, _ = orc.generate(prompts.numpy(), chunks * steps, temperature, noise.numpy())
    assert np.array_equal(long.cpu().numpy(), ref)
    # the continuation refuses to run on a rebuilt handle
    net.load_state_dict(net.state_dict())
    with pytest.raises(RuntimeError):
        net.generate_more(4)


def test_samplernn_chunks():
    """carry_state=True is one uninterrupted generation; carry_state=False re-prompts like the reference script: hidden
    reset + warm-up over the tail, i.e. the oracle run chunk by chunk on the tails.  A per-prompt temperature vector that
    drifts between chunks (generate_chunks.py:42-43) goes through temperature_update."""
    from mimikit_b200 import generate_chunks
    net, fs = _samplernn()
    g = torch.Generator().manual_seed(5)
    B, P, chunks, steps = 7, 40, 3, 19
    prompts = torch.randint(0, 256, (B, P), generator=g)
    noise = torch.rand(B, chunks * steps, generator=g)
    long = net.generate(prompts, chunks * steps, temperature=0.95, noise=noise)
    carried = generate_chunks(net, prompts, chunks, steps, temperature=0.95, noise=noise, carry_state=True)
    assert torch.equal(carried, long)
    temps = [torch.linspace(0.85, 0.999, B)]
    for i in range(1, chunks):
        temps.append((temps[-1] + 0.01 * i).clamp(0.85, 0.999))
    L = 24
    got = generate_chunks(net, prompts, chunks, steps, prompt_length=L, temperature=temps[0],
                          temperature_update=lambda i, T: temps[i], noise=noise)
    orc = restate.SampleRNNOracle({k: v.numpy() for k, v in net.state_dict().items()}, fs)
    track = prompts.numpy()
    for i in range(chunks):
        src = track if i == 0 else track[:, -L:]
        seq, _ = orc.generate(src, steps, temps[i].numpy(), noise[:, i * steps:(i + 1) * steps].numpy())
        track = np.concatenate([track, seq[:, -steps:]], 1)
    assert np.array_equal(got.cpu().numpy(), track)


def test_ensemble_generator_chains_networks():
    """Two events, two different networks at the base rate: the output is prompt | event 1 | event 2 | silence, every event
    generated from the last `prompt_length` samples of what precedes it (ensemble_generator.py:80-144) — checked against
    the oracle driven by hand through the same transforms."""
    from mimikit_b200 import EnsembleGenerator
    sr = 16000
    wn, blocks = _wavenet(1, sr)
    sr_net, fs = _samplernn(2, sr)
    Pw = 128
    prompt = torch.from_numpy(restate.mulaw_expand(restate.synthetic_prompts(3, Pw)))
    stream = [dict(generator=wn, seconds=40 / sr), dict(generator=sr_net, seconds=24 / sr, temperature=None),
              dict(generator=wn, seconds=1.0)]
    total_s = (Pw + 40 + 24 + 16) / sr
    out = EnsembleGenerator(prompt, max_seconds=total_s, base_sr=sr, stream=iter(stream)).run()
    assert tuple(out.shape) == (3, Pw + 80) and out.dtype == torch.float32 and out.is_cuda
    wave = prompt.numpy()
    o1 = restate.WaveNetOracle({k: v.numpy() for k, v in wn.state_dict().items()}, blocks)
    q = restate.mulaw_compress(wave[:, -Pw:])
    seq, _ = o1.generate(q, 40)
    wave = np.concatenate([wave, restate.mulaw_expand(seq[:, Pw:])], 1)
    o2 = restate.SampleRNNOracle({k: v.numpy() for k, v in sr_net.state_dict().items()}, fs)
    q = restate.mulaw_compress(wave[:, -Pw:])
    seq, _ = o2.generate(q, 24)
    wave = np.concatenate([wave, restate.mulaw_expand(seq[:, Pw:])], 1)
    got = out.cpu().numpy()
    np.testing.assert_allclose(got[:, :wave.shape[1]], wave, atol=1e-6)
    assert not got[:, wave.shape[1]:].any()        # the third event does not fit before max_seconds: silence
