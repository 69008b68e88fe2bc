"""GPU parity tests of the persistent SampleRNN kernel (through the C ABI) against the golden vectors produced by
the live reference (tests/golden/samplernn_*.npz) and against the oracle on seeded inputs."""
import os

import numpy as np
import pytest
import torch

from conftest import golden_state_dict, load_golden
from mimikit_b200 import _capi
from oracle import restate

pytestmark = pytest.mark.gpu

REL_TOL = 1e-3   # north star: teacher-forced logits within 1e-3 relative in fp32


def _rel_err(got, ref):
    return float(np.abs(got - ref).max() / max(1e-6, np.abs(ref).max()))


def make_net(frame_sizes, hidden, mlp_dim=128, sd=None, seed=0, rnn_class="gru"):
    from mimikit_b200 import IOSpec, SampleRNN
    torch.manual_seed(seed)
    cfg = SampleRNN.Config(io_spec=IOSpec.mulaw_io(IOSpec.MuLawIOConfig(mlp_dim=mlp_dim)),
                           frame_sizes=tuple(frame_sizes), hidden_dim=hidden, rnn_class=rnn_class)
    net = SampleRNN.from_config(cfg).to("cuda")
    if sd is not None:
        net.load_state_dict(sd)
    return net


@pytest.mark.parametrize("name", ["samplernn_821_small", "samplernn_821_small_ragged", "samplernn_1642_small",
                                  "samplernn_41_small"])
def test_golden_sequences_and_logits(name):
    d = load_golden(name)
    fs = tuple(int(f) for f in d["meta/frame_sizes"])
    net = make_net(fs, int(d["meta/hidden_dim"]), int(d["meta/mlp_dim"]), golden_state_dict(d))
    prompts, noise = torch.from_numpy(d["prompts"]), torch.from_numpy(d["noise"])
    P, n = prompts.shape[1], noise.shape[1]
    seq, logits = net.generate(prompts, n, return_logits=True)
    assert seq.dtype == torch.int64 and seq.is_cuda and tuple(seq.shape) == d["seq_argmax"].shape
    assert _rel_err(logits.cpu().numpy()[:, 0], d["logits_argmax"][:, 0]) <= REL_TOL
    assert np.array_equal(seq.cpu().numpy(), d["seq_argmax"])            # bit-exact argmax-decoded sequence
    assert _rel_err(logits.cpu().numpy(), d["logits_argmax"]) <= REL_TOL
    seq, logits = net.generate(prompts, n, temperature=1.0, noise=noise, return_logits=True)
    assert np.array_equal(seq.cpu().numpy(), d["seq_t1"])                # sampled with the supplied noise
    assert _rel_err(logits.cpu().numpy(), d["logits_t1"]) <= REL_TOL
    seq = net.generate(prompts, n, temperature=torch.from_numpy(d["tvec"]), noise=noise)
    assert np.array_equal(seq.cpu().numpy(), d["seq_tvec"])
    # step-wise logits on forced inputs (the teacher-forced definition for SampleRNN)
    lg, dec = net.teacher_forced(torch.from_numpy(d["seq_t1"]), P, 1.0, noise)
    assert np.array_equal(dec.cpu().numpy(), d["seq_t1"][:, P:])
    assert _rel_err(lg.cpu().numpy(), d["logits_t1"]) <= REL_TOL


@pytest.mark.parametrize("ctas", [None, "1", "5", "24"])
@pytest.mark.parametrize("fs,H,B,P", [((8, 2, 1), 64, 19, 43), ((4, 4), 32, 3, 16), ((16, 4, 2), 48, 33, 64),
                                      ((8, 4, 2, 1), 32, 6, 27)])
def test_vs_oracle_partitions(monkeypatch, ctas, fs, H, B, P):
    """Seeded weights and prompts; different row partitions (number of CTAs), ragged batches (B not a multiple of
    the 16-prompt chunk), prompt lengths that are not a multiple of the top frame size (warm-up offset quirk)."""
    monkeypatch.setenv("MMK_SR_KERNEL", "1")    # the general (grid-barrier) kernel
    if ctas is not None:
        monkeypatch.setenv("MMK_SR_CTAS", ctas)
    net = make_net(fs, H, mlp_dim=32, seed=5)
    try:
        net.launch_info(B)
    except _capi.MmkError as e:   # this partition cannot host the net (row tile width / shared memory)
        pytest.skip(str(e))
    orc = restate.SampleRNNOracle({k: v.numpy() for k, v in net.state_dict().items()}, fs)
    g = torch.Generator().manual_seed(17)
    n = 37
    prompts = torch.randint(0, 256, (B, P), generator=g)
    noise = torch.rand(B, n, generator=g)
    for temp in (None, 0.95):
        seq, logits = net.generate(prompts, n, temperature=temp, noise=noise, return_logits=True)
        ref_seq, ref_logits = orc.generate(prompts.numpy(), n, temp, noise.numpy())
        assert _rel_err(logits.cpu().numpy()[:, 0], ref_logits[:, 0]) <= REL_TOL
        assert np.array_equal(seq.cpu().numpy(), ref_seq), (ctas, temp)
        assert _rel_err(logits.cpu().numpy(), ref_logits) <= REL_TOL


@pytest.mark.parametrize("fs,H,B,P", [((8, 2, 1), 64, 19, 43), ((4, 4), 32, 3, 16), ((16, 4, 2), 48, 33, 64),
                                      ((8, 4, 2, 1), 32, 6, 27), ((8, 2, 1), 512, 128, 40)])
def test_default_geometry_hosts_every_net(fs, H, B, P):
    """No environment overrides: the launcher's own choice of kernel and geometry must host every net of this file — a
    geometry that should fit but regressed is a FAILURE here, not a skip."""
    for var in ("MMK_SR_KERNEL", "MMK_SR_CLUSTER", "MMK_SR_CTAS"):
        assert var not in os.environ
    net = make_net(fs, H, mlp_dim=32, seed=5)
    info = net.launch_info(B)
    assert info["sm_used"] >= 1 and info["threads"] >= 32
    orc = restate.SampleRNNOracle({k: v.numpy() for k, v in net.state_dict().items()}, fs)
    g = torch.Generator().manual_seed(17)
    prompts = torch.randint(0, 256, (min(B, 5), P), generator=g)
    seq = net.generate(prompts, 9)
    ref_seq, _ = orc.generate(prompts.numpy(), 9)
    assert np.array_equal(seq.cpu().numpy(), ref_seq)


@pytest.mark.parametrize("cluster", ["1", "2", "4", "8"])
@pytest.mark.parametrize("ctas", [None, "8"])
@pytest.mark.parametrize("fs,H,B,P", [((8, 2, 1), 64, 19, 43), ((4, 4), 32, 3, 16), ((16, 4, 2), 48, 33, 64),
                                      ((8, 4, 2, 1), 32, 6, 27), ((8, 2, 1), 64, 70, 40), ((4, 2), 128, 9, 10)])
def test_cluster_kernel_vs_oracle(monkeypatch, cluster, ctas, fs, H, B, P):
    """The cluster kernel (samplernn2.cu: cluster-local head over distributed shared memory, streamed tier
    contractions) at every cluster size, small grids (several prompt groups per cluster), ragged batches and the
    warm-up offset quirk; free-running argmax / sampled sequences bit-exact, logits within tolerance, and the
    teacher-forced decisions on the oracle's own sequence."""
    monkeypatch.setenv("MMK_SR_KERNEL", "2")
    monkeypatch.setenv("MMK_SR_CLUSTER", cluster)
    if ctas is not None:
        monkeypatch.setenv("MMK_SR_CTAS", ctas)
    net = make_net(fs, H, mlp_dim=32, seed=5)
    try:
        info = net.launch_info(B)
    except _capi.MmkError as e:   # this geometry cannot host the net
        pytest.skip(str(e))
    assert info["cluster_size"] == int(cluster)
    orc = restate.SampleRNNOracle({k: v.numpy() for k, v in net.state_dict().items()}, fs)
    g = torch.Generator().manual_seed(17)
    n = 37
    prompts = torch.randint(0, 256, (B, P), generator=g)
    noise = torch.rand(B, n, generator=g)
    for temp in (None, 0.95):
        seq, logits = net.generate(prompts, n, temperature=temp, noise=noise, return_logits=True)
        ref_seq, ref_logits = orc.generate(prompts.numpy(), n, temp, noise.numpy())
        assert _rel_err(logits.cpu().numpy()[:, 0], ref_logits[:, 0]) <= REL_TOL
        assert np.array_equal(seq.cpu().numpy(), ref_seq), (cluster, temp)
        assert _rel_err(logits.cpu().numpy(), ref_logits) <= REL_TOL
    lg, dec = net.teacher_forced(torch.from_numpy(ref_seq), P, 0.95, noise)
    assert np.array_equal(dec.cpu().numpy(), ref_seq[:, P:])
    assert _rel_err(lg.cpu().numpy(), ref_logits) <= REL_TOL


@pytest.mark.parametrize("cluster", [None, "1", "2", "8"])
@pytest.mark.parametrize("fs,H,B,P", [((8, 2, 1), 512, 37, 43), ((8, 2, 1), 512, 128, 24), ((4, 4), 256, 3, 16), ((8, 4, 2, 1), 128, 22, 27),
                                      ((4, 2), 128, 9, 10), ((8, 4, 2), 256, 70, 40), ((2, 2, 1), 128, 5, 9), ((2, 1, 1), 128, 4, 8),
                                      ((16, 8, 8), 256, 9, 48), ((16, 8, 8), 512, 40, 35), ((16, 2, 1), 128, 6, 33)])
def test_lane_major_engine_vs_oracle(monkeypatch, cluster, fs, H, B, P):
    """The lane-major frame-tier engine of samplernn2.cu (H in {128, 256, 512}: weights in registers, [prompt][H] rows
    prefetched into registers, transposing shuffle trees): every supported K-quarter count, up-sampling factor and frame
    size, ragged batches, all cluster sizes of the head; sequences bit-exact, logits in tolerance."""
    if cluster is not None:
        monkeypatch.setenv("MMK_SR_CLUSTER", cluster)
    net = make_net(fs, H, mlp_dim=32, seed=5)
    try:
        info = net.launch_info(B)
    except _capi.MmkError as e:   # this cluster size cannot host the net
        pytest.skip(str(e))
    assert info["threads"] == 256 and info["sm_used"] == H // 4, info     # the lane-major engine, one CTA per 4 hidden units
    orc = restate.SampleRNNOracle({k: v.numpy() for k, v in net.state_dict().items()}, fs)
    g = torch.Generator().manual_seed(17)
    n = 21
    prompts = torch.randint(0, 256, (B, P), generator=g)
    noise = torch.rand(B, n, generator=g)
    sub = list(range(B)) if B <= 40 else [0, 1, B // 2, B - 2, B - 1]
    for temp in (None, 0.95):
        seq, logits = net.generate(prompts, n, temperature=temp, noise=noise, return_logits=True)
        ref_seq, ref_logits = orc.generate(prompts[sub].numpy(), n, temp, noise[sub].numpy())
        assert _rel_err(logits.cpu().numpy()[sub][:, 0], ref_logits[:, 0]) <= REL_TOL
        assert np.array_equal(seq.cpu().numpy()[sub], ref_seq), (cluster, temp)
        assert _rel_err(logits.cpu().numpy()[sub], ref_logits) <= REL_TOL
    lg, dec = net.teacher_forced(torch.from_numpy(ref_seq), P, 0.95, noise[sub])
    assert np.array_equal(dec.cpu().numpy(), ref_seq[:, P:])
    assert _rel_err(lg.cpu().numpy(), ref_logits) <= REL_TOL


@pytest.mark.parametrize("fs,H,B,P,mlp", [((8, 2, 1), 512, 37, 48, 32), ((8, 2, 1), 512, 128, 24, 32), ((4, 4), 256, 3, 16, 32),
                                          ((8, 4, 2, 1), 128, 22, 32, 32), ((2, 2, 1), 128, 5, 10, 32), ((8, 4, 2), 256, 70, 40, 32),
                                          ((8, 2, 1), 512, 50, 24, 128), ((8, 2, 1), 256, 9, 16, 64), ((16, 8, 8), 256, 12, 48, 128)])
def test_tensor_core_mode_vs_oracle(fs, H, B, P, mlp):
    """compute_dtype bfloat16: the frame tiers' GRU and up-sampler contractions on tcgen05 (bf16 operands, fp32 accumulation in
    TMEM, frame Linear folded into the gate), cell / head / sampler in fp32.  Teacher-forced logits within the north star's 5e-2
    of the fp32 oracle, decisions = argmax of the kernel's own logits, free-running generation deterministic and consistent with
    its own logits.  The geometries cover both forms of the head: with the first Linear folded into the bottom tier's up-sampler
    (when up * mlp_dim rows split evenly over the H / 4 CTAs, e.g. BASELINE's 512 / 128) and without (512 / 32)."""
    net = make_net(fs, H, mlp_dim=mlp, seed=5).bfloat16()
    info = net.launch_info(B)
    assert info["threads"] == 256 and info["sm_used"] == H // 4, info
    orc = restate.SampleRNNOracle({k: v.numpy() for k, v in net.state_dict().items()}, fs)
    g = torch.Generator().manual_seed(17)
    n = 24
    prompts = torch.randint(0, 256, (B, P), generator=g)
    sub = list(range(B)) if B <= 24 else [0, 1, B // 2, B - 2, B - 1]
    ref_seq, ref_logits = orc.generate(prompts[sub].numpy(), n, None, None)
    full = torch.cat([prompts, torch.zeros(B, n, dtype=torch.long)], 1)
    full[sub] = torch.from_numpy(ref_seq)
    lg, dec = net.teacher_forced(full, P)
    rel = _rel_err(lg.cpu().numpy()[sub], ref_logits)
    print(f"bf16 SampleRNN {fs} H={H} B={B}: teacher-forced logits rel err {rel:.2e}")
    assert rel <= 5e-2
    assert np.array_equal(dec.cpu().numpy(), restate.argmax_first(lg.cpu().numpy()))
    seq, logits = net.generate(prompts, n, return_logits=True)
    seq2 = net.generate(prompts, n)
    assert torch.equal(seq, seq2)
    assert np.array_equal(seq.cpu().numpy()[:, P:], restate.argmax_first(logits.cpu().numpy()))
    # sampled: every decision is the oracle's inverse-CDF draw applied to the kernel's own logits (scalar and per-prompt T)
    noise = torch.rand(B, n, generator=g)
    for T in (torch.tensor([0.9]), torch.linspace(0.8, 1.1, B)):
        sseq, slog = net.generate(prompts, n, temperature=T, noise=noise, return_logits=True)
        sl = slog.cpu().numpy()
        for j in range(n):
            want = restate.sample_inverse_cdf(sl[:, j], T.numpy(), noise[:, j].numpy())
            assert np.array_equal(sseq.cpu().numpy()[:, P + j], want), (j, T.shape)
    # and back: the fp32 engine of the same object still reproduces the oracle bit for bit
    net.float()
    seq32 = net.generate(prompts[sub], n)
    assert np.array_equal(seq32.cpu().numpy(), ref_seq)


@pytest.mark.parametrize("fs,H,B,P,mlp", [((8, 2, 1), 512, 33, 24, 128), ((8, 4, 2, 1), 128, 22, 32, 32), ((4, 4), 256, 7, 16, 64),
                                          ((16, 8, 8), 256, 20, 48, 128)])
def test_tensor_core_mode_lstm(fs, H, B, P, mlp):
    """nn.LSTM tiers — the reference's DEFAULT rnn_class (sample_rnn_v2.py:40-66) — on the tcgen05 engine: gates i, f, g, o as the
    16 accumulator columns of a CTA, the cell state carried in fp32.  Same criteria as the GRU form; fp32 mode of the same net
    (general kernel) still reproduces the oracle bit for bit."""
    net = make_net(fs, H, mlp_dim=mlp, seed=7, rnn_class="lstm").bfloat16()
    info = net.launch_info(B)
    assert info["threads"] == 256 and info["sm_used"] == H // 4, info
    orc = restate.SampleRNNOracle({k: v.numpy() for k, v in net.state_dict().items()}, fs, rnn_class="lstm")
    g = torch.Generator().manual_seed(23)
    n = 24
    prompts = torch.randint(0, 256, (B, P), generator=g)
    sub = list(range(B)) if B <= 24 else [0, 1, B // 2, B - 2, B - 1]
    ref_seq, ref_logits = orc.generate(prompts[sub].numpy(), n, None, None)
    full = torch.cat([prompts, torch.zeros(B, n, dtype=torch.long)], 1)
    full[sub] = torch.from_numpy(ref_seq)
    lg, dec = net.teacher_forced(full, P)
    rel = _rel_err(lg.cpu().numpy()[sub], ref_logits)
    print(f"bf16 SampleRNN LSTM {fs} H={H} B={B}: teacher-forced logits rel err {rel:.2e}")
    assert rel <= 5e-2
    assert np.array_equal(dec.cpu().numpy(), restate.argmax_first(lg.cpu().numpy()))
    seq, logits = net.generate(prompts, n, return_logits=True)
    assert torch.equal(seq, net.generate(prompts, n))
    assert np.array_equal(seq.cpu().numpy()[:, P:], restate.argmax_first(logits.cpu().numpy()))
    # chunked continuation carries h and c: two halves == one launch
    a = net.generate(prompts, n // 2)
    b = net.generate_more(n - n // 2)
    assert torch.equal(torch.cat([a, b], 1), seq)
    net.float()
    try:
        seq32 = net.generate(prompts[sub], n)
    except _capi.MmkError as e:       # the general fp32 kernel keeps every weight resident: LSTM-512 does not fit it
        assert "shared memory" in str(e) and H == 512
    else:
        assert np.array_equal(seq32.cpu().numpy(), ref_seq)


@pytest.mark.parametrize("fs,H,B,P", [((8, 2, 1), 512, 37, 40), ((8, 4, 2, 1), 128, 22, 27), ((4, 4), 256, 3, 16), ((2, 2, 1), 128, 5, 9),
                                      ((16, 8, 8), 256, 9, 48), ((16, 8, 8), 128, 33, 37)])
def test_lane_major_lstm_vs_oracle(fs, H, B, P):
    """nn.LSTM tiers (the reference's default rnn_class) on the lane-major fp32 engine: 16 homogeneous gate columns fed by [x | h],
    cell state in fp32.  Sequences bit-exact with the oracle (argmax and sampled), logits within tolerance, teacher-forced
    decisions, chunked continuation == one launch.  (16, 8, 8) / 256 / LSTM is SampleRNN.Config()'s own default
    (sample_rnn_v2.py:124-131): frames of 16, up-sampling by 8, an 8-tap sample tier."""
    net = make_net(fs, H, mlp_dim=32, seed=9, rnn_class="lstm")
    info = net.launch_info(B)
    assert info["threads"] == 256 and info["sm_used"] == H // 4, info
    orc = restate.SampleRNNOracle({k: v.numpy() for k, v in net.state_dict().items()}, fs, rnn_class="lstm")
    g = torch.Generator().manual_seed(31)
    n = 24
    prompts = torch.randint(0, 256, (B, P), generator=g)
    noise = torch.rand(B, n, generator=g)
    sub = list(range(B)) if B <= 24 else [0, 1, B // 2, B - 2, B - 1]
    for temp in (None, 0.95):
        seq, logits = net.generate(prompts, n, temperature=temp, noise=noise, return_logits=True)
        ref_seq, ref_logits = orc.generate(prompts[sub].numpy(), n, temp, noise[sub].numpy())
        assert np.array_equal(seq.cpu().numpy()[sub], ref_seq), temp
        assert _rel_err(logits.cpu().numpy()[sub], ref_logits) <= REL_TOL
    lg, dec = net.teacher_forced(torch.from_numpy(ref_seq), P, 0.95, noise[sub])
    assert np.array_equal(dec.cpu().numpy(), ref_seq[:, P:])
    a = net.generate(prompts, n // 2)
    b = net.generate_more(n - n // 2)
    assert torch.equal(torch.cat([a, b], 1), net.generate(prompts, n))


@pytest.mark.parametrize("rnn,h0_init,H,mode", [("gru", "ones", 128, "f32"), ("lstm", "randn", 256, "f32"), ("lstm", "ones", 128, "bf16"),
                                                ("gru", "randn", 256, "bf16")])
def test_initial_state_on_the_fast_engines(rnn, h0_init, H, mode):
    """h0_init 'ones' / 'randn' (SampleRNNTier._init_h0, sample_rnn_v2.py:101-119) with explicit states: the lane-major (fp32) and
    tensor-core (bf16) engines take them through mmk_samplernn_set_hidden — fp32 bit-exact with the oracle fed the same states,
    bf16 within 5e-2 on teacher-forced logits."""
    from mimikit_b200 import IOSpec, SampleRNN
    fs = (8, 2, 1)
    torch.manual_seed(13)
    cfg = SampleRNN.Config(io_spec=IOSpec.mulaw_io(IOSpec.MuLawIOConfig(mlp_dim=32)), frame_sizes=fs, hidden_dim=H, rnn_class=rnn, h0_init=h0_init)
    net = SampleRNN.from_config(cfg).to("cuda")
    if mode == "bf16":
        net.bfloat16()
    B, P, n = 7, 27, 20
    assert net.launch_info(B)["threads"] == 256          # one of the two fast engines, not the general kernel
    g = torch.Generator().manual_seed(4)
    prompts = torch.randint(0, 256, (B, P), generator=g)
    noise = torch.rand(B, n, generator=g)
    h0 = {}
    for i in range(len(fs) - 1):
        for which in ((0, 1) if rnn == "lstm" else (0,)):
            h0[(i, 0, which)] = torch.ones(B, H) if h0_init == "ones" else torch.randn(B, H, generator=g)
    orc = restate.SampleRNNOracle({k: v.numpy() for k, v in net.state_dict().items()}, fs, rnn_class=rnn)
    ref_seq, ref_logits = orc.generate(prompts.numpy(), n, 0.9, noise.numpy(), h0={k: v.numpy() for k, v in h0.items()})
    if mode == "f32":
        seq, logits = net.generate(prompts, n, temperature=0.9, noise=noise, return_logits=True, h0=h0)
        assert np.array_equal(seq.cpu().numpy(), ref_seq)
        assert _rel_err(logits.cpu().numpy(), ref_logits) <= REL_TOL
    else:
        lg, dec = net.teacher_forced(torch.from_numpy(ref_seq), P, 0.9, noise, h0=h0)
        assert _rel_err(lg.cpu().numpy(), ref_logits) <= 5e-2
    # the states matter: the zero-state run differs
    z_seq, z_logits = orc.generate(prompts.numpy(), n, 0.9, noise.numpy())
    assert _rel_err(z_logits, ref_logits) > 1e-3


def test_tensor_core_mode_without_folds(monkeypatch):
    """MMK_SR_FOLD=0: the tensor-core engine without its algebraic folds (conditioning as a bf16 image and a K = 2H contraction,
    the head's first Linear in the cluster head) — the same criteria, and logits close to the folded form's."""
    fs, H, B, P, n = (8, 2, 1), 512, 19, 24, 16
    net = make_net(fs, H, mlp_dim=128, seed=5).bfloat16()
    g = torch.Generator().manual_seed(17)
    prompts = torch.randint(0, 256, (B, P), generator=g)
    orc = restate.SampleRNNOracle({k: v.numpy() for k, v in net.state_dict().items()}, fs)
    ref_seq, ref_logits = orc.generate(prompts.numpy(), n, None, None)
    folded, _ = net.teacher_forced(torch.from_numpy(ref_seq), P)
    monkeypatch.setenv("MMK_SR_FOLD", "0")
    net2 = make_net(fs, H, mlp_dim=128, seed=5).bfloat16()
    plain, dec = net2.teacher_forced(torch.from_numpy(ref_seq), P)
    assert _rel_err(plain.cpu().numpy(), ref_logits) <= 5e-2
    assert np.array_equal(dec.cpu().numpy(), restate.argmax_first(plain.cpu().numpy()))
    assert _rel_err(plain.cpu().numpy(), folded.cpu().numpy()) <= 5e-3
    assert not torch.equal(plain, folded)          # two different arithmetic forms did run


def test_stepwise_protocol_and_loop():
    """before_generate / generate_step / after_generate == whole-sequence path == oracle; GenerateLoopV2 integration as
    the reference's tests/test_sample_rnn.py:90-112 (batch 2, 512-sample prompt + 512 steps, temperature=(1.,))."""
    from mimikit_b200 import GenerateLoopV2
    fs = (8, 2, 1)
    net = make_net(fs, 32, mlp_dim=32, seed=2)
    orc = restate.SampleRNNOracle({k: v.numpy() for k, v in net.state_dict().items()}, fs)
    g = torch.Generator().manual_seed(3)
    B, P, n = 3, 41, 26
    prompts = torch.randint(0, 256, (B, P), generator=g)
    ref_seq, _ = orc.generate(prompts.numpy(), n)
    x = torch.cat([prompts, torch.zeros(B, n, dtype=torch.int64)], 1).cuda()
    net.before_generate((x[:, :P],), 0)
    for t in range(P, P + n):
        out = net.generate_step((x[:, t - net.rf:t],), t=t)
        assert isinstance(out, tuple) and tuple(out[0].shape) == (B, 1)
        x[:, t:t + 1] = out[0]
    net.after_generate((x,), 0)
    assert np.array_equal(x.cpu().numpy(), ref_seq)

    prompts = torch.randint(0, 256, (2, 512), generator=g)
    cfg = GenerateLoopV2.Config(parameters=dict(temperature=(1.,)), display_waveform=False)
    outs = list(GenerateLoopV2(cfg, net, 512, [[torch.arange(2), prompts]]).run())
    assert len(outs) == 1 and tuple(outs[0][0].shape) == (2, 1024) and outs[0][0].dtype == torch.float32
    assert float(outs[0][0][:, 512:].abs().max()) > 0
    cfg = GenerateLoopV2.Config(display_waveform=False, yield_inversed_outputs=False)
    out = list(GenerateLoopV2(cfg, net, 64, [[torch.arange(2), prompts]]).run())[0][0]
    ref_seq, _ = orc.generate(prompts.numpy(), 64)
    assert np.array_equal(out.cpu().numpy(), ref_seq)


def test_weight_norm_checkpoint_and_errors():
    from mimikit_b200 import IOSpec, SampleRNN
    fs = (4, 1)
    net = make_net(fs, 32, mlp_dim=32, seed=9)
    sd = net.state_dict()
    # a weight_norm'ed checkpoint stores every leaf as *_g / *_v (sample_rnn_v2.py:67-81): fold on load
    wn = {}
    for k, v in sd.items():
        if v.dim() >= 1 and not k.endswith("min_temp"):
            nrm = v.reshape(v.shape[0], -1).norm(dim=1).reshape((-1,) + (1,) * (v.dim() - 1))
            wn[k + "_g"], wn[k + "_v"] = nrm * 1.0, v * 3.0
        else:
            wn[k] = v
    net2 = make_net(fs, 32, mlp_dim=32, seed=123)
    net2.load_state_dict(wn)
    prompts = torch.randint(0, 256, (2, 16), generator=torch.Generator().manual_seed(0))
    a, la = net.generate(prompts, 12, return_logits=True)
    b, lb = net2.generate(prompts, 12, return_logits=True)
    assert _rel_err(lb.cpu().numpy()[:, 0], la.cpu().numpy()[:, 0]) <= 1e-5
    with pytest.raises(RuntimeError):
        net.generate(torch.zeros(2, 3, dtype=torch.int64), 4)      # prompt shorter than the top frame
    with pytest.raises(NotImplementedError):
        SampleRNN.from_config(SampleRNN.Config(io_spec=IOSpec.mulaw_io(IOSpec.MuLawIOConfig()), rnn_class="none"))
    with pytest.raises(NotImplementedError):
        SampleRNN.from_config(SampleRNN.Config(io_spec=IOSpec.mulaw_io(IOSpec.MuLawIOConfig()), rnn_class="gru", n_rnn=5))
    with pytest.raises(NotImplementedError):
        SampleRNN.from_config(SampleRNN.Config(io_spec=IOSpec.mulaw_io(IOSpec.MuLawIOConfig()), inputs_mode="prod"))  # not a ZipMode


def test_s3_full_width_properties():
    """BASELINE cfg 3 geometry ((8,2,1), GRU 512, batch 128): determinism, batch-permutation equivariance, and oracle
    parity on a subset over a short horizon."""
    fs = (8, 2, 1)
    net = make_net(fs, 512, mlp_dim=128, seed=0)
    g = torch.Generator().manual_seed(1234)
    B, P, n = 128, 256, 64
    prompts = torch.from_numpy(restate.synthetic_prompts(B, P))
    noise = torch.rand(B, n, generator=g)
    seq = net.generate(prompts, n, temperature=1.0, noise=noise)
    assert torch.equal(seq, net.generate(prompts, n, temperature=1.0, noise=noise))
    perm = torch.randperm(B, generator=g)
    seq_p = net.generate(prompts[perm], n, temperature=1.0, noise=noise[perm])
    assert torch.equal(seq_p.cpu(), seq.cpu()[perm])
    assert torch.equal(seq[:, :P].cpu(), prompts) and int(seq.max()) <= 255 and int(seq.min()) >= 0
    orc = restate.SampleRNNOracle({k: v.numpy() for k, v in net.state_dict().items()}, fs)
    sub = [0, 77, 127]
    ref_seq, ref_logits = orc.generate(prompts[sub].numpy(), 32, 1.0, noise[sub, :32].numpy())
    got_seq, got_logits = net.generate(prompts[sub], 32, temperature=1.0, noise=noise[sub, :32], return_logits=True)
    assert _rel_err(got_logits.cpu().numpy()[:, 0], ref_logits[:, 0]) <= REL_TOL
    assert np.array_equal(got_seq.cpu().numpy(), ref_seq)
    assert np.array_equal(seq.cpu().numpy()[sub][:, :P + 32], ref_seq)


def test_s3_long_horizon_vs_oracle():
    """BASELINE cfg 3 geometry over a longer horizon (1 000 generated samples = 625 tier firings after the warm-up) on the default
    (lane-major) engine: argmax and sampled sequences stay bit-exact with the oracle, logits within tolerance at the far end."""
    fs = (8, 2, 1)
    net = make_net(fs, 512, mlp_dim=128, seed=3)
    assert net.launch_info(128)["threads"] == 256
    g = torch.Generator().manual_seed(99)
    B, P, n = 128, 64, 1000
    prompts = torch.from_numpy(restate.synthetic_prompts(B, P))
    noise = torch.rand(B, n, generator=g)
    orc = restate.SampleRNNOracle({k: v.numpy() for k, v in net.state_dict().items()}, fs)
    sub = [5, 126]
    for temp in (None, 0.9):
        seq, logits = net.generate(prompts, n, temperature=temp, noise=noise, return_logits=True)
        ref_seq, ref_logits = orc.generate(prompts[sub].numpy(), n, temp, noise[sub].numpy())
        assert np.array_equal(seq.cpu().numpy()[sub], ref_seq), temp
        assert _rel_err(logits.cpu().numpy()[sub][:, -50:], ref_logits[:, -50:]) <= REL_TOL


@pytest.mark.parametrize("name", ["samplernn_lstm_default", "samplernn_lstm_2layers_ones", "samplernn_gru_3layers_randn_mlp2",
                                  "samplernn_rnn_tanh_mlp1", "samplernn_no_temperature", "samplernn_lstm_nobias",
                                  "samplernn_gru_nobias_2layers", "samplernn_static_mix", "samplernn_mean"])
def test_variant_goldens(name):
    """The rest of SampleRNNTier's configuration surface against the live reference (tests/golden, generated by
    oracle/make_golden.py samplernn_variants): rnn_class "lstm" (the reference DEFAULT) / "rnn", n_rnn 2 and 3, h0_init
    "ones" and "randn" (the reference's own draws, replayed), 1 and 2 hidden MLP layers (one shared Linear, mlp.py:47-50)."""
    from mimikit_b200 import IOSpec, SampleRNN
    from test_oracle_golden import variant_setup
    d = load_golden(name)
    m, kw, h0 = variant_setup(d)
    fs = tuple(int(f) for f in m["frame_sizes"])
    head = dict(min_temperature=None) if int(m.get("no_temperature", 0)) else {}      # MLP(min_temperature=None), mlp.py:29, 54-62
    cfg = SampleRNN.Config(io_spec=IOSpec.mulaw_io(IOSpec.MuLawIOConfig(mlp_dim=int(m["mlp_dim"]), n_mlp_layers=kw["n_mlp_hidden"],
                                                                        **head)),
                           frame_sizes=fs, hidden_dim=int(m["hidden_dim"]), rnn_class=kw["rnn_class"], n_rnn=kw["n_rnn"],
                           h0_init=str(m["h0_init"]), rnn_bias=bool(int(m.get("rnn_bias", 1))),
                           inputs_mode=str(m.get("inputs_mode", "sum")))
    net = SampleRNN.from_config(cfg).to("cuda")
    net.load_state_dict(golden_state_dict(d))
    prompts, noise = torch.from_numpy(d["prompts"]), torch.from_numpy(d["noise"])
    P, n = prompts.shape[1], noise.shape[1]
    explicit = None if str(m["h0_init"]) != "randn" else {k: torch.from_numpy(v) for k, v in h0.items()}
    for tag, T in (("argmax", None), ("t1", 1.0)):
        seq, logits = net.generate(prompts, n, temperature=T, noise=noise, return_logits=True, h0=explicit)
        assert _rel_err(logits.cpu().numpy()[:, 0], d["logits_" + tag][:, 0]) <= REL_TOL, (name, tag)
        assert np.array_equal(seq.cpu().numpy(), d["seq_" + tag]), (name, tag)
        assert _rel_err(logits.cpu().numpy(), d["logits_" + tag]) <= REL_TOL
    lg, dec = net.teacher_forced(torch.from_numpy(d["seq_t1"]), P, 1.0, noise, h0=explicit)
    assert np.array_equal(dec.cpu().numpy(), d["seq_t1"][:, P:])
    # the state survives between launches: generate_more continues, the step-wise protocol == the whole-sequence launch
    if str(m["h0_init"]) != "randn":
        first = net.generate(prompts, n // 2)
        more = net.generate_more(n - n // 2)
        assert np.array_equal(torch.cat([first, more], 1).cpu().numpy(), d["seq_argmax"])
        x = torch.cat([prompts, torch.zeros(prompts.shape[0], n, dtype=torch.int64)], 1).cuda()
        net.before_generate((x[:, :P],), 0)
        for t in range(P, P + n):
            x[:, t:t + 1] = net.generate_step((x[:, t - net.rf:t],), t=t)[0]
        assert np.array_equal(x.cpu().numpy(), d["seq_argmax"])


def test_lstm_is_the_default_and_randn_draws():
    """SampleRNN.Config() defaults (rnn_class 'lstm', sample_rnn_v2.py:127) build and generate; h0_init 'randn' draws its
    own states from a generator (reproducibly) when none are passed; bad explicit states are rejected."""
    from mimikit_b200 import IOSpec, SampleRNN
    torch.manual_seed(3)
    cfg = SampleRNN.Config(io_spec=IOSpec.mulaw_io(IOSpec.MuLawIOConfig(mlp_dim=32)), frame_sizes=(8, 2, 1), hidden_dim=48)
    assert cfg.rnn_class == "lstm"
    net = SampleRNN.from_config(cfg).to("cuda")
    orc = restate.SampleRNNOracle({k: v.numpy() for k, v in net.state_dict().items()}, (8, 2, 1), rnn_class="lstm")
    g = torch.Generator().manual_seed(2)
    prompts = torch.randint(0, 256, (21, 50), generator=g)
    noise = torch.rand(21, 20, generator=g)
    seq, lg = net.generate(prompts, 20, temperature=0.9, noise=noise, return_logits=True)
    ref_seq, ref_lg = orc.generate(prompts.numpy(), 20, 0.9, noise.numpy())
    assert np.array_equal(seq.cpu().numpy(), ref_seq) and _rel_err(lg.cpu().numpy(), ref_lg) <= REL_TOL
    cfg2 = SampleRNN.Config(io_spec=IOSpec.mulaw_io(IOSpec.MuLawIOConfig(mlp_dim=32)), frame_sizes=(4, 2), hidden_dim=32,
                            rnn_class="gru", h0_init="randn")
    net2 = SampleRNN.from_config(cfg2).to("cuda")
    gen = lambda seed: net2.generate(prompts[:4], 16, return_logits=True, generator=torch.Generator(device="cuda").manual_seed(seed))
    (a, la), (b, lb), (c, lc) = gen(5), gen(5), gen(6)
    assert torch.equal(a, b) and torch.equal(la, lb) and not torch.equal(la, lc)
    with pytest.raises(ValueError):
        net2.generate(prompts[:4], 4, h0={(0, 0, 0): torch.zeros(3, 32)})


def test_exported_checkpoint_generates_the_golden_sequence(tmp_path):
    """export_network -> file -> load_exported (mimikit_b200/checkpoint.py) of an LSTM network built from the
    live reference's state dict: the reloaded network generates the reference's sequence."""
    from mimikit_b200 import IOSpec, SampleRNN, load_exported, save_exported
    from test_oracle_golden import variant_setup
    d = load_golden("samplernn_lstm_2layers_ones")
    m, kw, _ = variant_setup(d)
    cfg = SampleRNN.Config(io_spec=IOSpec.mulaw_io(IOSpec.MuLawIOConfig(mlp_dim=int(m["mlp_dim"]))),
                           frame_sizes=tuple(int(f) for f in m["frame_sizes"]), hidden_dim=int(m["hidden_dim"]),
                           rnn_class="lstm", n_rnn=2, h0_init="ones")
    net = SampleRNN.from_config(cfg)
    net.load_state_dict(golden_state_dict(d))
    again = load_exported(save_exported(net, str(tmp_path / "m.b200.pt")), device="cuda")
    assert again.config.rnn_class == "lstm" and again.config.n_rnn == 2 and again.config.h0_init == "ones"
    seq = again.generate(torch.from_numpy(d["prompts"]), d["noise"].shape[1])
    assert np.array_equal(seq.cpu().numpy(), d["seq_argmax"])
