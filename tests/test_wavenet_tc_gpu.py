"""GPU parity tests of the bf16 tensor-core WaveNet kernel (csrc/wavenet_tc.cu: tcgen05.mma + TMEM), through the C ABI.

Bar (BASELINE.json north_star): teacher-forced logits within 5e-2 relative in bf16.  Sequences follow the bf16 logits, so
they are checked for internal consistency (decisions == the oracle's sampler applied to the kernel's own logits; a
generated sequence replays exactly when teacher-forced) rather than against the fp32 golden sequences."""
import ctypes

import numpy as np
import pytest
import torch

from mimikit_b200 import _capi
from oracle import restate
from test_wavenet_gpu import _rel_err, make_net

pytestmark = pytest.mark.gpu

BF16_TOL = 5e-2


@pytest.mark.parametrize("N,K", [(64, 128), (128, 128), (256, 64), (16, 128), (128, 64), (256, 128), (48, 64), (256, 256)])
def test_umma_descriptor_path_is_a_gemm(N, K):
    """D = A . B^T through the kernel's shared-memory descriptors (K-major SWIZZLE_128B), tcgen05.mma and TMEM loads."""
    g = torch.Generator().manual_seed(N * 1000 + K)
    A = torch.randn(128, K, generator=g).cuda()
    B = torch.randn(N, K, generator=g).cuda()
    D = torch.full((128, N), float("nan"), device="cuda")
    cycles = ctypes.c_longlong(0)
    _capi.check(_capi.lib().mmk_tc_gemm_check(A.data_ptr(), B.data_ptr(), D.data_ptr(), N, K, ctypes.byref(cycles),
                                              _capi.stream_ptr()))
    torch.cuda.synchronize()
    ref = A.bfloat16().float() @ B.bfloat16().float().T
    assert torch.isfinite(D).all()
    assert float((D - ref).abs().max()) <= 1e-3 * float(ref.abs().max())
    print(f"tcgen05.mma 128x{N}x16: {cycles.value / (8 * K / 16):.0f} cycles per instruction (issue to commit, 8 x {K // 16})")


def _bf16_net(blocks, dims, skips, mlp, seed=0):
    net = make_net(blocks, dims, dims, skips, mlp, seed=seed)
    orc = restate.WaveNetOracle({k: v.numpy() for k, v in net.state_dict().items()}, blocks)
    return net.bfloat16(), orc


@pytest.mark.parametrize("blocks,dims,skips,mlp,B,n", [((3, 3), 64, 64, 64, 5, 40), ((4, 4), 128, 128, 128, 130, 24),
                                                       ((2, 5), 64, 128, 64, 129, 33), ((8, 8, 7, 7), 128, 128, 128, 64, 24)])
def test_teacher_forced_logits_within_bf16_tolerance(blocks, dims, skips, mlp, B, n):
    net, orc = _bf16_net(blocks, dims, skips, mlp)
    g = torch.Generator().manual_seed(7)
    P = net.rf + 5
    prompts = torch.randint(0, 256, (B, P), generator=g)
    noise = torch.rand(B, n, generator=g)
    for temp in (None, 0.9):
        ref_seq, ref_logits = orc.generate(prompts.numpy(), n, temp, noise.numpy())          # fp32 oracle, its own feedback
        logits, dec = net.teacher_forced(torch.from_numpy(ref_seq), P, temp, noise)           # same inputs, bf16 arithmetic
        logits, dec = logits.cpu().numpy(), dec.cpu().numpy()
        assert np.isfinite(logits).all()
        assert _rel_err(logits, ref_logits) <= BF16_TOL, _rel_err(logits, ref_logits)
        # the sampler contract holds on the kernel's own logits, bit for bit
        if temp is None:
            want = restate.argmax_first(logits)
        else:
            T = restate.normalize_temperature(temp, B)
            want = np.stack([restate.sample_inverse_cdf(logits[:, i], T, noise.numpy()[:, i]) for i in range(n)], 1)
        assert np.array_equal(dec, want)
        assert (dec == ref_seq[:, P:]).mean() > 0.5      # and mostly agrees with the fp32 decisions


def test_generation_replays_and_is_deterministic():
    net, orc = _bf16_net((3, 3), 64, 64, 64)
    g = torch.Generator().manual_seed(11)
    B, n, P = 131, 50, net.rf + 9
    prompts = torch.randint(0, 256, (B, P), generator=g)
    noise = torch.rand(B, n, generator=g)
    tvec = torch.linspace(0.85, 0.999, B)
    seq, logits = net.generate(prompts, n, temperature=tvec, noise=noise, return_logits=True)
    seq2 = net.generate(prompts, n, temperature=tvec, noise=noise)
    assert torch.equal(seq, seq2)
    assert torch.equal(seq[:, :P].cpu(), prompts)
    lg, dec = net.teacher_forced(seq, P, tvec, noise)
    assert torch.equal(dec, seq[:, P:])
    assert torch.equal(lg, logits)
    # the fp32 oracle teacher-forced on this sequence stays within the bf16 tolerance
    _, ref_logits = orc.generate(prompts.numpy(), n, None, None, forced=seq.cpu().numpy())
    assert _rel_err(logits.cpu().numpy(), ref_logits) <= BF16_TOL
    # permuting the prompts permutes the outputs (rows are independent inside the 128-row MMA tile)
    perm = torch.randperm(B, generator=g)
    seqp = net.generate(prompts[perm], n, temperature=tvec[perm], noise=noise[perm])
    assert torch.equal(seqp.cpu(), seq.cpu()[perm])


def test_stepwise_protocol_matches_whole_sequence_launch():
    from mimikit_b200 import GenerateLoopV2
    net, _ = _bf16_net((3, 3), 64, 64, 64)
    g = torch.Generator().manual_seed(3)
    B, n, P = 4, 12, net.rf + 3
    prompts = torch.randint(0, 256, (B, P), generator=g)
    want = net.generate(prompts, n)
    seq = torch.cat([prompts, torch.zeros(B, n, dtype=torch.int64)], 1).cuda()
    net.before_generate((prompts,), 0)
    for t in range(P, P + n):
        out, = net.generate_step((seq[:, t - net.rf:t],), t=t)
        seq[:, t:t + 1] = out
    net.after_generate((seq,), 0)
    assert torch.equal(seq, want)


def test_unsupported_configurations_fail_loudly():
    net = make_net((3,), 128).bfloat16()           # no residual / skip convs: fp32 kernels only
    with pytest.raises(RuntimeError, match="unsupported configuration"):
        net.generate(torch.zeros(2, net.rf + 1, dtype=torch.int64), 4)
    net = make_net((3, 3), 32, 32, 32, 32).bfloat16()   # channel counts other than 64 / 128
    with pytest.raises(RuntimeError, match="unsupported configuration"):
        net.generate(torch.zeros(2, net.rf + 1, dtype=torch.int64), 4)
    assert net.float().generate(torch.zeros(2, net.rf + 1, dtype=torch.int64), 4).shape == (2, net.rf + 5)


@pytest.mark.parametrize("B,n,extra", [(1, 1, 0), (128, 3, 0), (257, 2, 1), (3, 70, 40)])
def test_edge_geometries(B, n, extra):
    """One prompt, exactly one full 128-row group, three groups with a ragged last one, a single step, prompts exactly
    as long as the receptive field: logits within tolerance, replay exact, outputs of dead rows never written."""
    net, orc = _bf16_net((2, 3), 64, 64, 64, seed=5)
    g = torch.Generator().manual_seed(B * 7 + n)
    P = net.rf + extra
    prompts = torch.randint(0, 256, (B, P), generator=g)
    noise = torch.rand(B, n, generator=g)
    seq, logits = net.generate(prompts, n, temperature=0.95, noise=noise, return_logits=True)
    assert tuple(seq.shape) == (B, P + n) and torch.equal(seq[:, :P].cpu(), prompts)
    assert int(seq.min()) >= 0 and int(seq.max()) <= 255
    _, ref_logits = orc.generate(prompts.numpy(), n, None, None, forced=seq.cpu().numpy())
    assert _rel_err(logits.cpu().numpy(), ref_logits) <= BF16_TOL
    lg, dec = net.teacher_forced(seq, P, 0.95, noise)
    assert torch.equal(dec, seq[:, P:]) and torch.equal(lg, logits)


@pytest.mark.parametrize("blocks,dims,skips,mlp,B,n", [((3, 3), 64, 64, 64, 9, 40), ((8, 8, 7, 7), 128, 128, 128, 6, 16)])
def test_logits_against_the_bf16_faithful_oracle(blocks, dims, skips, mlp, B, n):
    """Against oracle.restate.WaveNetBf16Oracle — the kernel's own arithmetic restated on the CPU (bf16 MMA operands,
    fp32 accumulation, bias folding) with exact transcendentals — the kernel must agree far more tightly than the 5e-2
    it is allowed against the fp32 reference: what is left is the MUFU tanh approximation and summation order."""
    net = make_net(blocks, dims, dims, skips, mlp, seed=2)
    sd = {k: v.numpy() for k, v in net.state_dict().items()}
    orc = restate.WaveNetBf16Oracle(sd, blocks)
    net.bfloat16()
    g = torch.Generator().manual_seed(13)
    P = net.rf + 3
    seq = torch.randint(0, 256, (B, P + n), generator=g)
    logits, _ = net.teacher_forced(seq, P)
    want = orc.logits_for(seq.numpy(), P)
    err = _rel_err(logits.cpu().numpy(), want)
    print(f"bf16 kernel vs bf16-faithful oracle, blocks={blocks}: rel err {err:.2e}")
    assert err <= 1e-2, err


# ---------------------------------------------------------------------------------------------------------------------
# the layer-pipelined tensor-core kernel (csrc/wavenet7.cu): weights resident as the MMA M operand, 16-prompt groups
# ---------------------------------------------------------------------------------------------------------------------
def _pipe_net(blocks, seed=0):
    return _bf16_net(blocks, 128, 128, 128, seed=seed)


@pytest.mark.parametrize("cluster", [None, "2", "4", "8", "16"])
def test_pipeline_kernel_geometries(monkeypatch, cluster):
    """Every cluster size (cluster boundaries are crossed through L2 mailboxes pulled in with bulk copies), ragged batch
    (37 prompts = two full 16-prompt groups + 5): logits within the bf16 tolerance of the fp32 oracle and within 1e-2 of
    the bf16-faithful oracle, replay exact, determinism, permutation equivariance, and the one-CTA kernel agrees."""
    if cluster is not None:
        monkeypatch.setenv("MMK_TC_CLUSTER", cluster)
    blocks = (4, 3)
    net, orc = _pipe_net(blocks, seed=4)
    g = torch.Generator().manual_seed(21)
    B, n, P = 37, 30, net.rf + 6
    info = net.launch_info(B)
    assert info["group_size"] == 16 and info["n_stages"] == sum(blocks) + 1 and info["sm_used"] == sum(blocks) + 1, info
    if cluster is not None:
        assert info["cluster_size"] == int(cluster)
    prompts = torch.randint(0, 256, (B, P), generator=g)
    noise = torch.rand(B, n, generator=g)
    tvec = torch.linspace(0.85, 0.999, B)
    seq, logits = net.generate(prompts, n, temperature=tvec, noise=noise, return_logits=True)
    assert torch.equal(seq, net.generate(prompts, n, temperature=tvec, noise=noise))
    assert torch.equal(seq[:, :P].cpu(), prompts) and int(seq.min()) >= 0 and int(seq.max()) <= 255
    lg, dec = net.teacher_forced(seq, P, tvec, noise)
    assert torch.equal(dec, seq[:, P:]) and torch.equal(lg, logits)
    _, ref_logits = orc.generate(prompts.numpy(), n, None, None, forced=seq.cpu().numpy())
    assert _rel_err(logits.cpu().numpy(), ref_logits) <= BF16_TOL
    faithful = restate.WaveNetBf16Oracle({k: v.numpy() for k, v in net.state_dict().items()}, blocks)
    err = _rel_err(logits.cpu().numpy(), faithful.logits_for(seq.cpu().numpy(), P))
    print(f"pipeline kernel vs bf16-faithful oracle (cluster {cluster}): rel err {err:.2e}")
    assert err <= 1e-2, err
    perm = torch.randperm(B, generator=g)
    seqp = net.generate(prompts[perm], n, temperature=tvec[perm], noise=noise[perm])
    assert torch.equal(seqp.cpu(), seq.cpu()[perm])
    # the first tensor-core kernel (one CTA per 128 prompts) on the same weights: same arithmetic up to summation order
    monkeypatch.setenv("MMK_TC_KERNEL", "4")
    net.float(); net.bfloat16()                       # drop the handle: the next call builds the other kernel
    assert net.launch_info(B)["group_size"] == 128
    lg4, _ = net.teacher_forced(seq, P, tvec, noise)
    assert _rel_err(lg4.cpu().numpy(), logits.cpu().numpy()) <= 1e-2


def test_pipeline_kernel_stepwise_and_continuation():
    """ARM protocol one sample at a time == whole-sequence launch; generate_more continues from the rings."""
    net, _ = _pipe_net((3, 3), seed=6)
    g = torch.Generator().manual_seed(3)
    B, n, P = 20, 14, net.rf + 3
    prompts = torch.randint(0, 256, (B, P), generator=g)
    want = net.generate(prompts, n)
    seq = torch.cat([prompts, torch.zeros(B, n, dtype=torch.int64)], 1).cuda()
    net.before_generate((prompts,), 0)
    for t in range(P, P + n):
        out, = net.generate_step((seq[:, t - net.rf:t],), t=t)
        seq[:, t:t + 1] = out
    net.after_generate((seq,), 0)
    assert torch.equal(seq, want)
    first = net.generate(prompts, 6)
    more = net.generate_more(n - 6)
    assert torch.equal(torch.cat([first, more], 1), want)
