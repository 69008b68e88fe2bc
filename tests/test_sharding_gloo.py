"""N > 1 host logic on CPU: world_size-2/3 `gloo` processes run the prompt-sharded generation driver with a
deterministic stand-in network (the kernels need a GPU; the sharding, the per-prompt parameter slicing and the single
gather do not).  The multi-process result must equal the single-process result prompt for prompt."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mimikit_b200 import sharding


class FakeARM:
    """Per-prompt deterministic 'generation': every output depends only on that prompt's row, its noise row and its
    temperature — exactly the independence the real networks have."""
    q_levels = 256

    def __init__(self):
        self.calls = []

    def generate(self, prompts, n_steps, temperature=None, noise=None):
        B, P = prompts.shape
        self.calls.append(B)
        seq = torch.zeros((B, P + n_steps), dtype=torch.int64)
        seq[:, :P] = prompts
        if temperature is None:
            T = torch.zeros(B)
        else:
            T = torch.as_tensor(temperature, dtype=torch.float32).reshape(-1).expand(B)
        for t in range(P, P + n_steps):
            u = noise[:, t - P] if noise is not None else torch.zeros(B)
            seq[:, t] = (seq[:, t - 1] * 7 + seq[:, t - P] + (u * 100).long() + (T * 10).long() + 3) % 256
        return seq


class FakeARM512(FakeARM):
    """A 512-level alphabet announced the way the real networks do it: through config.io_spec.targets[0].out_dim
    (no `q_levels` attribute on the class) — values >= 256 must survive the gather (uint8 truncation)."""
    q_levels = None

    class _T:
        out_dim = 512

    class _IO:
        pass

    class _Cfg:
        pass

    def __init__(self):
        super().__init__()
        io = self._IO(); io.targets = (self._T(),)
        self.config = self._Cfg(); self.config.io_spec = io

    def generate(self, prompts, n_steps, temperature=None, noise=None):
        seq = super().generate(prompts % 256, n_steps, temperature, noise)
        seq[:, prompts.shape[1]:] += 256 * (seq[:, prompts.shape[1]:] % 2)     # half of the outputs land in [256, 512)
        seq[:, :prompts.shape[1]] = prompts
        return seq


def _worker512(rank, world, port, B, P, n, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        prompts, noise, _ = _inputs(B, P, n)
        # numpy, not tensors: a tensor travels through the queue as a shared-memory handle that dies with this process
        q.put((rank, sharding.generate_sharded(FakeARM512(), prompts + 200, n, temperature=0.9, noise=noise).numpy()))
    finally:
        dist.destroy_process_group()


def test_gather_keeps_wide_alphabets():
    B, P, n, world = 5, 6, 7, 2
    prompts, noise, _ = _inputs(B, P, n)
    want = FakeARM512().generate(prompts + 200, n, temperature=0.9, noise=noise)
    assert int(want.max()) >= 256
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker512, args=(r, world, port, B, P, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in range(world):
        assert torch.equal(torch.from_numpy(got[r]), want)
    # a block that does not fit its declared alphabet is an error, never a silent truncation
    with pytest.raises(ValueError):
        class One:
            pass
        import unittest.mock as um
        with um.patch.object(sharding, "_world", lambda group=None: (0, 2)):
            sharding.gather_sequences(torch.full((3, 4), 300), 6, 256)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _inputs(B, P, n):
    g = torch.Generator().manual_seed(7)
    prompts = torch.randint(0, 256, (B, P), generator=g)
    noise = torch.rand(B, n, generator=g)
    tvec = torch.linspace(0.85, 0.999, B)
    return prompts, noise, tvec


def _worker(rank, world, port, B, P, n, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        prompts, noise, tvec = _inputs(B, P, n)
        net = FakeARM()
        res = {}
        res["argmax"] = sharding.generate_sharded(net, prompts, n)
        res["scalar_T"] = sharding.generate_sharded(net, prompts, n, temperature=0.9, noise=noise)
        res["vector_T"] = sharding.generate_sharded(net, prompts, n, temperature=tvec, noise=noise)
        res["local_only"] = sharding.generate_sharded(net, prompts, n, temperature=tvec, noise=noise, gather=False)
        res["calls"] = net.calls
        feats = sharding.extract_sharded(lambda x: x.float() * 2 + 1, prompts, gather=True)
        res["feats"] = feats
        q.put((rank, {k: (v.numpy() if isinstance(v, torch.Tensor) else v) for k, v in res.items()}))   # see _worker512
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,B", [(2, 8), (2, 5), (3, 4)])
def test_generate_sharded_equals_single_process(world, B):
    P, n = 12, 9
    prompts, noise, tvec = _inputs(B, P, n)
    ref = FakeARM()
    want = {
        "argmax": ref.generate(prompts, n),
        "scalar_T": ref.generate(prompts, n, temperature=0.9, noise=noise),
        "vector_T": ref.generate(prompts, n, temperature=tvec, noise=noise),
    }
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, B, P, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in range(world):
        lo, hi = sharding.shard_bounds(B, world, r)
        for k, v in want.items():
            assert torch.equal(torch.from_numpy(got[r][k]), v), (r, k)          # every rank holds the whole batch, in prompt order
        assert torch.equal(torch.from_numpy(got[r]["local_only"]), want["vector_T"][lo:hi])
        assert got[r]["calls"] == [hi - lo] * 4                # one kernel-path call per run, on its own block only
        assert torch.equal(torch.from_numpy(got[r]["feats"]), prompts.float() * 2 + 1)


def test_shard_bounds_partition():
    for n in (0, 1, 5, 64, 128, 1024, 3600):
        for world in (1, 2, 3, 4, 8):
            spans = [sharding.shard_bounds(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))          # contiguous, no gaps, no overlap
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)
    with pytest.raises(ValueError):
        sharding.shard_bounds(4, 2, 2)


def test_single_process_passthrough():
    prompts, noise, tvec = _inputs(4, 6, 5)
    net = FakeARM()
    out = sharding.generate_sharded(net, prompts, 5, temperature=tvec, noise=noise)
    assert torch.equal(out, FakeARM().generate(prompts, 5, temperature=tvec, noise=noise))
    with pytest.raises(ValueError):
        sharding.shard_of(torch.zeros(3), 2, 0, n_units=4)
