"""GPU test of the package's multi-GPU entry point over NCCL: `sharding.generate_sharded` on 2 (or more) B200s must equal
the single-GPU run prompt for prompt (prompts, per-prompt temperatures and noise are sharded together; one all-gather of
the index blocks).  Skipped on a box with one GPU (the gloo tests cover the host logic there)."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, kind, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from mimikit_b200 import IOSpec, SampleRNN, WaveNet, sharding
        torch.manual_seed(0)
        if kind == "wavenet":
            cfg = WaveNet.Config(io_spec=IOSpec.mulaw_io(IOSpec.MuLawIOConfig(input_module_type="embedding", mlp_dim=64)),
                                 blocks=(4, 3), dims_dilated=(64,), residuals_dim=64, skips_dim=64)
            net = WaveNet.from_config(cfg).to(dev)
            P = net.rf + 5
        else:
            cfg = SampleRNN.Config(io_spec=IOSpec.mulaw_io(IOSpec.MuLawIOConfig(mlp_dim=32)), frame_sizes=(8, 2, 1),
                                   hidden_dim=64, rnn_class="gru")
            net = SampleRNN.from_config(cfg).to(dev)
            P = 43
        g = torch.Generator().manual_seed(7)
        B, n = 13, 29                                   # ragged: 13 prompts over 2 ranks = 7 + 6
        prompts = torch.randint(0, 256, (B, P), generator=g).to(dev)
        noise = torch.rand(B, n, generator=g).to(dev)
        tvec = (0.5 + torch.rand(B, generator=g)).to(dev)
        res = {}
        for tag, T in (("argmax", None), ("t", 0.9), ("tvec", tvec)):
            got = sharding.generate_sharded(net, prompts, n, temperature=T, noise=noise)
            assert got.shape == (B, P + n) and got.dtype == torch.int64
            res[tag] = got.cpu().numpy()
        if rank == 0:                                   # the same batch on one GPU, no collective
            for tag, T in (("argmax", None), ("t", 0.9), ("tvec", tvec)):
                res["single_" + tag] = net.generate(prompts, n, temperature=T, noise=noise).cpu().numpy()
        q.put((rank, res))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("kind", ["wavenet", "samplernn"])
def test_generate_sharded_over_nccl_equals_single_gpu(kind):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, kind, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=600) for _ in range(world))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for tag in ("argmax", "t", "tvec"):
        want = got[0]["single_" + tag]
        for r in range(world):
            assert np.array_equal(got[r][tag], want), (kind, tag, r)
