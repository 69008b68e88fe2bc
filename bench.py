#!/usr/bin/env python
"""bench.py — generated audio samples/s of the batched autoregressive generation path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload wavenet|samplernn|features]

Default workload = BASELINE.json configs[1]: WaveNet mu-law q=256, blocks (8,8,7,7) = 30 layers, 128 channels,
batch 64 prompts of 1 s, generating 10 s at 16 kHz per prompt on each GPU (weak scaling: every rank generates for
its own 64 prompts; one all_gather of the uint8 outputs at the end of the step, no collective inside it).
A "step" = one pass of the hot path over one batch: prefill + all n autoregressive samples for every prompt.
--workload samplernn is BASELINE.json configs[2]: a FIXED batch of 128 prompts sharded over the ranks (strong scaling);
--workload features is configs[4]: every rank extracts its own 10 h shard (weak scaling).

The JSON line (rank 0) follows the driver contract: value (device-timed, inputs resident in HBM), e2e (same metric
through the public GenerateLoopV2 API from pinned HOST buffers, H2D/D2H inside the timed region), roofline,
cpu_baseline, clocks, gpu_launches.  `--impl reference` times the reference's own CPU algorithm (oracle torch port;
the Python reference itself cannot travel to the GPU box) on a bounded sample of the same workload.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SR = 16000
W30 = dict(blocks=(8, 8, 7, 7), dims=128, residuals_dim=128, skips_dim=128, mlp_dim=128)
S3 = dict(frame_sizes=(8, 2, 1), hidden_dim=512, mlp_dim=128)
# algorithmic work per generated sample per prompt
FLOP_PER_SAMPLE = {"wavenet": 2 * 2_982_016, "samplernn": 2 * 1_476_224}
# work of the prefill inside the same launch, per prompt sample it pushes through the net (no head, no sampler):
# WaveNet: the 30 layers (2 932 736 MAC) for the last rf = 765 prompt samples; SampleRNN: the frame tiers (tier 0 every 8 samples, tier 1
# every 2: GRUs + up-samplers 1 376 256 + frame Linears 1 024 = 1 377 280 MAC) for every prompt sample (before_generate's warm-up)
PREFILL_FLOP_PER_SAMPLE = {"wavenet": 2 * 2_932_736, "samplernn": 2 * 1_377_280}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="wavenet", choices=["wavenet", "samplernn", "features"])
    ap.add_argument("--batch", type=int, default=None, help="prompts per GPU (default: 64 wavenet, 128/N samplernn)")
    ap.add_argument("--seconds", type=float, default=10.0, help="generated audio per prompt")
    ap.add_argument("--prompt-seconds", type=float, default=1.0)
    ap.add_argument("--temperature", type=float, default=None, help="default: argmax (what GenerateLoopV2 does)")
    ap.add_argument("--dtype", default="f32", choices=["f32", "bf16"],
                    help="f32 = FFMA kernels (bit-exact sequences, the default); bf16 = tcgen05 tensor-core kernels (logits "
                         "within 5e-2): WaveNet layer pipeline, SampleRNN frame tiers (<= 128 prompts per GPU)")
    ap.add_argument("--cpu-budget-s", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the short side measurements of BASELINE configs 3, 4, 5")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------------------
def make_network(workload, device=None):
    import torch
    from mimikit_b200 import IOSpec, SampleRNN, WaveNet
    torch.manual_seed(0)
    if workload == "wavenet":
        cfg = WaveNet.Config(io_spec=IOSpec.mulaw_io(IOSpec.MuLawIOConfig(sr=SR, input_module_type="embedding",
                                                                          mlp_dim=W30["mlp_dim"])),
                             blocks=W30["blocks"], dims_dilated=(W30["dims"],), residuals_dim=W30["residuals_dim"],
                             skips_dim=W30["skips_dim"])
        net = WaveNet.from_config(cfg)
    else:
        cfg = SampleRNN.Config(io_spec=IOSpec.mulaw_io(IOSpec.MuLawIOConfig(sr=SR, mlp_dim=S3["mlp_dim"])),
                               frame_sizes=S3["frame_sizes"], hidden_dim=S3["hidden_dim"], rnn_class="gru")
        net = SampleRNN.from_config(cfg)
    return net.to(device) if device is not None else net


def synthetic_prompts(B, P, rank=0):
    """two-sine + noise mix, peak-normalised, mu-law 256 — built on the host with torch."""
    import math
    import torch
    g = torch.Generator().manual_seed(1234 + rank)
    t = torch.arange(P, dtype=torch.float64) / SR
    phi = 2 * math.pi * torch.arange(B, dtype=torch.float64)[:, None] / max(B, 1)
    x = 0.6 * torch.sin(2 * math.pi * 220 * t[None] + phi) + 0.3 * torch.sin(2 * math.pi * 659 * t[None]) \
        + 0.05 * torch.randn(B, P, generator=g, dtype=torch.float64)
    x = (x / x.abs().amax(dim=1, keepdim=True)).float()
    mu = 255.0
    xm = torch.sign(x) * torch.log1p(mu * x.abs()) / math.log1p(mu)
    return ((xm + 1) / 2 * mu + 0.5).to(torch.int64)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "200", "-i", str(self.idx)], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.f.read().splitlines():
            c = [v.strip() for v in line.split(",")]
            if len(c) < 8:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.f.name)
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


# ------------------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the reference's own CPU algorithm (oracle/torch_port.py), bounded sample
# ------------------------------------------------------------------------------------------------------------
def cpu_port(workload, state_dict, net):
    from oracle import torch_port
    if workload == "wavenet":
        return torch_port.WaveNetPort(state_dict, net.dilations)
    return torch_port.SampleRNNPort(state_dict, net.frame_sizes)


def cpu_sample_steps(workload):
    # WaveNet: the reference recomputes the receptive field per sample (~0.7 s/step at B=64 on 8 cores) -> few steps;
    # SampleRNN: ms per step, plus its warm-up over the prompt.
    return 8 if workload == "wavenet" else 1024


def time_cpu(workload, net, prompts, n_gen, n_full=None):
    """(samples/s, seconds spent) of the CPU port.  WaveNet: every sample recomputes the receptive field, so the cost per
    sample is constant and `n_gen` samples give the rate.  SampleRNN: the reference's before_generate walks the whole prompt
    once (warm-up) before the first sample; charging that to a short sample of the horizon would bias the rate low, so the
    warm-up (n_steps = 0) and the steady-state samples are timed separately and the rate is the one of the FULL workload:
    B * n_full / (T_warm + n_full * t_step)."""
    import torch
    torch.set_num_threads(os.cpu_count() or 1)
    port = cpu_port(workload, net.state_dict(), net)
    B = prompts.shape[0]
    t0 = time.perf_counter()
    if workload == "samplernn":
        port.generate(prompts, 0, None)
        t_warm = time.perf_counter() - t0
        port.generate(prompts, n_gen, None)
        dt = time.perf_counter() - t0
        t_step = max(dt - 2 * t_warm, 1e-9) / n_gen
        n_full = n_full or n_gen
        return B * n_full / (t_warm + n_full * t_step), dt, {"t_warm_s": t_warm, "t_step_us": t_step * 1e6}
    port.generate(prompts, n_gen, None)
    dt = time.perf_counter() - t0
    return B * n_gen / dt, dt, {}


def run_reference(args):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = args.workload
    B = args.batch or (64 if wl == "wavenet" else 128)
    P, n_full = int(SR * args.prompt_seconds), int(SR * args.seconds)
    net = make_network(wl)
    prompts = synthetic_prompts(B, P)
    n_gen = cpu_sample_steps(wl)
    _, probe_dt, _ = time_cpu(wl, net, prompts, n_gen, n_full)   # CPU warm-up pass; also sizes each step to ~8 s of CPU work
    n_gen = max(n_gen, min(n_full, int(n_gen * 8.0 / max(probe_dt, 1e-3))))
    times, vals = [], []
    for _ in range(args.steps):
        v, dt, _ = time_cpu(wl, net, prompts, n_gen, n_full)
        times.append(dt); vals.append(v)
    dt = statistics.mean(times)
    val = statistics.mean(vals)
    cores = torch.get_num_threads()
    line = {
        "impl": "reference", "metric": "generated audio samples/sec", "value": val, "unit": "samples/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(wl, B, P, n_full, args.gpus),
        "cpu_baseline": {"value": val, "unit": "samples/s", "cores": cores, "kind": "port",
                         "sample": f"first {n_gen} autoregressive samples of the same {B}-prompt batch per step "
                                   f"(step cost is constant in t; full workload is {n_full} samples per prompt"
                                   + ("; prompt warm-up timed separately and charged once per full horizon)" if wl == "samplernn" else ")")},
        "e2e": {"value": val, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def workload_config(wl, B, P, n, n_gpus):
    if wl == "wavenet":
        return {"workload": f"WaveNet mu-law q=256 blocks=(8,8,7,7) 30 layers 128 res/skip channels, batch {B} "
                            f"prompts/GPU x {P} prompt samples -> {n} generated samples each (16 kHz)",
                "batch_per_gpu": B, "prompt_len": P, "n_steps": n, "decode": "argmax",
                "sharding": f"prompts x{n_gpus}", "l2": "256 MB scratch written between timed steps"}
    return {"workload": f"SampleRNN (8,2,1) GRU-512 mu-law q=256, batch {B} prompts/GPU x {P} prompt samples -> {n} "
                        f"generated samples each (16 kHz)",
            "batch_per_gpu": B, "prompt_len": P, "n_steps": n, "decode": "argmax",
            "sharding": f"prompts x{n_gpus}", "l2": "256 MB scratch written between timed steps"}


# ------------------------------------------------------------------------------------------------------------
# ncu `dram__bytes_read.sum + dram__bytes_write.sum` of ONE launch of the dominant kernel, from the `--set full` capture
# under profiles/ (the persistent generation kernels keep weights on chip: traffic is ring spill + mailboxes + outputs)
NCU_TRAFFIC = {}
try:
    with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as _f:
        NCU_TRAFFIC = json.load(_f)
except (OSError, ValueError):
    pass


def latency_floor_us(wl, sm_mhz):
    """SURVEY §8(d) asks for the dependency-chain floor next to the throughput roofline: dependent stages x the latency
    a stage cannot go below in this decomposition.  WaveNet (wavenet6.cu): 30 layers x (2 DSMEM hops of ~275 cycles
    [215 transit + mbarrier wake, B300_MICROARCH.md] + the FFMA issue time of the two critical contractions of a 4-prompt
    group on one SM: 2048 + 1024 warp-FMAs at 4 per cycle = 768 cycles) + ~5 000 cycles of head, sampler and the L2
    feedback of the sampled index.  SampleRNN: see DESIGN.md §4.2."""
    if wl == "wavenet":
        return (30 * (2 * 275 + 768) + 5000) / sm_mhz
    return None


def run_b200(args):
    import torch
    import torch.distributed as dist
    from mimikit_b200 import GenerateLoopV2, sharding

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    wl = args.workload
    B = args.batch or (64 if wl == "wavenet" else max(1, 128 // world))
    P, n = int(SR * args.prompt_seconds), int(SR * args.seconds)
    net = make_network(wl, dev)
    if args.dtype == "bf16":
        net.bfloat16()
    # the GLOBAL prompt batch (world x B prompts, rank r's block generated with r's seed); every rank holds it, as
    # sharding.generate_sharded expects, and generates for its own block
    prompts_host = torch.cat([synthetic_prompts(B, P, r) for r in range(world)], 0).pin_memory()
    prompts_dev = prompts_host.to(dev)
    lo, hi = sharding.shard_bounds(world * B, world, rank)
    scratch = torch.empty(256 << 20, dtype=torch.uint8, device=dev)     # > L2 (126 MB)
    temp = args.temperature
    gen = torch.Generator(device=dev).manual_seed(4321 + rank)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device():
        # the package's multi-GPU entry point: this rank's block through ONE persistent kernel launch, then the single
        # collective of the path (all-gather of the uint8 index blocks)
        return sharding.generate_sharded(net, prompts_dev, n, temperature=temp, generator=gen)

    # ---- warm-up (also: per-step device timestamps for the p50 step latency) ----
    p50_us = None
    for w in range(max(args.warmup, 1)):
        if w == 0:
            _, ts = net.generate(prompts_dev[lo:hi], n, temperature=temp, generator=gen, return_step_timestamps=True)
            d = (ts[1:] - ts[:-1]).double()
            d = d[-n + 1:] if d.numel() >= n else d     # generation steps only (the prefill steps come first)
            p50_us = float(d.median()) / 1e3
        else:
            step_device()
        scratch.zero_()
    # ---- timed: exactly K steps ----
    clocks = ClockSampler(local_rank)
    barrier()
    clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step_device()
        scratch.zero_()
    e1.record()
    barrier()
    ck = clocks.stop()
    ms = e0.elapsed_time(e1)
    t_ms = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms = float(t_ms)
    value = world * B * n * args.steps / (ms / 1e3)

    # ---- e2e: public API (GenerateLoopV2), pinned host prompts in, host waveform out; at N > 1 the loop's real output
    #      block goes through the package's gather (sharding.gather_sequences) ----
    cfg = GenerateLoopV2.Config(parameters=None if temp is None else {"temperature": temp}, display_waveform=False,
                                yield_inversed_outputs=False)
    out_host = torch.empty((B, P + n), dtype=torch.float32).pin_memory()
    inv = net.config.io_spec.targets[0].inv

    def step_e2e():
        loop = GenerateLoopV2(cfg, net, n, [[torch.arange(lo, hi), prompts_host[lo:hi]]])
        for outs in loop.run():
            seq = outs[0]
            if world > 1:
                sharding.gather_sequences(seq, world * B, net.q_levels)
            out_host.copy_(inv(seq), non_blocking=False)

    step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
        scratch.zero_()
    barrier()
    e2e_s = time.perf_counter() - t0
    t_e = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
    e2e_val = world * B * n * args.steps / float(t_e)

    extras = None if args.no_extras else run_extras(args, dev, rank, world, barrier)

    if rank == 0:
        peaks = {}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                peaks = json.load(f)
        except OSError:
            pass
        peak_tf = peaks.get("bf16_tflops_sustained", 1400.0)
        n_prefill = min(P, net.rf) if wl == "wavenet" else P
        flop_launch = FLOP_PER_SAMPLE[wl] * B * n + PREFILL_FLOP_PER_SAMPLE[wl] * B * n_prefill
        kernel_ms = ms / args.steps            # the persistent kernel IS the step (prefill included)
        ach = flop_launch / (kernel_ms / 1e3) / 1e12
        sm_mhz = ck["sm_mhz"] or peaks.get("sm_max_mhz", 1965.0)
        fp32_peak = 148 * 128 * 2 * sm_mhz * 1e6 / 1e12
        info = net.launch_info(B)
        # bf16: the layer-pipelined tcgen05 kernel hosts 16-prompt groups on one CTA per layer; anything else is the fall-back
        tc_name = "wavenet7_kernel" if (info.get("group_size") == 16 and info.get("n_stages", 0) > 1) else "wavenet_tc_kernel"
        kname = {"wavenet": tc_name if args.dtype == "bf16" else "wavenet6_kernel",
                 "samplernn": "samplernn_cluster_kernel" + ("<2>" if args.dtype == "bf16" else "")}[wl]
        tr = NCU_TRAFFIC.get(kname)
        if args.dtype == "bf16":
            roof = {"bound": "tensor", "kernel": kname, "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf,
                    "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained" if peaks else "fallback 1.4 PFLOP/s sustained"}
        else:
            roof = {"bound": "fp32_fma", "kernel": kname, "achieved": ach, "peak": fp32_peak, "unit": "TFLOP/s",
                    "frac": ach / fp32_peak,
                    "peak_source": "148 SMs x 128 FFMA lanes x 2 flop x the SM clock sampled during the run: the path computes in "
                                   "fp32 FFMA (bit-exact argmax parity), so this, not the tensor peak, is its ceiling",
                    "tensor_bf16": {"peak": peak_tf, "frac": ach / peak_tf},
                    "latency_floor_us": latency_floor_us(wl, sm_mhz),
                    "latency_floor_note": "dependency-chain floor of one generated sample in this decomposition (see "
                                          "bench.py latency_floor_us); p50_step_latency_us is the measured counterpart"}
        # dram bytes of one launch from the committed ncu capture (profiles/ncu_traffic.json), scaled from the captured launch to
        # this one by the samples a launch pushes through the net (prompt prefill + generated, per prompt: sequence reads, ring
        # spill, mailboxes, sequence writes all go with them); null when no capture is committed
        roof["flops_counted"] = (f"{B} prompts x ({n} generated samples x {FLOP_PER_SAMPLE[wl]} + {n_prefill} prefill samples x "
                                 f"{PREFILL_FLOP_PER_SAMPLE[wl]}) flop per launch (the launch covers the prefill)")
        roof["traffic"] = None if not tr else tr["dram_bytes"] * (B * (P + n)) / max(1, tr["prompts"] * (tr["prompt_len"] + tr["n_steps"]))
        if tr:
            roof["traffic_source"] = tr.get("source")
        line = {
            "metric": "generated audio samples/sec", "value": value, "unit": "samples/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong" if (wl == "samplernn" and args.batch is None) else "weak", "vs_baseline": None,
            "dtype": args.dtype, "data": "synthetic",
            "config": workload_config(wl, B, P, n, world),
            "p50_step_latency_us": p50_us,
            "e2e": {"value": e2e_val, "unit": "samples/s", "h2d_bytes_per_step": B * P * 8,
                    "d2h_bytes_per_step": B * (P + n) * 4},
            "gpu_launches": 2 * args.steps,
            "launch": info,
            "clocks": {"sm_mhz": ck["sm_mhz"], "sm_max_mhz": ck["sm_max_mhz"], "reasons": ck["reasons"]},
            "roofline": roof,
        }
        if extras:
            line["other_configs"] = extras
        if not args.no_cpu_baseline:
            blk = prompts_host[lo:hi]
            n_gen = cpu_sample_steps(wl)
            _, probe_dt, _ = time_cpu(wl, net, blk, n_gen, n)             # probe, then size the sample to ~15 s
            n_gen = max(n_gen, min(n, int(n_gen * 15.0 / max(probe_dt, 1e-3))))
            cpu_val, cpu_dt, cpu_parts = time_cpu(wl, net, blk, n_gen, n)
            line["cpu_baseline"] = {
                "value": cpu_val, "unit": "samples/s", "cores": torch.get_num_threads(), "kind": "port",
                "sample": f"first {n_gen} autoregressive samples of the same {B}-prompt batch ({cpu_dt:.1f} s); the "
                          f"port runs the reference's own per-sample algorithm (oracle/torch_port.py)"
                          + ("; prompt warm-up and steady-state steps timed separately, rate extrapolated to the full "
                             f"{n}-sample horizon" if wl == "samplernn" else ""), **cpu_parts}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_extras(args, dev, rank, world, barrier):
    """Short measurements of the other BASELINE.json configs on the same ranks, so that the driver's run sees them next to
    the headline: cfg 3 (SampleRNN, 128 prompts sharded over the ranks), cfg 4 (WaveNet bf16 on tcgen05, 128 prompts per
    GPU), cfg 5 (mu-law + STFT/mel, 1 h per GPU).  Device-timed, max over ranks, a fraction of a second each."""
    import torch
    import torch.distributed as dist
    from mimikit_b200 import MagSpec, MelSpec, MuLawCompress, sharding
    out = {}

    def timed(fn, reps=1):
        fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1) / reps], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    try:
        # cfg 3: strong scaling of a fixed 128-prompt batch
        net = make_network("samplernn", dev)
        Bg, P, n = 128, SR // 4, SR // 2
        pr = synthetic_prompts(Bg, P).to(dev)
        ms = timed(lambda: sharding.generate_sharded(net, pr, n))
        out["cfg3_samplernn_b128_sharded"] = {"value": Bg * n / (ms / 1e3), "unit": "samples/s", "ms": ms, "scaling": "strong",
                                              "dtype": "f32", "workload": f"SampleRNN (8,2,1) GRU-512, 128 prompts over {world} GPU(s), "
                                                                          f"{P}-sample prompt -> {n} samples, one gather"}
        if world == 1:          # the same config with the frame tiers on tcgen05 (bf16 operands; one M = 128 tile of prompts)
            net.bfloat16()
            ms = timed(lambda: sharding.generate_sharded(net, pr, n))
            out["cfg3_samplernn_b128_bf16"] = {"value": Bg * n / (ms / 1e3), "unit": "samples/s", "ms": ms, "scaling": "strong",
                                               "dtype": "bf16", "workload": f"SampleRNN (8,2,1) GRU-512 frame tiers on tcgen05, 128 prompts, "
                                                                            f"{P}-sample prompt -> {n} samples"}
        del net
        # cfg 4: tensor-core WaveNet, 128 prompts per GPU
        net = make_network("wavenet", dev).bfloat16()
        Bg, n = 128 * world, SR // 2          # long enough that the 1 s prefill does not dominate the figure
        pr = torch.cat([synthetic_prompts(128, P, r) for r in range(world)], 0).to(dev)
        ms = timed(lambda: sharding.generate_sharded(net, pr, n))
        out["cfg4_wavenet_bf16_b128_per_gpu"] = {"value": Bg * n / (ms / 1e3), "unit": "samples/s", "ms": ms, "scaling": "weak",
                                                 "dtype": "bf16", "workload": f"WaveNet W-30 on tcgen05, {Bg} prompts over {world} GPU(s), "
                                                                              f"{P}-sample prompt -> {n} samples, one gather"}
        del net
        # cfg 5: features, 1 h of 22.05 kHz audio per GPU
        n_clips, L = 360, FEAT["clip"]
        g = torch.Generator(device=dev).manual_seed(99 + rank)
        x = torch.rand((n_clips, L), generator=g, device=dev) * 2 - 1
        mu, ms_, mel = MuLawCompress(256, 1.), MagSpec(FEAT["n_fft"], FEAT["hop"]), MelSpec(FEAT["n_mels"])
        ms = timed(lambda: (mu(x), ms_.mel(x, mel)), reps=5)
        out["cfg5_features_1h_per_gpu"] = {"value": world * n_clips * L / (ms / 1e3), "unit": "samples/s", "ms": ms, "scaling": "weak",
                                           "dtype": "f32", "workload": f"mu-law (int64) + STFT(2048/512)->mel(128), {n_clips} clips x {L} "
                                                                       f"samples per GPU, inputs resident in HBM"}
    except Exception as e:   # the headline line must survive a failing side measurement
        out["error"] = f"{type(e).__name__}: {e}"
    return out


# ------------------------------------------------------------------------------------------------------------
# features workload (BASELINE.json configs[4]): mu-law + fused STFT->mel over 10 h of 22.05 kHz audio
# ------------------------------------------------------------------------------------------------------------
FEAT = dict(sr=22050, clip=220500, clips_10h=3600, n_fft=2048, hop=512, n_mels=128)


def time_cpu_features(steps, n_clips=36):
    """The reference's CPU feature chain (torch ops of MuLawCompress / MagSpec / MelSpec, oracle/torch_port.py) on a bounded
    sample: 36 clips = 6 min of audio per step.  Returns (samples/s, seconds per step)."""
    import torch
    from oracle import restate, torch_port
    torch.set_num_threads(os.cpu_count() or 1)
    g = torch.Generator().manual_seed(1234)
    x = torch.rand(n_clips, FEAT["clip"], generator=g) * 2 - 1
    fb = torch.from_numpy(restate.mel_filterbank(FEAT["n_fft"], FEAT["n_mels"]))

    def one():
        torch_port.mulaw_compress(x)
        torch_port.melspec(torch_port.magspec(x, FEAT["n_fft"], FEAT["hop"]), fb)
    one()
    t0 = time.perf_counter()
    for _ in range(steps):
        one()
    dt = (time.perf_counter() - t0) / steps
    return n_clips * FEAT["clip"] / dt, dt


def run_features_reference(args):
    import torch
    if int(os.environ.get("RANK", "0")) != 0:
        return
    n_clips = 36
    val, dt = time_cpu_features(args.steps, n_clips)
    print(json.dumps({
        "impl": "reference", "metric": "feature-extracted audio samples/sec", "value": val, "unit": "samples/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "mu-law q=256 + STFT(2048/512)->mag->mel(128), 22.05 kHz"},
        "cpu_baseline": {"value": val, "unit": "samples/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": f"{n_clips} clips x {FEAT['clip']} samples per step (full workload: 3600 clips)"},
        "e2e": {"value": val, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}), flush=True)


def run_features(args):
    import torch
    import torch.distributed as dist
    from mimikit_b200 import MagSpec, MelSpec, MuLawCompress

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n_clips = args.batch or FEAT["clips_10h"]      # weak scaling: every rank extracts its own 10 h shard
    L = FEAT["clip"]
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    x = torch.rand((n_clips, L), generator=g, device=dev) * 2 - 1          # 3.18 GB, >> L2
    mu, ms_, mel = MuLawCompress(256, 1.), MagSpec(FEAT["n_fft"], FEAT["hop"]), MelSpec(FEAT["n_mels"])
    n_frames = L // FEAT["hop"] + 1

    def step():
        q = mu(x)
        m = ms_.mel(x, mel)
        return q, m

    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    for _ in range(max(args.warmup, 1)):
        step()
    torch.cuda.synchronize()
    clocks = ClockSampler(local_rank)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        ev[i][0].record()
        q = mu(x)
        ev[i][1].record()
        m = ms_.mel(x, mel)
        ev[i][2].record()
        del q, m
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ck = clocks.stop()
    ms = e0.elapsed_time(e1)
    t_ms = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms = float(t_ms)
    value = world * n_clips * L * args.steps / (ms / 1e3)
    mu_ms = statistics.mean(e[0].elapsed_time(e[1]) for e in ev)
    st_ms = statistics.mean(e[1].elapsed_time(e[2]) for e in ev)

    # e2e: pinned host waveform in, host mel + mu-law (uint8) out, for a 1/10 shard (pageable host RAM is finite)
    e_clips = max(1, n_clips // 10)
    xh = torch.rand((e_clips, L)).mul_(2).sub_(1).pin_memory()
    qh = torch.empty((e_clips, L), dtype=torch.int64).pin_memory()
    mh = torch.empty((e_clips, n_frames, FEAT["n_mels"]), dtype=torch.float32).pin_memory()

    def e2e_step():
        xd = xh.to(dev, non_blocking=True)
        qh.copy_(mu(xd), non_blocking=True)
        mh.copy_(ms_.mel(xd, mel), non_blocking=True)
        torch.cuda.synchronize()
    e2e_step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    e2e_s = time.perf_counter() - t0
    t_e = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
    e2e_val = world * e_clips * L * args.steps / float(t_e)

    if rank == 0:
        peaks = {}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                peaks = json.load(f)
        except OSError:
            pass
        peak = peaks.get("hbm_gbs", 6650.0)
        n_samp = n_clips * L
        mu_bytes = 12 * n_samp                                            # fp32 in + int64 out
        st_bytes = 4 * n_samp + n_clips * n_frames * FEAT["n_mels"] * 4   # fp32 in + mel out
        mu_gbs, st_gbs = mu_bytes / (mu_ms / 1e3) / 1e9, st_bytes / (st_ms / 1e3) / 1e9
        dom = ("stft2048_warp_kernel", st_gbs) if st_ms >= mu_ms else ("mulaw_compress_table_kernel", mu_gbs)
        line = {
            "metric": "feature-extracted audio samples/sec", "value": value, "unit": "samples/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"mu-law q=256 (int64 out) + fused STFT(2048/512, hann, centered)->mag->mel(128) "
                                   f"over {n_clips} clips x {L} samples/GPU (22.05 kHz); inputs (3.2 GB) exceed L2",
                       "clips_per_gpu": n_clips, "clip_len": L},
            "e2e": {"value": e2e_val, "unit": "samples/s", "h2d_bytes_per_step": e_clips * L * 4,
                    "d2h_bytes_per_step": e_clips * L * 8 + e_clips * n_frames * FEAT["n_mels"] * 4,
                    "note": f"{e_clips}-clip shard per step"},
            "gpu_launches": 2 * args.steps,
            "clocks": {"sm_mhz": ck["sm_mhz"], "sm_max_mhz": ck["sm_max_mhz"], "reasons": ck["reasons"]},
            "kernels": {"mulaw_compress_table_kernel": {"ms": mu_ms, "GB/s": mu_gbs, "frac": mu_gbs / peak},
                        "stft2048_warp_kernel": {"ms": st_ms, "GB/s": st_gbs, "frac": st_gbs / peak}},
            # dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture of each kernel on 360 clips
            # (profiles/r01_features_ncu_full_s6.txt), scaled to this launch's clip count: the kernels are streaming, traffic
            # is linear in the clips.  STFT: 317.74 + 66.15 MB; mu-law: 317.81 + 585.84 MB (algorithmic: 3.97 / 9.53 GB for 10 h).
            "roofline": {"bound": "hbm", "kernel": dom[0], "achieved": dom[1], "peak": peak, "unit": "GB/s",
                         "frac": dom[1] / peak,
                         "traffic": ((317.74e6 + 66.15e6) if dom[0].startswith("stft") else (317.81e6 + 585.84e6)) * n_clips / 360.0,
                         "traffic_source": "ncu --set full on a 360-clip launch, scaled by the clip count",
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6.65 TB/s"},
        }
        if not args.no_cpu_baseline:
            cpu_val, cpu_dt = time_cpu_features(2)
            line["cpu_baseline"] = {"value": cpu_val, "unit": "samples/s", "cores": torch.get_num_threads(), "kind": "port",
                                    "sample": f"36 clips x {L} samples per step ({cpu_dt:.1f} s; full workload: {n_clips} clips): the "
                                              f"reference's torch ops on the host (oracle/torch_port.py)"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.workload == "features":
        run_features_reference(args) if args.impl == "reference" else run_features(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
