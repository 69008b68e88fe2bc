"""Ad-hoc: SampleRNN (8,2,1) hidden 512 with nn.LSTM tiers (the reference's default rnn_class), 128 prompts: samples/s of the lane-major
fp32 engine and of the tcgen05 engine, next to the GRU form (device-timed, one warm-up)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mimikit_b200 import IOSpec, SampleRNN
from oracle import restate

B, P, n = 128, 16000, 32000
prompts = torch.from_numpy(restate.synthetic_prompts(B, P)).cuda()
for rnn, fs, H in (("gru", (8, 2, 1), 512), ("lstm", (8, 2, 1), 512), ("lstm", (16, 8, 8), 256)):   # the last: SampleRNN.Config()'s defaults
    torch.manual_seed(0)
    cfg = SampleRNN.Config(io_spec=IOSpec.mulaw_io(IOSpec.MuLawIOConfig(sr=16000, mlp_dim=128)), frame_sizes=fs, hidden_dim=H, rnn_class=rnn)
    net = SampleRNN.from_config(cfg).to("cuda")
    print(fs, H, net.launch_info(B))
    for mode in ("f32", "bf16"):
        net.bfloat16() if mode == "bf16" else net.float()
        net.generate(prompts, 800)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record(); net.generate(prompts, n); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print(f"{rnn} {mode}: {B * n / ms * 1e3 / 1e6:.2f} M samples/s ({ms:.0f} ms for {B} x ({P} prompt + {n} generated))", flush=True)
