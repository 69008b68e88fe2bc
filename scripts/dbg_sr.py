import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
os.environ["MMK_SR_KERNEL"] = sys.argv[1] if len(sys.argv) > 1 else "2"
if len(sys.argv) > 2: os.environ["MMK_SR_CLUSTER"] = sys.argv[2]
from test_samplernn_gpu import make_net
from oracle import restate
fs, H, B, P = (8, 2, 1), 64, 70, 24
net = make_net(fs, H, mlp_dim=32, seed=5)
print(net.launch_info(B))
orc = restate.SampleRNNOracle({k: v.numpy() for k, v in net.state_dict().items()}, fs)
g = torch.Generator().manual_seed(17)
n = 37
prompts = torch.randint(0, 256, (B, P), generator=g)
noise = torch.rand(B, n, generator=g)
for temp in (None, 0.95):
    seq, logits = net.generate(prompts, n, temperature=temp, noise=noise, return_logits=True)
    ref_seq, ref_logits = orc.generate(prompts.numpy(), n, temp, noise.numpy())
    s = seq.cpu().numpy(); l = logits.cpu().numpy()
    bad = np.argwhere(s != ref_seq)
    print("temp", temp, "mismatches", len(bad), "first", bad[:5].tolist())
    if len(bad):
        b, t = bad[0]
        print("  got", s[b, t], "ref", ref_seq[b, t], "logit err at step", np.abs(l[b, t - P] - ref_logits[b, t - P]).max(),
              "max before", np.abs(l[b, :t - P] - ref_logits[b, :t - P]).max() if t > P else None)
        lg, dec = net.teacher_forced(torch.from_numpy(ref_seq), P, temp, noise)
        lg = lg.cpu().numpy(); dec = dec.cpu().numpy()
        print("  teacher-forced: decision mismatches", int((dec != ref_seq[:, P:]).sum()), "max logit err", np.abs(lg - ref_logits).max(), "rel", np.abs(lg - ref_logits).max() / np.abs(ref_logits).max())
        bb = np.argwhere(dec != ref_seq[:, P:])
        for b, k in bb[:5]:
            # CDF margin
            z = ref_logits[b, k].astype(np.float64) / temp
            pr = np.exp(z - z.max()); c = np.cumsum(pr); u = noise[b, k].item() * c[-1]
            print("   b", b, "k", k, "dec", dec[b, k], "ref", ref_seq[b, P + k], "cdf margin", np.abs(c - u).min() / c[-1])
