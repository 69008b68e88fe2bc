import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mimikit_b200 import MuLawCompress
x = torch.rand((3600, 220500), device="cuda") * 2 - 1
mu = MuLawCompress()
ref = None
for w in ("0", "1", "2"):
    os.environ["MMK_MULAW_WIDE"] = w
    ts = []
    for _ in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record(); q = mu(x); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ref = q if ref is None else ref
    print(f"wide={w}: best {min(ts):.3f} ms -> {12*x.numel()/min(ts)/1e6:.0f} GB/s; equal to wide=0: {torch.equal(q, ref)}")
