import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
from test_wavenet_gpu import make_net
blocks = (4, 3)
for B, n, cl in [(16, 160, "4"), (37, 60, "4"), (37, 60, "16"), (130, 40, "16")]:
    g = torch.Generator().manual_seed(21)
    res = {}
    for cs in ("2", cl):
        os.environ["MMK_TC_CLUSTER"] = cs
        net = make_net(blocks, 128, 128, 128, 128, seed=4).bfloat16()
        P = net.rf + 6
        if cs == "2":
            seq = torch.randint(0, 256, (B, P + n), generator=g)
        lg, _ = net.teacher_forced(seq, P)
        res[cs] = lg.cpu().numpy()
    d = np.abs(res[cl] - res["2"]).max(axis=(0, 2))
    bad = np.nonzero(d > 0)[0]
    print(f"B={B} n={n} cluster {cl}: first bad step {bad[0] if len(bad) else None} of {n}; units before = {(bad[0] + net.rf - 1) * ((B + 15) // 16) if len(bad) else None}")
