"""GPU probe for ncu: the extractor chain (Normalize -> RemoveDC, fused Normalize + mu-law) on 360 clips of 10 s."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mimikit_b200 import Compose, MuLawCompress, Normalize, RemoveDC
x = torch.rand((360, 220500), device="cuda") * 2 - 1
for _ in range(2):
    y = RemoveDC()(Normalize()(x))
    q = Compose(Normalize(), MuLawCompress())(x)
torch.cuda.synchronize()
print("ok", tuple(y.shape), tuple(q.shape))
