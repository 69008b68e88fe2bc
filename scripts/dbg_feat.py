"""GPU-box probe: device time of the feature kernels alone at several sizes (events around each call)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mimikit_b200 import MagSpec, MelSpec, MuLawCompress, MuLawExpand

def timeit(fn, reps=5):
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record(); r = fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1)); del r
    return ts

L = 220500
ms_, mel, mu, ex = MagSpec(2048, 512), MelSpec(128), MuLawCompress(), MuLawExpand()
for clips in (36, 360, 1800, 3600):
    x = torch.rand((clips, L), device="cuda") * 2 - 1
    t = timeit(lambda: ms_.mel(x, mel))
    print(f"stft+mel clips={clips}: ms {['%.3f' % v for v in t]}  -> {4*clips*L/min(t)/1e6:.1f} GB/s in", flush=True)
    for mode in ("0", "1"):
        os.environ["MMK_MULAW_EXACT"] = mode
        t = timeit(lambda: mu(x))
        print(f"  mulaw compress exact={mode}: ms {['%.3f' % v for v in t]} -> {12*clips*L/min(t)/1e6:.1f} GB/s")
        t = timeit(lambda: mu.torch_func(x, out_dtype=torch.uint8))
        print(f"  mulaw compress u8 exact={mode}: ms {['%.3f' % v for v in t]} -> {5*clips*L/min(t)/1e6:.1f} GB/s")
        q = mu(x)
        t = timeit(lambda: ex(q))
        print(f"  mulaw expand exact={mode}: ms {['%.3f' % v for v in t]} -> {12*clips*L/min(t)/1e6:.1f} GB/s")
        del q
    os.environ.pop("MMK_MULAW_EXACT")
    del x

# Normalize and the fused Normalize -> mu-law pipeline
from mimikit_b200 import Compose, Normalize
x = torch.rand((3600, L), device="cuda") * 2 - 1
t = timeit(lambda: Normalize()(x))
print(f"normalize 3600 clips: ms {['%.3f' % v for v in t]} -> {12*3600*L/min(t)/1e6:.1f} GB/s (12 B/sample)")
t = timeit(lambda: Compose(Normalize(), MuLawCompress())(x))
print(f"fused normalize+mulaw: ms {['%.3f' % v for v in t]} -> {16*3600*L/min(t)/1e6:.1f} GB/s (16 B/sample)")
t = timeit(lambda: MuLawCompress()(Normalize()(x)))
print(f"unfused normalize, mulaw: ms {['%.3f' % v for v in t]}")
from mimikit_b200 import RemoveDC
t = timeit(lambda: RemoveDC()(x))
print(f"remove_dc 3600 clips: ms {['%.3f' % v for v in t]} -> {8*3600*L/min(t)/1e6:.1f} GB/s (8 B/sample)")
xx = torch.rand((36000, 22050), device="cuda") * 2 - 1
t = timeit(lambda: RemoveDC()(xx))
print(f"remove_dc 36000 x 1 s clips: ms {['%.3f' % v for v in t]} -> {8*36000*22050/min(t)/1e6:.1f} GB/s")
