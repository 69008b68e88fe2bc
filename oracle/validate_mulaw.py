"""TEST INFRASTRUCTURE — exhaustive pin of oracle/c/oracle_feat.c:orc_log1pf against torch.log1p (CPU, fp32)
over every fp32 bit pattern in [0, hi] (default hi = 255.0 = mu*C for q=256, C=1), plus a random sweep of the
whole compress pipeline against the live reference when /root/reference is present.
Run: python -m oracle.validate_mulaw [hi]      (container only; ~1 min)"""
import ctypes
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))


def main(hi=255.0):
    lib = ctypes.CDLL(os.path.join(HERE, "_build", "liboracle.so"))
    top = int(np.float32(hi).view(np.uint32)) + 1
    chunk = 1 << 26
    bad = 0
    for lo in range(0, top, chunk):
        n = min(chunk, top - lo)
        bits = torch.arange(lo, lo + n, dtype=torch.int64).to(torch.int32)
        x = bits.view(torch.float32).contiguous()
        out = torch.empty_like(x)
        lib.orc_log1pf_arr(ctypes.c_void_p(x.data_ptr()), ctypes.c_void_p(out.data_ptr()), ctypes.c_int64(n))
        ref = torch.log1p(x)
        # compare bit patterns (so that -0/+0 and NaNs would count)
        bad += (ref.view(torch.int32) != out.view(torch.int32)).sum().item()
    print(f"log1pf: {top} fp32 values in [0, {hi}] checked against torch.log1p (CPU): {bad} mismatches")
    return bad


if __name__ == "__main__":
    sys.exit(1 if main(float(sys.argv[1]) if len(sys.argv) > 1 else 255.0) else 0)
