"""Brute-force top-k timing (bf16 tcgen05 kernel):  python tools/time_topk.py [nq] [nc] [k]"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import two_tower_b200 as tt  # noqa: E402

nq = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
nc = int(sys.argv[2]) if len(sys.argv) > 2 else 2_000_000
k = int(sys.argv[3]) if len(sys.argv) > 3 else 100
d = 128
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev); g.manual_seed(5678)
cand = (torch.randn((nc, d), device=dev, generator=g) / d ** 0.5).to(torch.bfloat16)
q = (torch.randn((nq, d), device=dev, generator=g) / d ** 0.5).to(torch.bfloat16)
index = tt.layers.factorized_top_k.BruteForce(k=k, precision="bf16").index(cand)
for _ in range(2):
    s, i = index(q)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record()
reps = 3
for _ in range(reps):
    s, i = index(q)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
tf = 2.0 * nq * nc * d / (ms * 1e-3) / 1e12
print(f"nq={nq} nc={nc} k={k}: {ms:.2f} ms  {nq / (ms * 1e-3):.0f} queries/s  {tf:.0f} TFLOP/s  ({nc * d * 2 * (nq / 128 / 148 if nq >= 128 * 148 else 1) / (ms * 1e-3) / 1e9:.0f} GB/s of candidate stream per pass)")
# spot check against torch on a slice
ref = (q[:256].float() @ cand.float().T)
rs, ri = torch.topk(ref, k, dim=1)
print("max |score diff| on 256 queries:", float((rs - s[:256]).abs().max()), " id agreement:", float((ri == i[:256]).float().mean()))
# per-kernel times of one call (event pairs around every launch)
import ctypes
lib = tt._lib.load()
lib.tt_profile_enable(1)
for _ in range(3):
    index(q)
buf = ctypes.create_string_buffer(1 << 14)
lib.tt_profile_collect(buf, len(buf))
lib.tt_profile_enable(0)
for ln in buf.value.decode().splitlines():
    name, cnt, total = ln.split()
    print(f"   {name:28s} {1e3 * float(total) / int(cnt):10.1f} us per launch")
print("   splits:", lib.tt_topk_num_splits(1, nq, nc, d, k))
