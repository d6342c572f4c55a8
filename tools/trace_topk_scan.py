"""clock64() timeline of CTA (0,0) of the second-phase top-k scan kernel (256 resident queries per CTA):
   python tools/trace_topk_scan.py [nq] [nc]"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import two_tower_b200 as tt  # noqa: E402

nq = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
nc = int(sys.argv[2]) if len(sys.argv) > 2 else 4_000_000
d, k, T = 128, 100, 48
g = torch.Generator(device="cuda"); g.manual_seed(5678)
cand = (torch.randn((nc, d), device="cuda", generator=g) / d ** 0.5).to(torch.bfloat16)
q = (torch.randn((nq, d), device="cuda", generator=g) / d ** 0.5).to(torch.bfloat16)
index = tt.layers.factorized_top_k.BruteForce(k=k, precision="bf16").index(cand)
for _ in range(2):
    index(q)
lib = tt._lib.load()
buf = torch.zeros(3 * T * 4, dtype=torch.int64, device="cuda")
tt._lib.check(lib.tt_debug_topk_scan_trace(buf.data_ptr()))
index(q)
torch.cuda.synchronize()
tt._lib.check(lib.tt_debug_topk_scan_trace(None))
r = buf.cpu().numpy().reshape(3, T, 4)
t0 = r[2, 0, 0]
print("tile | MMA thread: wait, tile landed, buffer free, issued | selection 0: wait, ready, released, done | selection 1: ...")
for t in range(T):
    print(f"{t:4d} | " + " ".join(f"{int(x - t0):7d}" for x in r[2, t]) + " | " + " ".join(f"{int(x - t0):7d}" for x in r[0, t]) + " | " +
          " ".join(f"{int(x - t0):7d}" for x in r[1, t]))
print("steady state, cycles per tile (2 units):", (r[2, T - 1, 3] - r[2, T // 2, 3]) / (T - 1 - T // 2))
