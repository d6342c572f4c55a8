"""Per-CTA phase timeline (globaltimer ns) of the fused tower kernels at cfg2 shape (2 towers x 8192 rows).
   python tools/trace_tower.py [--dedup]"""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import two_tower_b200 as tt  # noqa: E402

ops = tt.ops
lib = tt._lib.load()
g = torch.Generator(device="cuda"); g.manual_seed(1)
B, d_in, d_hid, d_out = 8192, 128, 256, 128


def tower(V):
    table = torch.rand((V, d_in), device="cuda", generator=g) * 0.1 - 0.05
    ids = torch.randint(0, V, (B,), device="cuda", generator=g)
    w1 = (torch.randn((d_in, d_hid), device="cuda", generator=g) * 0.1).to(torch.bfloat16)
    w2 = (torch.randn((d_hid, d_out), device="cuda", generator=g) * 0.1).to(torch.bfloat16)
    return dict(features=[(table, ids, None, "sum")], batch=B, w1=w1, b1=torch.zeros(d_hid, device="cuda"),
                w2=w2, b2=torch.zeros(d_out, device="cuda"))


specs = [tower(1_000_000), tower(500_000)]
parts = [torch.randn((2, B, d_out), device="cuda", generator=g) for _ in specs]
buf = torch.zeros(2 * 16 * 256, dtype=torch.int64, device="cuda")


def run():
    if "--dedup" in sys.argv:             # the forward's dedup warp fills a fresh sparse-optimizer workspace (off by default)
        for s in specs:
            s["prepare_ws"] = ops.SparseWorkspace(B, d_in, "cuda")
    outs = ops.tower_mlp2_fwd(specs)
    ops.tower_mlp2_bwd([dict(s, x=x, h=h, dy_parts=p, dy_splits=2) for s, (x, h, y), p in zip(specs, outs, parts)])


for _ in range(3):
    run()
tt._lib.check(lib.tt_debug_tower_trace(buf.data_ptr()))
run()
torch.cuda.synchronize()
tt._lib.check(lib.tt_debug_tower_trace(None))
tr = buf.cpu().numpy().reshape(2, 256, 16)
labels = {0: ["entry", "setup", "gathered", "mma1 half0", "h half0", "h half1", "mma2", "y staged", "exit", "dedup: ids", "dedup: done"],
          1: ["entry", "setup", "dy tile", "mma a", "dh staged", "dW2 out", "mma c", "dx out", "mma d", "dW1 out", "exit"]}
for k, name in enumerate(["forward", "backward"]):
    rec = tr[k][tr[k][:, 0] > 0]
    t0 = rec[:, 0].min()
    n = len(labels[k])
    print(f"#### tower {name}: {len(rec)} CTAs, ns since the first CTA entered (min / median / max over CTAs)")
    for i in range(n):
        col = rec[:, i] - t0
        print(f"   {labels[k][i]:10s} {col.min():7d} {int(np.median(col)):7d} {col.max():7d}")
