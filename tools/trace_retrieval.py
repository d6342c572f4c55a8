"""Software-pipeline timeline of CTA (0,0) of the bf16 retrieval kernels (tt_debug_trace_buffer).
   python tools/trace_retrieval.py            # runs fwd + bwd at B=8192, d=128 on cuda:0, prints per-tile cycles"""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import two_tower_b200 as tt  # noqa: E402

T = 64
N1 = 16 * T + 16
ops = tt.ops
lib = tt._lib.load()
g = torch.Generator(device="cuda"); g.manual_seed(1)
q = (torch.randn((8192, 128), device="cuda", generator=g) * 0.3).to(torch.bfloat16)
c = (torch.randn((8192, 128), device="cuda", generator=g) * 0.3).to(torch.bfloat16)
buf = torch.zeros(3 * N1 + 3 * 4 * 256, dtype=torch.int64, device="cuda")
for _ in range(3):
    loss, lse, _ = ops.retrieval_loss_fwd("bf16", q, c, 10.0)
    ops.retrieval_loss_bwd_parts(q, c, 10.0, lse)
tt._lib.check(lib.tt_debug_trace_buffer(buf.data_ptr()))
loss, lse, _ = ops.retrieval_loss_fwd("bf16", q, c, 10.0)
ops.retrieval_loss_bwd_parts(q, c, 10.0, lse)
torch.cuda.synchronize()
tt._lib.check(lib.tt_debug_trace_buffer(None))
allbuf = buf.cpu().numpy()
tr = allbuf[:3 * N1].reshape(3, N1)
cta = allbuf[3 * N1:].reshape(3, 256, 4)
for k, nm in enumerate(["fwd", "dQ", "dC"]):
    rec = cta[k][cta[k][:, 0] > 0]
    t0 = rec[:, 0].min()
    print(f"#### {nm}: {len(rec)} CTAs; entry spread {rec[:,0].max()-t0} ns; setup done (min/median/max) {np.min(rec[:,1]-t0)}/{int(np.median(rec[:,1]-t0))}/{np.max(rec[:,1]-t0)} ns; "
          f"exit (min/median/max) {np.min(rec[:,2]-t0)}/{int(np.median(rec[:,2]-t0))}/{np.max(rec[:,2]-t0)} ns; CTA(0,0) exit {rec[0,2]-t0} ns on sm {rec[0,3]}")
    slow = np.argsort(-(rec[:, 2] - t0))[:6]
    print("     slowest CTAs (index, sm, exit ns):", [(int(i), int(rec[i,3]), int(rec[i,2]-t0)) for i in slow])
names = {0: ("fwd", ["wait_S", "S_ready", "loaded", "done"]), 1: ("dQ", ["wait_S", "S_ready", "exp_done", "stored"]),
         2: ("dC", ["wait_S", "S_ready", "exp_done", "stored"])}
for k in range(3):
    ev = tr[k, :16 * T].reshape(4, T, 4)
    t0 = ev[ev > 0].min()
    name, labels = names[k]
    xs = tr[k, 16 * T:16 * T + 8]
    print(f"#### {name} CTA(0,0) thread 0 milestones (cycles): entry, setup, loop done, [3], [4], [5], before last sync, exit:",
          [int(x - t0) if x > 0 else None for x in xs])
    print(f"==== {name}: cycles relative to the first stamp; softmax WG0 | WG1 rows then MMA thread, TMA thread")
    for role, rn in enumerate(["WG0", "WG1", "MMA", "TMA"]):
        print(f"-- {rn}")
        for t in range(T):
            if ev[role, t].max() == 0:
                continue
            print(f"   tile {t:2d}: " + "  ".join(f"{int(x - t0):7d}" if x > 0 else "      -" for x in ev[role, t]))
