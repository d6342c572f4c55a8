"""Software-pipeline timeline (clock64) of CTA (0,0) of the one-pass loss forward + dQ kernel at the cfg2 shape.
   python tools/trace_fwd_dq.py [B]"""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import two_tower_b200 as tt  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
T = 64
N1 = 16 * T + 16
ops = tt.ops
lib = tt._lib.load()
g = torch.Generator(device="cuda"); g.manual_seed(1)
q = (torch.randn((B, 128), device="cuda", generator=g) * 0.09).to(torch.bfloat16)
c = (torch.randn((B, 128), device="cuda", generator=g) * 0.09).to(torch.bfloat16)
buf = torch.zeros(3 * N1 + 3 * 4 * 256, dtype=torch.int64, device="cuda")
for _ in range(5):
    ops.retrieval_loss_fwd_dq(q, c, 10.0)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    ops.retrieval_loss_fwd_dq(q, c, 10.0)
e1.record(); torch.cuda.synchronize()
print(f"fwd+dQ (+ fold + loss sum) back to back: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us per call")
tt._lib.check(lib.tt_debug_trace_buffer(buf.data_ptr()))
ops.retrieval_loss_fwd_dq(q, c, 10.0)
torch.cuda.synchronize()
tt._lib.check(lib.tt_debug_trace_buffer(None))
allbuf = buf.cpu().numpy()
cta = allbuf[3 * N1:].reshape(3, 256, 4)[1]
rec = cta[cta[:, 0] > 0]
t0c = rec[:, 0].min()
print(f"{len(rec)} CTAs (globaltimer ns): entry spread {rec[:, 0].max() - t0c}; setup done min/median/max "
      f"{np.min(rec[:, 1] - t0c)}/{int(np.median(rec[:, 1] - t0c))}/{np.max(rec[:, 1] - t0c)}; exit min/median/max "
      f"{np.min(rec[:, 2] - t0c)}/{int(np.median(rec[:, 2] - t0c))}/{np.max(rec[:, 2] - t0c)}; CTA(0,0): entry {rec[0, 0] - t0c} exit {rec[0, 2] - t0c} on sm {rec[0, 3]}")
dur = rec[:, 2] - rec[:, 1]
print(f"per-CTA work span (setup done -> exit) min/median/max: {dur.min()}/{int(np.median(dur))}/{dur.max()} ns")
tr = allbuf[:3 * N1].reshape(3, N1)[1]
ev = tr[:16 * T].reshape(4, T, 4)
t0 = ev[ev > 0].min()
lab = {0: "WG0 (even slots) / WG2 (odd slots): wait_S S_ready exp_done P_stored", 1: "WG1 (even slots) / WG3 (odd slots)",
       2: "S issuer: start tile_landed s_free issued", 3: "O issuer: start p_full issued"}
for role in range(4):
    print("--", lab[role])
    for t in range(T):
        if ev[role, t].max() == 0:
            continue
        print(f"   slot {t:2d}: " + "  ".join(f"{int(x - t0):7d}" if x > 0 else "      -" for x in ev[role, t]))
# summary: per-tile period and softmax turn of warpgroup 0 (even slots of role 0 = tiles 0, 1, 2, ...)
rows = [t for t in range(0, T, 2) if ev[0, t].max() > 0]
if len(rows) > 3:
    lat = [int(ev[0, t, 3] - ev[0, t, 1]) for t in rows[1:-1]]
    wait = [int(ev[0, t, 1] - ev[0, t, 0]) for t in rows[1:-1]]
    per = [int(ev[0, rows[i + 1], 1] - ev[0, rows[i], 1]) for i in range(1, len(rows) - 2)]
    print(f"WG0: softmax turn S_ready->P_stored median {int(np.median(lat))} cyc; wait for S median {int(np.median(wait))}; "
          f"period per 128 x 128 tile median {int(np.median(per))}")
