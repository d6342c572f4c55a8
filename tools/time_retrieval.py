"""Back-to-back timing of the bf16 retrieval kernels alone (no host gaps: N launches between two events),
with the SM clock sampled while they run.
   python tools/time_retrieval.py [B] [d]"""
import subprocess
import sys
import threading
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import two_tower_b200 as tt  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
NC = int(sys.argv[2]) if len(sys.argv) > 2 else B
d = 128
ops = tt.ops
g = torch.Generator(device="cuda"); g.manual_seed(1)
q = (torch.randn((B, d), device="cuda", generator=g) * 0.3).to(torch.bfloat16)
c = (torch.randn((NC, d), device="cuda", generator=g) * 0.3).to(torch.bfloat16)
loss, lse, _ = ops.retrieval_loss_fwd("bf16", q, c, 10.0)


def timed(fn, n):
    for _ in range(5):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


rows = []
stop = False
def pump():
    p = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.active", "--format=csv,noheader,nounits",
                          "-i", "0", "-lms", "50"], stdout=subprocess.PIPE, text=True)
    while not stop:
        ln = p.stdout.readline()
        if ln:
            rows.append(ln.strip())
    p.terminate()
th = threading.Thread(target=pump, daemon=True); th.start()
time.sleep(0.3)
n_idle = len(rows)
fl = 2.0 * B * NC * d
for n in (20, 2000):
    tf = timed(lambda: ops.retrieval_loss_fwd("bf16", q, c, 10.0), n)
    tb = timed(lambda: ops.retrieval_loss_bwd_parts(q, c, 10.0, lse), n)
    print(f"n={n:5d}: fwd {tf:7.2f} us ({fl / tf / 1e6:6.0f} TFLOP/s)   bwd dQ+dC {tb:7.2f} us ({2 * fl / tb / 1e6:6.0f} algorithmic, "
          f"{4 * fl / tb / 1e6:6.0f} executed TFLOP/s)")
stop = True
time.sleep(0.2)
print("clocks idle   :", rows[:n_idle][-3:])
print("clocks loaded :", rows[n_idle:][::max(1, len(rows[n_idle:]) // 12)])
