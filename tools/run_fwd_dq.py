"""A few launches of the one-pass loss forward + dQ kernel at the cfg2 shape (ncu target)."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import two_tower_b200 as tt  # noqa: E402
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
g = torch.Generator(device="cuda"); g.manual_seed(1)
q = (torch.randn((B, 128), device="cuda", generator=g) * 0.09).to(torch.bfloat16)
c = (torch.randn((B, 128), device="cuda", generator=g) * 0.09).to(torch.bfloat16)
for _ in range(6):
    loss, lse, pos, dq, ws = tt.ops.retrieval_loss_fwd_dq(q, c, 10.0)
    tt.ops.retrieval_loss_bwd_parts(q, c, 10.0, lse, want_dq=False)
torch.cuda.synchronize()
print(float(loss.item()))
