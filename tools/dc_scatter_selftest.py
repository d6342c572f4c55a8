"""Single-GPU emulation of the dC pass that scatters its row blocks through tensor maps (tt_peer_retrieval_bwd_dc):
both "ranks" live on this GPU (the owner maps point at two local receive areas).  Checks the slots against the
combine path and replays the launch inside a CUDA graph.   python tools/dc_scatter_selftest.py"""
import sys
from pathlib import Path
from types import SimpleNamespace

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import two_tower_b200 as tt  # noqa: E402

ops = tt.ops
world, rank, b, d = 2, 1, 9600, 128
nq, nc = b, world * b
g = torch.Generator(device="cuda"); g.manual_seed(3)
q = (torch.randn((nq, d), device="cuda", generator=g) * 0.3).to(torch.bfloat16)
c = (torch.randn((nc, d), device="cuda", generator=g) * 0.3).to(torch.bfloat16)
loss, lse, _ = ops.retrieval_loss_fwd("bf16", q, c, 5.0, rank * b)
_none, dc_parts = ops.retrieval_loss_bwd_parts(q, c, 5.0, lse, rank * b, want_dq=False)
assert dc_parts.shape[0] == 1, dc_parts.shape
ref = dc_parts[0]
recv = [torch.zeros((world * b, d), dtype=torch.float32, device="cuda") for _ in range(world)]   # receive area of each "rank"
maps = ops.peer_row_maps([r.data_ptr() for r in recv], world * b, d, q.device)
ws = SimpleNamespace(world=world, rank=rank)
scratch = torch.empty(16, dtype=torch.float32, device="cuda")
ops.peer_retrieval_bwd_dc(ws, maps, q, c, 5.0, lse, rank * b, None, scratch)
torch.cuda.synchronize()
for o in range(world):
    got = recv[o][rank * b:(rank + 1) * b]
    want = ref[o * b:(o + 1) * b]
    print(f"owner {o}: max |diff| {float((got - want).abs().max()):.3e}  untouched slot zero: {bool((recv[o][(1 - rank) * b:(2 - rank) * b] == 0).all())}")
    assert torch.equal(got, want)
side = torch.cuda.Stream()
side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    ops.peer_retrieval_bwd_dc(ws, maps, q, c, 5.0, lse, rank * b, None, scratch)
torch.cuda.current_stream().wait_stream(side)
graph = torch.cuda.CUDAGraph()
with torch.cuda.graph(graph):
    ops.peer_retrieval_bwd_dc(ws, maps, q, c, 5.0, lse, rank * b, None, scratch)
for r in recv:
    r.zero_()
for _ in range(300):
    graph.replay()
torch.cuda.synchronize()
for o in range(world):
    assert torch.equal(recv[o][rank * b:(rank + 1) * b], ref[o * b:(o + 1) * b])
print("graph replay x300 ok")
