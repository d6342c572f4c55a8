"""In-graph timeline of one cfg2 training step (tt_debug_timeline): when each kernel's first CTA starts its work (after
the PDL wait) and its last CTA ends inside the replayed CUDA graph, i.e. the real gaps and overlaps the per-kernel
event timings hide.  Kernels launched several times per step (peer pushes, barriers 1 and 3) show one merged span.
   python tools/step_timeline.py [--no-graph]"""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench  # noqa: E402
import two_tower_b200 as tt  # noqa: E402
from two_tower_b200 import synth  # noqa: E402

import os  # noqa: E402
import torch.distributed as dist  # noqa: E402

world, rank, local_rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local_rank)
group = None
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    group = dist.group.WORLD
tt.set_precision("bf16")
cfg = synth.CONFIGS["cfg2"]
dev = torch.device("cuda", local_rank)
model = bench.build_model(tt, cfg, world, rank, group)
pool = [{k: torch.from_numpy(v).to(dev) for k, v in synth.make_batch(cfg, 1000 * rank + i).items()} for i in range(4)]
model.test_step(pool[0])
step = model.train_step if "--no-graph" in sys.argv else model.make_graphed_train_step(pool[0], warmup=3)
for i in range(10):
    step(pool[i % 4])
lib = tt._lib.load()
I64MAX = np.iinfo(np.int64).max
names = ["tower fwd", "loss fwd", "loss bwd dQ", "loss bwd dC", "tower bwd", "optimizer step", "sparse prepare (side stream)",
         "combine partials (N>1: + scatter to owners)", "fold dense parts", "peer pushes (candidates+ids ... dense bucket)",
         "peer barrier 0", "peer sum dense+loss (+barrier 2)", "peer sum (no barrier)", "peer push rows", "peer barriers 1 .. 3",
         "loss fold (block entries)"]
rows = []
for rep in range(5):
    buf = torch.tensor([I64MAX, 0] * 16, dtype=torch.int64, device=dev)
    torch.cuda.synchronize()
    tt._lib.check(lib.tt_debug_timeline(buf.data_ptr()))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for i in range(3):                      # the third of three back-to-back replays is the one kept (min/max are re-armed by reading)
        if i == 2:
            torch.cuda.synchronize()
            buf.copy_(torch.tensor([I64MAX, 0] * 16, dtype=torch.int64))
            torch.cuda.synchronize()
            e0.record()
        step(pool[i % 4])
    e1.record()
    torch.cuda.synchronize()
    tt._lib.check(lib.tt_debug_timeline(None))
    rows.append((buf.cpu().numpy().reshape(16, 2), e0.elapsed_time(e1) * 1e3))
tl, us = rows[-1]
t0 = min(int(tl[k, 0]) for k in range(16) if tl[k, 1] > 0)
if rank != 0:
    torch.cuda.synchronize(); dist.barrier(); os._exit(0)
print(f"step (events around one replay): {us:.1f} us")
print(f"{'kernel':32s} {'first CTA in':>12s} {'last CTA out':>12s} {'span':>8s}   gap to previous end")
prev_end = None
for k in sorted([k for k in range(16) if tl[k, 1] > 0], key=lambda k: tl[k, 0]):
    if tl[k, 1] == 0:
        continue
    a, b = (int(tl[k, 0]) - t0) / 1e3, (int(tl[k, 1]) - t0) / 1e3
    gap = "" if prev_end is None or k == 6 else f"{a - prev_end:6.2f}"
    print(f"{names[k]:32s} {a:12.2f} {b:12.2f} {b - a:8.2f}   {gap}")
    if k != 6:
        prev_end = b
print("(optimizer step / sparse prepare / combine / fold: 'last CTA out' is the latest block ENTRY; their blocks are short)")
if world > 1:
    sys.stdout.flush(); torch.cuda.synchronize(); dist.barrier(); os._exit(0)
