"""Determinism stress of the fused step kernels: the same launch repeated N times on the same inputs, every output
compared BITWISE with the first launch's.  Output buffers are poisoned (NaN) before they go back to the caching
allocator, so an element a kernel fails to write -- or reads before writing -- shows up at once.
   python tools/stress_determinism.py [iters] [B]"""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import two_tower_b200 as tt  # noqa: E402
from two_tower_b200 import synth  # noqa: E402

ops = tt.ops
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
g = torch.Generator(device="cuda"); g.manual_seed(1)
d_in, d_hid, d_out = 128, 256, 128


def tower(V, B, zipf):
    table = torch.rand((V, d_in), device="cuda", generator=g) * 0.1 - 0.05
    rng = synth.rng_for(V + B)
    ids = torch.from_numpy(synth.draw_ids(rng, B, V, 1.3)).cuda() if zipf else torch.randint(0, V, (B,), device="cuda", generator=g)
    w1 = (torch.randn((d_in, d_hid), device="cuda", generator=g) * 0.1).to(torch.bfloat16)
    w2 = (torch.randn((d_hid, d_out), device="cuda", generator=g) * 0.1).to(torch.bfloat16)
    return dict(features=[(table, ids, None, "sum")], batch=B, w1=w1, b1=torch.randn(d_hid, device="cuda", generator=g) * 0.1,
                w2=w2, b2=torch.randn(d_out, device="cuda", generator=g) * 0.1)


def report(name, it, ref, got):
    bad = (ref.view(torch.int32) if ref.dtype == torch.float32 else ref.view(torch.int16)) != \
          (got.view(torch.int32) if got.dtype == torch.float32 else got.view(torch.int16))
    rows = bad.reshape(bad.shape[0], -1).any(1).nonzero().flatten().cpu().numpy() if bad.dim() > 1 else bad.nonzero().flatten().cpu().numpy()
    diff = (ref.float() - got.float()).abs()
    print(f"  MISMATCH {name} at launch {it}: {int(bad.sum())} elements, {len(rows)} rows (first {rows[:6]}, last {rows[-3:]}), "
          f"nan {int(torch.isnan(got.float()).sum())}, max |diff| {float(torch.nan_to_num(diff, nan=-1).max()):.3e}", flush=True)


def poison(ts):
    for t in ts:
        if t.dtype in (torch.float32, torch.bfloat16):
            t.fill_(float("nan"))


def stress(B, zipf):
    specs = [tower(3000, B, zipf), tower(2500, B, zipf)]
    parts = [torch.randn((2, B, d_out), device="cuda", generator=g) * 0.01 for _ in specs]
    q = (torch.randn((B, d_out), device="cuda", generator=g) * 0.09).to(torch.bfloat16)
    c = (torch.randn((B, d_out), device="cuda", generator=g) * 0.09).to(torch.bfloat16)
    ref = None
    fails = {}
    for it in range(iters):
        outs = ops.tower_mlp2_fwd(specs)
        bw = ops.tower_mlp2_bwd([dict(s, x=x, h=h, dy_parts=p, dy_splits=2) for s, (x, h, y), p in zip(specs, outs, parts)])
        cur = {}
        for t, ((x, h, y), o) in enumerate(zip(outs, bw)):
            cur.update({f"fwd{t}.x": x, f"fwd{t}.h": h, f"fwd{t}.y": y, f"bwd{t}.dx": o["dx"], f"bwd{t}.dw1": o["dw1"],
                        f"bwd{t}.dw2": o["dw2"], f"bwd{t}.db1": o["db1"], f"bwd{t}.db2": o["db2"]})
        if ops.retrieval_fwd_dq_supported(B, B, d_out):
            loss, lse, pos, dq, ws = ops.retrieval_loss_fwd_dq(q, c, 2.0, fork=True)
            dc_parts = ops.retrieval_loss_bwd_dc_fused(q, c, 2.0, ws)
            ops.join_side_work()
            cur.update({"loss.lse": lse, "loss.pos": pos, "loss.dq": dq, "loss.dc_parts": dc_parts, "loss.loss": loss})
        if ref is None:
            ref = {k: v.clone() for k, v in cur.items()}
        else:
            for k, v in cur.items():
                if not torch.equal(ref[k].view(torch.uint8), v.view(torch.uint8)):
                    fails[k] = fails.get(k, 0) + 1
                    if fails[k] <= 3:
                        report(k, it, ref[k], v)
        poison(cur.values())
        del outs, bw, cur
    torch.cuda.synchronize()
    print(f"B={B} zipf={zipf}: {iters} launches, mismatching launches per output: {fails if fails else 'none'}", flush=True)


for B in ([int(sys.argv[2])] if len(sys.argv) > 2 else [1000, 8192]):
    for zipf in (True, False):
        stress(B, zipf)
