#!/usr/bin/env python
"""bench.py -- train examples/s of the two-tower hot path (BASELINE.json metric) on N B200s.

    python bench.py --gpus 1 --steps 50 --warmup 5
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the TFRS-equivalent CPU restatement (oracle) timed on host cores

A step = gather (both towers) -> tower MLPs -> in-batch softmax loss fwd -> loss bwd -> MLP bwd ->
sparse Adagrad on the tables + dense Adagrad on the MLP, on one synthetic batch of cfg2
(1M users x 500K items, d=128, B=8192 per GPU, MLP 256-128, bf16 tensor-core compute, fp32 tables).
Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "train examples/s (fwd+bwd, B=8192, d=128)"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=500)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="cfg2", choices=["cfg1", "cfg2", "cfg3"])
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-graph", action="store_true", help="time eager launches instead of CUDA-graph replay")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-serving", action="store_true", help="skip the top-100 brute-force retrieval line (queries/s)")
    return ap.parse_args()


def workload_name(cfg, n_gpus):
    s = (f"{cfg.name}: ID-only two-tower, {cfg.v_user} users x {cfg.v_item} items, d={cfg.dim}, "
         f"B={cfg.batch}/GPU, MLP {'-'.join(map(str, cfg.mlp)) or 'none'}, in-batch softmax T={cfg.temperature}, Adagrad")
    if cfg.bags:
        s = s.replace("ID-only", "ID + multi-hot " + "/".join(cfg.bags))
    if n_gpus > 1:
        s += f"; tables row-sharded over {n_gpus} GPUs, candidates all-gathered (B_glob={cfg.batch * n_gpus})"
    return s


# --------------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples nvidia-smi during the timed region (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.proc = None
        self.thread = None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def __exit__(self, *exc):
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        return False

    def summary(self):
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for row in self.rows:
            f = [x.strip() for x in row.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx = max(mx, float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------- reference arm
def oracle_state(cfg, seed_offset=0):
    import oracle
    from two_tower_b200 import synth
    rng = synth.rng_for(cfg.seed + seed_offset)
    qs = oracle.TowerSpec([("user_id_encoded", "id", cfg.v_user, None)], cfg.dim, cfg.mlp)
    cs = oracle.TowerSpec([("item_id_encoded", "id", cfg.v_item, None)], cfg.dim, cfg.mlp)
    qp, cp = oracle.init_tower(qs, rng, np.float32), oracle.init_tower(cs, rng, np.float32)
    mk = lambda p: {"tables": {k: np.full(v.shape, 0.1, np.float32) for k, v in p["tables"].items()},
                    "kernels": [np.full(k.shape, 0.1, np.float32) for k in p["kernels"]],
                    "biases": [np.full(b.shape, 0.1, np.float32) for b in p["biases"]]}
    return qs, cs, qp, cp, mk(qp), mk(cp)


def time_oracle_steps(cfg, steps, warmup, budget_s=None):
    """TFRS-equivalent restatement (oracle, numpy fp32 with the host BLAS on all cores) on full cfg batches."""
    import oracle
    from two_tower_b200 import synth
    qs, cs, qp, cp, qsl, csl = oracle_state(cfg)
    times = []
    t_start = time.perf_counter()
    for i in range(warmup + steps):
        b = synth.make_batch(cfg, i)
        t0 = time.perf_counter()
        oracle.two_tower_train_step(qs, cs, qp, cp, qsl, csl, {"user_id_encoded": b["user_id_encoded"]},
                                    {"item_id_encoded": b["item_id_encoded"]}, temperature=cfg.temperature,
                                    lr=0.001, dtype=np.float32, inplace=True)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
        if budget_s is not None and time.perf_counter() - t_start > budget_s and len(times) >= 2:
            break
    return times


def run_reference(args):
    from two_tower_b200 import synth
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = synth.CONFIGS[args.config]
    cores = os.cpu_count() or 1
    steps = max(1, min(args.steps, 20))
    times = time_oracle_steps(cfg, steps, min(args.warmup, 2), budget_s=150)
    ms = 1e3 * sum(times) / len(times)
    value = cfg.batch / (ms / 1e3)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "examples/s", "n_gpus": args.gpus,
        "steps": len(times), "warmup": min(args.warmup, 2), "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(cfg, 1), "global_batch": cfg.batch,
                   "sample": "full cfg2 steps at B=8192 on the host cores (rank 0 only under torchrun); the N-GPU arm's global "
                             "batch is N x 8192 with in-batch negatives over all of it"},
        "cpu_baseline": {"value": value, "unit": "examples/s", "cores": cores, "kind": "port",
                         "sample": f"{len(times)} full {cfg.name} steps (B={cfg.batch}) of the numpy TFRS-equivalent restatement "
                                   "(oracle/; TensorFlow/TFRS are not installable here), host BLAS on all cores"},
        "e2e": {"value": value, "unit": "examples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# -------------------------------------------------------------------------------- our arm
def build_model(tt, cfg, world, rank, group):
    if world > 1:
        from two_tower_b200 import parallel
        # TT_EXCHANGE = peer (default: every exchange is a peer-memory kernel), nccl (P2P lookups + NCCL collectives),
        # a2a (NCCL all-to-all lookups)
        mode = os.environ.get("TT_EXCHANGE", "peer")
        return parallel.build_sharded_two_tower(cfg, group, lr=0.001, peer={"peer": "exchange", "nccl": True, "a2a": False}[mode])

    class TwoTower(tt.models.Model):
        def __init__(self):
            super().__init__()
            def tower(vocab):
                layers = [tt.layers.Embedding(vocab, cfg.dim)]
                for j, u in enumerate(cfg.mlp):
                    layers.append(tt.layers.Dense(u, "relu" if j < len(cfg.mlp) - 1 else None))
                return tt.Sequential(layers)
            self.user_model = tower(cfg.v_user)
            self.item_model = tower(cfg.v_item)
            self.task = tt.tasks.Retrieval(temperature=cfg.temperature)

        def compute_loss(self, features, training=False):
            return self.task(self.user_model(features["user_id_encoded"]), self.item_model(features["item_id_encoded"]))

    model = TwoTower()
    model.compile(optimizer=tt.optimizers.Adagrad(learning_rate=0.001))
    return model


def ncu_traffic(kernel, world):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` from the committed ncu --set full capture
    of this command (profiles/traffic.json, written by tools/summarize_ncu.py traffic); None if not captured."""
    p = ROOT / "profiles" / "traffic.json"
    if world != 1 or not p.exists():
        return None
    rec = json.loads(p.read_text()).get(kernel)
    return None if rec is None else rec["dram_bytes_per_launch"]


def algorithmic_flops(cfg, world):
    """K3+K4 = 3 * 2*b*B_glob*d per GPU; K2 = 3 * 2*b*sum(in*out) per tower."""
    b, d = cfg.batch, cfg.dim
    dims = [d, *cfg.mlp]
    mlp = sum(dims[i] * dims[i + 1] for i in range(len(cfg.mlp)))
    d_out = dims[-1]
    return 3 * 2 * b * (b * world) * d_out + 2 * 3 * 2 * b * mlp


def run_ours(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    group = None
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"          # keep stdout to the one JSON line
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        group = dist.group.WORLD

    import two_tower_b200 as tt
    from two_tower_b200 import ops, synth

    ops.device_check()
    tt.set_precision(args.precision)
    cfg = synth.CONFIGS[args.config]
    dev = torch.device("cuda", local_rank)
    model = build_model(tt, cfg, world, rank, group)

    # a pool of distinct synthetic batches: pinned host copies (e2e) and device-resident copies (kernel-only)
    n_pool = 8
    def to_t(b, pin):
        out = {}
        for k, v in b.items():
            if isinstance(v, tuple):
                out[k] = tuple(torch.from_numpy(a).pin_memory() if pin else torch.from_numpy(a).to(dev) for a in v)
            else:
                out[k] = torch.from_numpy(v).pin_memory() if pin else torch.from_numpy(v).to(dev)
        return out
    host_pool = [to_t(synth.make_batch(cfg, 1000 * rank + i), True) for i in range(n_pool)]
    dev_pool = [to_t(synth.make_batch(cfg, 1000 * rank + i), False) for i in range(n_pool)]
    h2d_bytes = sum(t.numel() * t.element_size() for v in host_pool[0].values() for t in (v if isinstance(v, tuple) else (v,)))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    model.test_step(dev_pool[0])                      # builds the Dense layers
    # N > 1: the step contains NCCL collectives with static shapes (capacity-padded all-to-all buckets); they are
    # captured into the CUDA graph together with the kernels
    use_graph = not args.no_graph
    if use_graph:
        step = model.make_graphed_train_step(dev_pool[0], warmup=max(3, args.warmup))
    else:
        step = model.train_step
    for i in range(max(3, args.warmup)):
        step(dev_pool[i % n_pool])
    barrier()

    # ---- kernel-side timed region: inputs already in HBM, EXACTLY K steps, CUDA events, max over ranks
    launches0 = ops.LAUNCHES
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    clocks = ClockSampler(local_rank)
    clocks.__enter__()                                # sampled from here to the end of the e2e region (all under load)
    if args.steps >= 20:                              # >= 0.6 s of replays: loaded clocks, and enough nvidia-smi samples
        t_pre = time.perf_counter()
        i = 0
        while time.perf_counter() - t_pre < 0.6 or i < args.warmup:
            step(dev_pool[i % n_pool])
            i += 1
            if i % 64 == 0:
                torch.cuda.synchronize()
    else:
        for i in range(args.warmup):
            step(dev_pool[i % n_pool])
    barrier()
    e0.record()
    for i in range(args.steps):
        out = step(dev_pool[i % n_pool])
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    gpu_launches = ops.LAUNCHES - launches0
    if world > 1:
        t = torch.tensor([ms_total], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_step = ms_total / args.steps
    value = cfg.batch * world / (ms_step / 1e3)

    # ---- end to end through the public API: pinned host ids -> H2D -> step -> loss D2H, every step.  The loss of
    # every step is copied to its own slot of a pinned host array on the step's stream (a training loop that logs each
    # loss without stalling the device); all K copies complete inside the timed region.
    loss_host_all = torch.empty(args.steps, dtype=torch.float32).pin_memory()
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        hb = host_pool[i % n_pool]
        if use_graph:
            res = step(hb)                            # copies into the graph's static inputs (H2D), replays
        else:
            res = step({k: (tuple(a.to(dev, non_blocking=True) for a in v) if isinstance(v, tuple) else v.to(dev, non_blocking=True))
                        for k, v in hb.items()})
        loss_host_all[i:i + 1].copy_(res["loss"], non_blocking=True)      # D2H read of the step's result
    barrier()
    e2e_s = time.perf_counter() - t0
    loss_host = float(loss_host_all[-1])
    assert bool(torch.isfinite(loss_host_all).all()), "a step produced a non-finite loss"
    clocks.__exit__(None, None, None)
    if world > 1:
        t = torch.tensor([e2e_s], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = cfg.batch * world * args.steps / e2e_s

    # ---- in-graph spans: first CTA in -> last CTA out of every step kernel inside one replayed step (tt_debug_timeline)
    in_graph = None
    if use_graph:
        I64MAX = np.iinfo(np.int64).max
        names = {0: "tower_mlp2_fwd_kernel", 1: "retrieval_fwd_dq_tc_kernel", 3: "retrieval_bwd_tc_kernel(dC)", 4: "tower_mlp2_bwd_kernel",
                 5: "optimizer_step_kernel(first to last block entry)", 15: "retrieval_dq_finalize_kernel(block entries)"}
        tl = torch.tensor([I64MAX, 0] * 16, dtype=torch.int64, device=dev)
        lib0 = tt._lib.load()
        for i in range(3):
            if i == 2:
                torch.cuda.synchronize()
                tl.copy_(torch.tensor([I64MAX, 0] * 16, dtype=torch.int64))
                torch.cuda.synchronize()
                tt._lib.check(lib0.tt_debug_timeline(tl.data_ptr()))
            step(dev_pool[i % n_pool])
        torch.cuda.synchronize()
        tt._lib.check(lib0.tt_debug_timeline(None))
        t = tl.cpu().numpy().reshape(16, 2)
        in_graph = {names[k]: round((int(t[k, 1]) - int(t[k, 0])) / 1e3, 2) for k in names if t[k, 1] > 0}
        barrier()

    # ---- per-kernel durations, live, with a CUDA event pair around every launch (eager pass)
    lib = tt._lib.load()
    lib.tt_profile_enable(1)
    prof_steps = min(args.steps, 20)
    for i in range(prof_steps):
        # keep the GPU busy while the host enqueues the step, so that every event pair brackets a kernel
        # that starts back to back with its predecessor (otherwise the interval also holds the host's
        # launch latency, several us per launch in eager mode)
        torch.cuda._sleep(4_000_000)
        model.train_step(dev_pool[i % n_pool])
    import ctypes
    buf = ctypes.create_string_buffer(1 << 16)
    lib.tt_profile_collect(buf, len(buf))
    lib.tt_profile_enable(0)
    kernels = {}
    for ln in buf.value.decode().splitlines():
        name, cnt, total = ln.split()
        kernels[name] = {"launches_per_step": int(cnt) / prof_steps, "us_per_launch": 1e3 * float(total) / int(cnt),
                         "us_per_step": 1e3 * float(total) / prof_steps}
    barrier()

    def finish():
        # graphs that captured NCCL collectives must be gone before the communicator is torn down; a rank that
        # still hangs in the teardown must not keep the box busy
        if world > 1:
            sys.stdout.flush()
            torch.cuda.synchronize()
            dist.barrier()
            os._exit(0)

    if rank != 0:
        finish()
        return

    peaks = {}
    pk = ROOT / "MEASURED_PEAKS.json"
    if pk.exists():
        peaks = json.loads(pk.read_text())
    peak_tf = peaks.get("bf16_tflops", 1590.0)
    peak_src = "measured (MEASURED_PEAKS.json bf16_tflops, burst: kernel timed alone)" if peaks else "fallback 1.59 PFLOP/s"
    # dominant kernel = the retrieval kernel with the largest share of the step.  bf16: the one-pass loss forward + dQ
    # (2 algorithmic GEMMs, none recomputed) or the dC pass (1 algorithmic GEMM + the S recompute)
    d_out = (cfg.mlp[-1] if cfg.mlp else cfg.dim)
    gemm = 2.0 * cfg.batch * (cfg.batch * world) * d_out
    cands = {"retrieval_fwd_dq_tc_kernel": (2 * gemm, 2 * gemm), "retrieval_bwd_tc_kernel": (gemm, 2 * gemm),
             "retrieval_fwd_tc_kernel": (gemm, gemm), "retrieval_kernel": (gemm, gemm)}
    present = [k for k in cands if k in kernels]
    roofline = None
    if present:
        dom = max(present, key=lambda k: kernels[k]["us_per_step"])
        flops_per_launch, executed = cands[dom]
        us = kernels[dom]["us_per_launch"]
        achieved = flops_per_launch / (us * 1e-6) / 1e12
        roofline = {"bound": "tensor", "kernel": dom, "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s",
                    "frac": achieved / peak_tf, "traffic": ncu_traffic(dom, world), "us_per_launch": us,
                    "algorithmic_flops_per_launch": flops_per_launch, "peak_source": peak_src,
                    "executed_flops_per_launch": executed}
    step_flops = algorithmic_flops(cfg, world)
    line = {
        "metric": METRIC, "value": value, "unit": "examples/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(3, args.warmup), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
        "config": {"workload": workload_name(cfg, world), "global_batch": cfg.batch * world,
                   "l2_policy": "inputs larger than L2: tables+Adagrad slots are 1.5 GB of random rows vs 126 MB L2; no flush",
                   "launch": "cuda graph replay" if use_graph else "eager",
                   "exchange": None if world == 1 else os.environ.get("TT_EXCHANGE", "peer") +
                   " (peer: every exchange is a kernel on the symmetric NVLink workspace, no NCCL in the step)"},
        "clocks": clocks.summary(),
        "e2e": {"value": e2e_value, "unit": "examples/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4,
                "last_loss": loss_host,
                "readback": "each step's loss is copied D2H (async, stream-ordered) into its own pinned slot; all inside the timed region"},
        "gpu_launches": gpu_launches,
        "logit_pairs_per_s": float(cfg.batch) * cfg.batch * world * world / (ms_step * 1e-3),
        "step_tflops": step_flops / (ms_step * 1e-3) / 1e12 / world,
        "step_frac_of_bf16_peak": step_flops / (ms_step * 1e-3) / 1e12 / world / peaks.get("bf16_tflops_sustained", 1384.0),
        "roofline": roofline,
        "kernels_us_per_step": {k: round(v["us_per_step"], 2) for k, v in sorted(kernels.items(), key=lambda kv: -kv[1]["us_per_step"])},
        "kernels_us_in_graph": in_graph,
    }
    if world == 1 and in_graph and "optimizer_step_kernel(first to last block entry)" in in_graph and not cfg.bags:
        # second roofline: the HBM-bound sparse scatter + row-wise Adagrad (SURVEY.md 8d K5 bytes: gradient rows read, table
        # + accumulator rows read and written, ids), timed inside the replayed step
        hbm = peaks.get("hbm_gbs", 6650.0)
        nnz = 2 * cfg.batch
        algo = nnz * cfg.dim * 4 + nnz * cfg.dim * 4 * 4 + nnz * 8
        us_o = in_graph["optimizer_step_kernel(first to last block entry)"]
        line["roofline_hbm"] = {"bound": "hbm", "kernel": "optimizer_step_kernel", "achieved": algo / (us_o * 1e-6) / 1e9, "peak": hbm,
                                "unit": "GB/s", "frac": algo / (us_o * 1e-6) / 1e9 / hbm, "traffic": ncu_traffic("optimizer_step_kernel", world),
                                "us_per_launch": us_o, "algorithmic_bytes_per_launch": algo,
                                "note": "upper bound on rows touched (all ids distinct); span = first to last block ENTRY inside the "
                                        "replayed graph (blocks are short), the eager event-pair time is in kernels_us_per_step",
                                "peak_source": "measured (MEASURED_PEAKS.json hbm_gbs)" if peaks else "fallback 6.65 TB/s"}
    if world == 1 and not args.no_cpu_baseline:
        times = time_oracle_steps(cfg, 8, 1, budget_s=25)
        cpu_ms = 1e3 * sum(times) / len(times)
        line["cpu_baseline"] = {"value": cfg.batch / (cpu_ms / 1e3), "unit": "examples/s", "cores": os.cpu_count() or 1,
                                "kind": "port",
                                "sample": f"{len(times)} full {cfg.name} steps (B={cfg.batch}) of the numpy TFRS-equivalent restatement, host BLAS on all cores"}
    else:
        line["cpu_baseline"] = None
    if not args.no_serving and world == 1:
        line["serving"] = serving_bench(tt, torch, dev, peaks)
    print(json.dumps(line))
    finish()


def serving_bench(tt, torch, dev, peaks, nq=16384, nc=10_000_000, d=128, k=100):
    """Brute-force top-100: a bounded slice of cfg5 -- the full 10 M candidate set (bf16, fp32 accumulate, what one GPU
    holds when the queries are sharded), 16384 of the 1 M queries."""
    g = torch.Generator(device=dev); g.manual_seed(5678)
    cand = (torch.randn((nc, d), device=dev, generator=g) / d ** 0.5).to(torch.bfloat16)
    q = (torch.randn((nq, d), device=dev, generator=g) / d ** 0.5).to(torch.bfloat16)
    index = tt.layers.factorized_top_k.BruteForce(k=k, precision="bf16").index(cand)
    for _ in range(2):
        index(q)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    reps = 3
    for _ in range(reps):
        index(q)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    tf = 2.0 * nq * nc * d / (ms * 1e-3) / 1e12
    return {"metric": "top-100 queries/s", "value": nq / (ms * 1e-3), "unit": "queries/s", "queries": nq, "candidates": nc,
            "ms": ms, "tflops": tf, "frac_of_bf16_peak": tf / peaks.get("bf16_tflops", 1590.0)}


if __name__ == "__main__":
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
