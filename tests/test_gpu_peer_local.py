"""csrc/peer.cu on ONE GPU: the "ranks" are emulated by separate workspaces on the same device (every base pointer is
local), called one rank after the other -- same kernels, same address arithmetic as over NVLink.  Flag barriers are
exercised with world = 1 only (a multi-rank barrier on one GPU would wait for a kernel that cannot run)."""
from types import SimpleNamespace

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
W, B, D = 2, 384, 128


def _workspaces(tt, nbytes):
    bufs = [torch.zeros(nbytes, dtype=torch.uint8, device="cuda") for _ in range(W)]
    bases = torch.tensor([b.data_ptr() for b in bufs], dtype=torch.int64, device="cuda")
    return bufs, [tt.ops.PeerWorkspace(bufs[r], bases, W, r) for r in range(W)]


def test_push_combine_scatter_push_rows_pull_rows_and_slab_sum(tt):
    ops = tt.ops
    g = torch.Generator(device="cuda"); g.manual_seed(12)
    off_c, off_ids, off_dc, off_rows, off_bucket = 1024, 1024 + 512 * 1024, 2 * 1024 * 1024, 4 * 1024 * 1024, 8 * 1024 * 1024
    bufs, wss = _workspaces(tt, 9 * 1024 * 1024)
    view = lambda r, off, shape, dt: wss[r].view(off, shape, dt)

    # all-gather by producer-side writes: candidates (bf16) + ids, two segments in one launch per rank
    cand = [torch.randn((B, D), device="cuda", generator=g).to(torch.bfloat16) for _ in range(W)]
    ids = [torch.randint(0, 1000, (B,), device="cuda", generator=g) for _ in range(W)]
    for r in range(W):
        ops.peer_push(wss[r], [(cand[r], off_c + r * B * D * 2), (ids[r], off_ids + r * B * 8)])
    for r in range(W):
        assert torch.equal(view(r, off_c, (W * B, D), torch.bfloat16), torch.cat(cand))
        assert torch.equal(view(r, off_ids, (W * B,), torch.int64), torch.cat(ids))

    # reduce-scatter, producer side: ordered sum of 3 split partials, row i -> slot [rank] of owner i // B
    parts = [torch.randn((3, W * B, D), device="cuda", generator=g) for _ in range(W)]
    for r in range(W):
        ops.peer_combine_scatter(wss[r], parts[r], B, off_dc)
    for o in range(W):
        slots = view(o, off_dc, (W, B, D), torch.float32)
        for r in range(W):
            want = (parts[r][0] + parts[r][1]) + parts[r][2]
            assert torch.equal(slots[r], want[o * B:(o + 1) * B])
        # the owner folds its slots in rank order (what the backward tower kernel does with dy_splits = W)
        out = torch.empty((B, D), dtype=torch.float32, device="cuda")
        ops.peer_sum(wss[o], off_dc, out, slot=-1, local_stride=B * D * 4)
        assert torch.equal(out, slots[0] + slots[1])

    # gradient rows to the owners of their table rows (id % W), and the owner-side pull of the same rows
    rows = [torch.randn((B, D), device="cuda", generator=g) for _ in range(W)]
    for r in range(W):
        ops.peer_push_rows(wss[r], [(ids[r], rows[r], off_rows)], B, D)
    all_ids, all_rows = torch.cat(ids), torch.cat(rows)
    for o in range(W):
        got = view(o, off_rows, (W * B, D), torch.float32)
        mine = all_ids % W == o
        assert torch.equal(got[mine], all_rows[mine]) and bool((got[~mine] == 0).all())
    src_off = off_bucket                                          # each rank's own rows at the same offset (pull source)
    for r in range(W):
        view(r, src_off, (B, D), torch.float32).copy_(rows[r])
    for o in range(W):
        out = torch.zeros((W * B, D), dtype=torch.float32, device="cuda")
        ops.peer_pull_rows(wss[o], [(all_ids, src_off, out)], B, D, slot=-1)
        mine = all_ids % W == o
        assert torch.equal(out[mine], all_rows[mine]) and bool((out[~mine] == 0).all())


def test_barrier_epochs_advance_with_one_rank(tt):
    ops = tt.ops
    buf = torch.zeros(8192, dtype=torch.uint8, device="cuda")
    bases = torch.tensor([buf.data_ptr()], dtype=torch.int64, device="cuda")
    ws = ops.PeerWorkspace(buf, bases, 1, 0)
    for k in range(1, 4):
        ops.peer_barrier(ws, 2)
        torch.cuda.synchronize()
        assert int(ws.step[2]) == k + 1 and int(ws.step[8 + 2]) == 0          # epoch advanced once, block counter cleared
        assert int(buf.view(torch.int64)[2 * 16 + 0]) == k                    # flag slot 2, source rank 0
    out = torch.empty(1024, dtype=torch.float32, device="cuda")
    src = ws.view(1024, (1024,), torch.float32); src.copy_(torch.arange(1024, dtype=torch.float32))
    ops.peer_sum(ws, 1024, out, slot=2)                                       # barrier fused into a multi-block kernel
    torch.cuda.synchronize()
    assert torch.equal(out, src) and int(ws.step[2]) == 5 and int(ws.step[8 + 2]) == 0


def test_dc_pass_scattering_through_tensor_maps_matches_the_combine_path(tt):
    """tt_peer_retrieval_bwd_dc with both owners local: bit-identical slots, also under CUDA-graph replay."""
    ops = tt.ops
    world, rank, d = 2, 1, 128
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    b = ((sms // 2 + 2) // world + 1) * 128                    # world * b / 128 row blocks > SMs / 2: an unsplit dC pass
    nq, nc = b, world * b
    g = torch.Generator(device="cuda"); g.manual_seed(3)
    q = (torch.randn((nq, d), device="cuda", generator=g) * 0.3).to(torch.bfloat16)
    c = (torch.randn((nc, d), device="cuda", generator=g) * 0.3).to(torch.bfloat16)
    _loss, lse, _ = ops.retrieval_loss_fwd("bf16", q, c, 5.0, rank * b)
    _none, dc_parts = ops.retrieval_loss_bwd_parts(q, c, 5.0, lse, rank * b, want_dq=False)
    assert dc_parts.shape[0] == 1
    recv = [torch.zeros((world * b, d), dtype=torch.float32, device="cuda") for _ in range(world)]
    maps = ops.peer_row_maps([r.data_ptr() for r in recv], world * b, d, q.device)
    ws = SimpleNamespace(world=world, rank=rank)
    scratch = torch.empty(16, dtype=torch.float32, device="cuda")
    ops.peer_retrieval_bwd_dc(ws, maps, q, c, 5.0, lse, rank * b, None, scratch)
    for o in range(world):
        assert torch.equal(recv[o][rank * b:(rank + 1) * b], dc_parts[0][o * b:(o + 1) * b])
        assert bool((recv[o][(1 - rank) * b:(2 - rank) * b] == 0).all())
    graph = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        ops.peer_retrieval_bwd_dc(ws, maps, q, c, 5.0, lse, rank * b, None, scratch)
    torch.cuda.current_stream().wait_stream(side)
    with torch.cuda.graph(graph):
        ops.peer_retrieval_bwd_dc(ws, maps, q, c, 5.0, lse, rank * b, None, scratch)
    for r in recv:
        r.zero_()
    for _ in range(20):
        graph.replay()
    torch.cuda.synchronize()
    for o in range(world):
        assert torch.equal(recv[o][rank * b:(rank + 1) * b], dc_parts[0][o * b:(o + 1) * b])


def test_topk_partials_written_into_owner_receive_areas_and_merged(tt):
    """Candidate-sharded serving on one GPU: each emulated rank scores ALL queries against its shard; the re-rank epilogue
    writes the partial list of query qi into list slot [rank] of the receive area of rank qi // qpr; every owner then
    merges its W lists.  Result == the single-device search == the oracle (ids bit-exact, incl. cross-shard ties)."""
    import oracle
    from two_tower_b200 import synth
    ops = tt.ops
    d, k, qpr, per = 128, 100, 96, 6000 + 64
    nq = W * qpr
    rng = synth.rng_for(808)
    cand = oracle.bf16_round((rng.normal(size=(W * per, d)) / np.sqrt(d)).astype(np.float32))
    cand[per + 5] = cand[5]                                   # an exact duplicate in the other shard
    q = oracle.bf16_round((rng.normal(size=(nq, d)) / np.sqrt(d)).astype(np.float32))
    off_s = 1024
    off_i = off_s + W * qpr * k * 4
    bufs, wss = _workspaces(tt, off_i + W * qpr * k * 8)
    qd = torch.as_tensor(q).cuda().to(torch.bfloat16)
    unc = torch.zeros(1, dtype=torch.int32, device="cuda")
    for r in range(W):
        shard = torch.as_tensor(cand[r * per:(r + 1) * per]).cuda().to(torch.bfloat16)
        ops.topk_bruteforce_peer("bf16", qd, shard, k, r * per, wss[r], qpr, off_s, off_i, uncertain=unc)
    ref_s, ref_i = oracle.brute_force_topk(q, cand, k, score_dtype=np.float32, block=4096)
    for o in range(W):
        s, i = ops.topk_merge(wss[o].view(off_s, (W, qpr, k), torch.float32), wss[o].view(off_i, (W, qpr, k), torch.int64), k)
        assert np.array_equal(i.cpu().numpy(), ref_i[o * qpr:(o + 1) * qpr])
        assert np.array_equal(s.cpu().numpy(), ref_s[o * qpr:(o + 1) * qpr])
    assert int(unc.item()) == 0
    s1, i1 = ops.topk_bruteforce("bf16", qd, torch.as_tensor(cand).cuda().to(torch.bfloat16), k)
    assert np.array_equal(i1.cpu().numpy(), ref_i)
