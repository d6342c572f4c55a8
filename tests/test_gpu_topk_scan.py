"""Threshold-scan form of the bf16 brute-force top-k (csrc/topk_scan.cu) against the oracle's IndexFlatIP / tf.math.top_k
restatement (oracle.brute_force_topk): sample -> guaranteed per-row threshold -> one streaming pass that
only keeps survivors -> per-row sort -> exact re-rank.  Ids and scores must be bit-exact in every mode, including the
device-side fallback to the list-keeping kernel when a row's survivor buffer overflows.
"""
import numpy as np
import pytest
import torch

import oracle
from two_tower_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops(tt):
    return tt.ops


@pytest.fixture()
def scan_mode(tt):
    lib = tt._lib.load()
    prev = lib.tt_debug_topk_scan_mode(-1)
    yield lambda m: lib.tt_debug_topk_scan_mode(m)
    lib.tt_debug_topk_scan_mode(prev)


def bf(x):
    return torch.as_tensor(np.ascontiguousarray(x)).cuda().to(torch.bfloat16).contiguous()


def gaussian(rng, n, d):
    return oracle.bf16_round((rng.normal(size=(n, d)) / np.sqrt(d)).astype(np.float32))


def run(ops, q, c, k, **kw):
    unc = torch.zeros(1, dtype=torch.int32, device="cuda")
    s, i = ops.topk_bruteforce("bf16", bf(q), bf(c), k, uncertain=unc, **kw)
    torch.cuda.synchronize()
    return s.cpu().numpy(), i.cpu().numpy(), int(unc.item())


class TestThresholdScan:
    @pytest.mark.parametrize("nq,nc,d,k", [(64, 300_000, 128, 100), (1, 70_001, 128, 100), (129, 131_072, 64, 10),
                                           (300, 99_999, 128, 1), (1000, 262_144 + 77, 128, 100), (40, 500_000, 64, 50)])
    def test_gaussian_ids_and_scores_bit_exact(self, tt, ops, scan_mode, nq, nc, d, k):
        assert tt._lib.load().tt_topk_num_launches(1, nq, nc, d, k) >= 6      # this shape takes the scan path
        rng = synth.rng_for(91 + nq + k)
        q, c = gaussian(rng, nq, d), gaussian(rng, nc, d)
        ref_s, ref_i = oracle.brute_force_topk(q, c, k, score_dtype=np.float32, block=32768)
        s, i, unc = run(ops, q, c, k)
        assert np.array_equal(i, ref_i) and np.array_equal(s, ref_s) and unc == 0
        scan_mode(0)                                                            # the list-keeping kernel alone
        assert tt._lib.load().tt_topk_num_launches(1, nq, nc, d, k) <= 3
        s0, i0, _ = run(ops, q, c, k)
        assert np.array_equal(i0, ref_i) and np.array_equal(s0, ref_s)

    def test_overflowing_rows_fall_back_on_the_device(self, tt, ops, scan_mode):
        """64-entry survivor buffers: every row overflows, the flag routes the batch through the list-keeping kernel."""
        rng = synth.rng_for(404)
        nq, nc, d, k = 200, 150_000, 128, 100
        q, c = gaussian(rng, nq, d), gaussian(rng, nc, d)
        ref_s, ref_i = oracle.brute_force_topk(q, c, k, score_dtype=np.float32, block=32768)
        scan_mode(2)
        s, i, _ = run(ops, q, c, k)
        assert np.array_equal(i, ref_i) and np.array_equal(s, ref_s)

    def test_massive_ties(self, ops, scan_mode):
        """Dyadic data: thousands of candidates tie at the threshold score.  Whatever path ends up producing the pool,
        the ids follow the (score desc, index asc) rule exactly."""
        rng = synth.rng_for(405)
        nq, nc, d, k = 100, 120_000, 64, 100
        q, c = synth.exact_matrix(rng, nq, d, 2), synth.exact_matrix(rng, nc, d, 2)
        ref_s, ref_i = oracle.brute_force_topk(q, c, k)
        s, i, _ = run(ops, q, c, k)
        assert np.array_equal(i, ref_i) and np.array_equal(s.astype(np.float64), ref_s)

    def test_sorted_candidates_identifiers_and_base(self, ops, scan_mode):
        """Adversarial order for a strided sample (candidates sorted by their score against query 0), identifiers and
        cand_index_base applied after the re-rank."""
        rng = synth.rng_for(406)
        nq, nc, d, k = 32, 200_000, 128, 100
        q, c = gaussian(rng, nq, d), gaussian(rng, nc, d)
        c = c[np.argsort(c @ q[0])]
        ident = rng.permutation(nc).astype(np.int64) + 10_000_000_000
        ref_s, ref_i = oracle.brute_force_topk(q, c, k, score_dtype=np.float32, block=32768)
        s, i, _ = run(ops, q, c, k, identifiers=torch.as_tensor(ident).cuda())
        assert np.array_equal(i, ident[ref_i]) and np.array_equal(s, ref_s)
        s, i, _ = run(ops, q, c, k, cand_index_base=12345)
        assert np.array_equal(i, ref_i + 12345)

    def test_repeated_calls_reuse_the_workspace(self, ops):
        """Counters and the fallback flag are reset on the device by every call (graph-replay safe)."""
        rng = synth.rng_for(407)
        q, c = gaussian(rng, 64, 128), gaussian(rng, 100_000, 128)
        _, ref_i = oracle.brute_force_topk(q, c, 100, score_dtype=np.float32, block=32768)
        for _ in range(3):
            _, i, _ = run(ops, q, c, 100)
            assert np.array_equal(i, ref_i)
