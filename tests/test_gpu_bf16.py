"""GPU parity, bf16 precision (tcgen05 tensor-core kernels) through the C-ABI against the
oracle.  Tolerance 2e-2 relative (north_star bf16 bound); on exact-arithmetic inputs (small
dyadic values: every product and partial sum representable) the fp32-accumulated results must
be EXACT, which pins descriptor/layout bugs that a loose tolerance would hide."""
import numpy as np
import pytest
import torch

import oracle
from two_tower_b200 import synth

pytestmark = pytest.mark.gpu
BF16_RTOL = 2e-2


def dev(x, dtype=None):
    t = torch.as_tensor(np.ascontiguousarray(x))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda().contiguous()


def bf(x):
    return dev(x, torch.bfloat16)


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


@pytest.fixture(scope="module")
def ops(tt):
    tt.ops.device_check()
    return tt.ops


class TestDenseTensorCore:
    @pytest.mark.parametrize("M,i,o", [(128, 64, 128), (256, 128, 256), (8192, 128, 256), (8192, 256, 128), (1000, 64, 48), (72, 320, 16)])
    def test_forward_exact_on_dyadic_inputs(self, ops, M, i, o):
        rng = synth.rng_for(M + i + o)
        x = synth.exact_matrix(rng, M, i, 4); k = synth.exact_matrix(rng, i, o, 4)
        b = synth.exact_matrix(rng, 1, o, 4)[0]
        y, y_t, y_f = ops.dense_fwd("bf16", bf(x), bf(k.T.copy()), dev(b), relu=False, want_t=True, want_f32=True)
        ref = x.astype(np.float64) @ k + b
        assert np.array_equal(y_f.cpu().numpy().astype(np.float64), ref)
        assert np.array_equal(y.float().cpu().numpy(), oracle.bf16_round(ref.astype(np.float32)))
        assert np.array_equal(y_t.float().cpu().numpy(), oracle.bf16_round(ref.astype(np.float32)).T)
        y2, _, _ = ops.dense_fwd("bf16", bf(x), bf(k.T.copy()), dev(b), relu=True)
        assert np.array_equal(y2.float().cpu().numpy(), oracle.bf16_round(np.maximum(ref, 0).astype(np.float32)))

    @pytest.mark.parametrize("M,i,o", [(256, 128, 256), (8192, 128, 256), (8192, 256, 128), (1000, 64, 48)])
    def test_backward_exact_on_dyadic_inputs(self, ops, M, i, o):
        rng = synth.rng_for(M * 3 + i + o)
        x = np.maximum(synth.exact_matrix(rng, M, i, 4), 0); k = synth.exact_matrix(rng, i, o, 2)
        dy = synth.exact_matrix(rng, M, o, 2)
        dx, dx_t, dx_f, dk, P, db = ops.dense_bwd("bf16", bf(dy), bf(dy.T.copy()), bf(x), bf(x.T.copy()), bf(k),
                                                  relu_mask_x=True, want_dx=True, want_dx_t=True, want_dx_f32=True)
        ref_dx = (dy.astype(np.float64) @ k.T) * (x > 0)
        assert np.array_equal(dx_f.cpu().numpy().astype(np.float64), ref_dx)
        assert np.array_equal(dx.float().cpu().numpy(), oracle.bf16_round(ref_dx.astype(np.float32)))
        assert np.array_equal(dx_t.float().cpu().numpy(), oracle.bf16_round(ref_dx.astype(np.float32)).T)
        assert P == dk.shape[0] and P >= 1
        assert np.array_equal(dk.sum(0).cpu().numpy().astype(np.float64), x.astype(np.float64).T @ dy)
        assert np.array_equal(db.cpu().numpy().astype(np.float64), dy.astype(np.float64).sum(0))

    def test_gaussian_within_bf16_tolerance(self, ops):
        rng = synth.rng_for(12)
        M, i, o = 4096, 128, 256
        x = rng.normal(size=(M, i)).astype(np.float32); k = oracle.glorot_uniform(rng, i, o); b = rng.normal(size=o).astype(np.float32)
        y, _, _ = ops.dense_fwd("bf16", bf(x), bf(k.T.copy()), dev(b), relu=True)
        ref = oracle.dense_forward(x.astype(np.float64), k, b, "relu")
        assert rel_err(y.float().cpu().numpy(), ref) < BF16_RTOL
