"""GPU: the tcgen05 split-bf16 GEMM of the BATCH_NORM training path (csrc/gemm_tc.cu) against float64 matmul, at every
shape / transpose / leading-dimension combination bn_train.cu issues (trunk, skip, feature, ddir, heads; forward, dX, dW)."""
import numpy as np
import pytest
import torch

from tests.util import cuda

pytestmark = pytest.mark.gpu


def _gemm(ta, tb, M, N, K, A, lda, B, ldb, beta, C, ldc, precise=False):
    from nerf_keras_b200 import _lib
    from nerf_keras_b200.models import _ptr, _stream
    _lib.check(_lib.lib().nerf_selftest_gemm_f32(int(ta), int(tb), M, N, K, _ptr(A), lda, _ptr(B), ldb, float(beta), _ptr(C),
                                                 ldc, int(precise), _stream()), "gemm")
    torch.cuda.synchronize()


# (M rows of A / samples, N, K, B stored transposed, beta, extra leading-dimension padding of A / B / C)
ROWS_CASES = [
    (1000, 256, 63, False, 0.0, (0, 0, 0)),      # first trunk layer
    (777, 256, 256, False, 0.0, (0, 0, 0)),      # trunk layer, ragged last tile
    (512, 256, 63, False, 1.0, (0, 0, 0)),       # skip part, accumulating
    (640, 128, 256, False, 0.0, (0, 0, 0)),      # ddir
    (640, 128, 27, False, 1.0, (0, 0, 0)),       # ddir, direction part
    (900, 256, 128, True, 0.0, (0, 0, 0)),       # dX through ddir
    (300, 256, 256, True, 0.0, (0, 0, 0)),       # dX through a trunk layer
    (5000, 256, 256, False, 0.0, (3, 5, 1)),     # unaligned leading dimensions
    (300, 1, 256, False, 0.0, (0, 0, 3)),        # sigma head -> preds[:, 3]
    (300, 3, 128, False, 0.0, (0, 0, 1)),        # rgb head -> preds[:, :3]
    (300, 128, 3, True, 0.0, (1, 0, 0)),         # dX through the rgb head (d_preds has a stride of 4)
    (300, 256, 1, True, 1.0, (3, 0, 0)),         # += d sigma (x) W_sigma
    (20000, 256, 256, False, 0.0, (0, 0, 0)),    # more tiles than SMs
]


@pytest.mark.parametrize("precise", [False, True])
@pytest.mark.parametrize("M,N,K,tb,beta,pads", ROWS_CASES)
def test_rows_gemm(M, N, K, tb, beta, pads, precise):
    g = torch.Generator("cuda").manual_seed(M + N + K)
    lda, ldc = K + pads[0], N + pads[2]
    ldb = (K if tb else N) + pads[1]
    A = torch.randn(M, lda, device="cuda", generator=g)
    B = torch.randn(N if tb else K, ldb, device="cuda", generator=g) * 0.1
    C = torch.randn(M, ldc, device="cuda", generator=g)
    C0 = C.clone()
    _gemm(False, tb, M, N, K, A, lda, B, ldb, beta, C, ldc, precise)
    Bop = (B[:, :K].T if tb else B[:, :N]).double()
    ref = A[:, :K].double() @ Bop + beta * C0[:, :N].double()
    scale = float((A[:, :K].double().abs() @ Bop.abs()).max())
    err = float((C[:, :N].double() - ref).abs().max())
    # two-way split: 2^-16 per product; three-way: what is left is the tensor core's fp32 accumulation (measured 8e-7 of the
    # scale at K = 63, not correctly rounded fp32 adds) -- 25x tighter
    assert err <= (2e-6 if precise or N <= 4 or K <= 4 else 2e-5) * scale, (err, scale)
    if pads[2]:
        assert torch.equal(C[:, N:], C0[:, N:])          # padding columns of C untouched


TN_CASES = [
    (3000, 256, 256, (0, 0, 0)),       # trunk dW
    (1111, 63, 256, (0, 0, 0)),        # first layer / skip rows
    (4096, 256, 128, (0, 0, 0)),       # ddir dW, feature part
    (4096, 27, 128, (0, 0, 0)),        # ddir dW, direction part
    (2000, 128, 3, (0, 1, 0)),         # rgb head dW (d_preds stride 4)
    (2000, 256, 1, (0, 3, 0)),         # sigma head dW
    (40000, 256, 256, (2, 6, 4)),      # many tiles, unaligned leading dimensions
]


@pytest.mark.parametrize("Ms,Mo,N,pads", TN_CASES)
def test_transposed_gemm(Ms, Mo, N, pads):
    g = torch.Generator("cuda").manual_seed(Ms + Mo + N)
    lda, ldb, ldc = Mo + pads[0], N + pads[1], N + pads[2]
    A = torch.randn(Ms, lda, device="cuda", generator=g)
    B = torch.randn(Ms, ldb, device="cuda", generator=g) * 0.1
    C = torch.randn(Mo, ldc, device="cuda", generator=g)
    C0 = C.clone()
    _gemm(True, False, Mo, N, Ms, A, lda, B, ldb, 1.0, C, ldc)
    ref = A[:, :Mo].double().T @ B[:, :N].double() + C0[:, :N].double()
    scale = float((A[:, :Mo].double().abs().T @ B[:, :N].double().abs()).max())
    err = float((C[:, :N].double() - ref).abs().max())
    assert err <= 2e-5 * scale, (err, scale)
    if pads[2]:
        assert torch.equal(C[:, N:], C0[:, N:])


def test_gemm_rejects_unsupported():
    from nerf_keras_b200 import _lib
    from nerf_keras_b200.models import _ptr, _stream
    A = torch.zeros(128, 512, device="cuda")
    B = torch.zeros(512, 256, device="cuda")
    C = torch.zeros(128, 256, device="cuda")
    L = _lib.lib()
    assert L.nerf_selftest_gemm_f32(0, 0, 128, 256, 512, _ptr(A), 512, _ptr(B), 256, 0.0, _ptr(C), 256, 0, _stream()) != 0
    assert L.nerf_selftest_gemm_f32(1, 0, 256, 256, 128, _ptr(A), 512, _ptr(B), 256, 0.0, _ptr(C), 256, 0, _stream()) != 0
