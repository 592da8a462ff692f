"""GPU: the tcgen05 GEMM building block of the batched path against a plain PyTorch fp32 reference of the
same op (operands rounded to fp16 exactly as the kernel stores them, fp32 accumulation)."""
import ctypes as C

import pytest
import torch

from tacotron2_subword_b200 import _cabi

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("M,N,K,splits", [(128, 16, 64, 1), (128, 64, 256, 1), (256, 8, 512, 2), (512, 64, 1792, 2),
                                          (4096, 64, 4096, 4), (1024, 128, 1024, 1), (384, 33, 768, 3)])
def test_tcgen05_gemm_matches_torch(M, N, K, splits):
    lib = _cabi.load_library()
    g = torch.Generator(device="cpu").manual_seed(M + N + K)
    A = (torch.randn(M, K, generator=g) * 0.05).cuda()
    X = torch.randn(N, K, generator=g).cuda()
    out = torch.full((M, N), float("nan"), device="cuda")
    rc = lib.taco2dec_test_gemm(M, N, K, splits, C.c_void_p(A.data_ptr()), C.c_void_p(X.data_ptr()),
                                C.c_void_p(out.data_ptr()), C.c_void_p(torch.cuda.current_stream().cuda_stream))
    _cabi.check(rc)
    torch.backends.cuda.matmul.allow_tf32 = False
    ref = A.half().float() @ X.half().float().t()
    err = float((out - ref).abs().max())
    scale = float(ref.abs().max())
    assert torch.isfinite(out).all()
    assert err <= 2e-4 * max(1.0, scale), (err, scale)   # fp32 accumulation-order noise only
