"""-m gpu: the contractions of the training step that are not part of the per-frame recurrence (csrc/wgrad.cuh), through the C-ABI.

taco2dec_wgrad_gemm   C[M, N] (+)= sum over (t, b) Y[t, b, :]^T X[t, b, :]   weight gradients (the sum PyTorch's autograd forms for
                                                                         every nn.Linear / LSTMCell of reference model.py:209-345)
taco2dec_sgemm_nn     C = A . B (optionally masked, x2)                  input gradients of the prenet / memory layers
taco2dec_bmm_tn       C[b] = A[b]^T . B[:, b]                            d(memory) from the attention weights

Bounds: wgrad operands are fp16 (Y pre-scaled by a power of two so that tiny gradients keep their mantissa), fp32 accumulation:
2e-3 of the result's scale; the two SIMT kernels are fp32 throughout: 1e-5."""
import ctypes as C

import pytest
import torch

from tacotron2_subword_b200 import _cabi

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib_handle():
    from tacotron2_subword_b200 import Decoder, create_hparams
    dec = Decoder(create_hparams()).cuda()
    eng = dec._engine(torch.device("cuda", 0))
    yield eng.lib, eng.handle
    del dec


def _p(t):
    return C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _wgrad(lib, h, Y, X, out, accumulate=False, reuse=False, ws=None):
    T, B, M = Y.shape
    N = X.shape[2]
    need = int(lib.taco2dec_wgrad_workspace_bytes(h, M, N, T, B))
    if ws is None:
        ws = torch.empty(need, dtype=torch.uint8, device="cuda")
    assert ws.numel() >= need
    _cabi.check(lib.taco2dec_wgrad_gemm(h, _p(Y), Y.stride(0), Y.stride(1), M, _p(X), X.stride(0), X.stride(1), N, T, B, _p(out),
                                        out.stride(0), int(accumulate), int(reuse), _p(ws), ws.numel(), _stream()))
    return ws


def _ref(Y, X):
    return torch.einsum("tbm,tbn->mn", Y.double(), X.double())


@pytest.mark.parametrize("T,B,M,N", [(5, 3, 1, 80), (7, 2, 128, 512), (33, 4, 200, 96), (200, 16, 1024, 256), (64, 8, 4096, 1024),
                                      (1, 1, 80, 1536)])
def test_wgrad_matches_float64_sum(lib_handle, T, B, M, N):
    lib, h = lib_handle
    g = torch.Generator(device="cuda").manual_seed(T * 31 + M)
    Y = torch.randn(T, B, M, device="cuda", generator=g) * 3e-6        # gradient-sized values: far below fp16's normal range
    X = torch.randn(T, B, N, device="cuda", generator=g)
    out = torch.full((M, N), float("nan"), device="cuda")
    _wgrad(lib, h, Y, X, out)
    want = _ref(Y, X)
    scale = float(want.abs().max())
    assert float((out.double() - want).abs().max()) <= 2e-3 * scale


def test_wgrad_strided_views_accumulate_and_operand_reuse(lib_handle):
    """Operands addressed in place: utterance-major storage seen as [T, B, .], a column block of a wider output, accumulation,
    and a second product that reuses the packed Y of the first."""
    lib, h = lib_handle
    g = torch.Generator(device="cuda").manual_seed(11)
    T, B, M, N1, N2 = 40, 6, 320, 100, 256
    Yst = torch.randn(B, T, M, device="cuda", generator=g) * 1e-3
    Y = Yst.transpose(0, 1)                                     # strides (M, T*M, 1)
    Xall = torch.randn(T + 1, 2, B, N1, device="cuda", generator=g)
    X1 = Xall[1:, 1]                                            # strides (2*B*N1, N1, 1)
    X2 = torch.randn(T, B, N2, device="cuda", generator=g)
    out = torch.randn(M, N1 + N2 + 7, device="cuda", generator=g)
    before = out.clone()
    need = max(int(lib.taco2dec_wgrad_workspace_bytes(h, M, n, T, B)) for n in (N1, N2))
    ws = torch.empty(need, dtype=torch.uint8, device="cuda")
    _wgrad(lib, h, Y, X1, out[:, :N1], ws=ws)
    _wgrad(lib, h, Y, X2, out[:, N1:N1 + N2], accumulate=True, reuse=True, ws=ws)
    w1, w2 = _ref(Y, X1), _ref(Y, X2)
    assert float((out[:, :N1].double() - w1).abs().max()) <= 2e-3 * float(w1.abs().max())
    assert float((out[:, N1:N1 + N2].double() - before[:, N1:N1 + N2].double() - w2).abs().max()) <= 2e-3 * float(w2.abs().max())
    assert torch.equal(out[:, N1 + N2:], before[:, N1 + N2:])   # columns outside the block are untouched


def test_wgrad_zero_gradient_rows_give_exact_zero(lib_handle):
    lib, h = lib_handle
    Y = torch.zeros(9, 2, 64, device="cuda")
    X = torch.randn(9, 2, 48, device="cuda")
    out = torch.ones(64, 48, device="cuda")
    _wgrad(lib, h, Y, X, out)
    assert float(out.abs().max()) == 0.0


def test_wgrad_rejects_small_workspace(lib_handle):
    lib, h = lib_handle
    Y = torch.zeros(4, 2, 64, device="cuda")
    X = torch.zeros(4, 2, 48, device="cuda")
    out = torch.zeros(64, 48, device="cuda")
    ws = torch.empty(1024, dtype=torch.uint8, device="cuda")
    rc = lib.taco2dec_wgrad_gemm(h, _p(Y), Y.stride(0), Y.stride(1), 64, _p(X), X.stride(0), X.stride(1), 48, 4, 2, _p(out), 48, 0, 0,
                                 _p(ws), ws.numel(), _stream())
    assert rc != 0 and b"workspace" in lib.taco2dec_last_error()


@pytest.mark.parametrize("R,N,K,masked,acc", [(1, 1, 1, False, False), (70, 130, 33, False, True), (513, 256, 256, True, False),
                                             (3200, 512, 128, False, True)])
def test_sgemm_nn(lib_handle, R, N, K, masked, acc):
    lib, _ = lib_handle
    g = torch.Generator(device="cuda").manual_seed(R + N)
    A = torch.randn(R, K, device="cuda", generator=g)
    Bm = torch.randn(K, N + 3, device="cuda", generator=g)[:, :N]             # ldb > N
    mask = torch.relu(torch.randn(R, N, device="cuda", generator=g)) if masked else None
    out = torch.randn(R, N, device="cuda", generator=g)
    want = A.double() @ Bm.double()
    if masked:
        want = want * 2.0 * (mask > 0)
    if acc:
        want = want + out.double()
    _cabi.check(lib.taco2dec_sgemm_nn(_p(A), K, _p(Bm), Bm.stride(0), _p(out), N, R, N, K, _p(mask) if masked else None,
                                      N if masked else 0, int(acc), _stream()))
    assert float((out.double() - want).abs().max()) <= 1e-5 * max(1.0, float(want.abs().max()))


@pytest.mark.parametrize("batch,M,N,T", [(1, 1, 1, 1), (3, 37, 512, 50), (16, 149, 512, 200)])
def test_bmm_tn(lib_handle, batch, M, N, T):
    lib, _ = lib_handle
    g = torch.Generator(device="cuda").manual_seed(batch + T)
    A = torch.rand(batch, T, M, device="cuda", generator=g)
    Bm = torch.randn(T, batch, N, device="cuda", generator=g)
    out = torch.full((batch, M, N), float("nan"), device="cuda")
    _cabi.check(lib.taco2dec_bmm_tn(_p(A), A.stride(0), A.stride(1), _p(Bm), Bm.stride(0), Bm.stride(1), _p(out), out.stride(0), N,
                                    batch, M, N, T, _stream()))
    want = torch.einsum("btm,tbn->bmn", A.double(), Bm.double())
    assert float((out.double() - want).abs().max()) <= 1e-5 * max(1.0, float(want.abs().max()))
