import ctypes as C, os, sys, torch
sys.path.insert(0, "/root/repo")
from tacotron2_subword_b200 import _cabi
lib = _cabi.load_library()
fmt = os.environ.get("TACO2DEC_GEMM_FMT", "")
M, N, K, splits = 256, 64, 1024, 2
torch.manual_seed(0)
A = torch.randn(M, K, device="cuda") * 0.05; X = torch.randn(N, K, device="cuda") * 1e-3
out = torch.empty(M, N, device="cuda")
_cabi.check(lib.taco2dec_test_gemm(M, N, K, splits, C.c_void_p(A.data_ptr()), C.c_void_p(X.data_ptr()), C.c_void_p(out.data_ptr()), C.c_void_p(torch.cuda.current_stream().cuda_stream)))
exact = A.double() @ X.double().t()
ra = (A.bfloat16() if fmt == "bf16" else A.half()).double(); rx = (X.bfloat16() if fmt in ("bf16", "mixed") else X.half()).double()
model = ra @ rx.t()
print(fmt or "fp16", "vs rounded-operand model:", float((out.double() - model).abs().max() / model.abs().max()), " vs exact:", float((out.double() - exact).abs().max() / exact.abs().max()))
