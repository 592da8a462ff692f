"""Timeline of the tcgen05 GEMM kernel (globaltimer stamps inside the kernel), for the shapes the tensor path launches.
usage: TACO2DEC_GEMM_STAMPS=1 python tools/gemm_stamps.py"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tacotron2_subword_b200 import _cabi
lib = _cabi.load_library()
for (M, N, K, splits) in [(128, 64, 1024, 4), (4096, 64, 4096, 4), (4096, 64, 1792, 2), (1792, 64, 4096, 4)]:
    A = torch.randn(M, K, device="cuda") * 0.05
    X = torch.randn(N, K, device="cuda")
    out = torch.empty(M, N, device="cuda")
    _cabi.check(lib.taco2dec_test_gemm(M, N, K, splits, C.c_void_p(A.data_ptr()), C.c_void_p(X.data_ptr()),
                                       C.c_void_p(out.data_ptr()), C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    ref = A.half().float() @ X.half().float().t()
    print(M, N, K, splits, "max err", float((out - ref).abs().max()), flush=True)
