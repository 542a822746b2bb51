#!/usr/bin/env python
"""Throughput of the DMMA fp64 GEMM against torch.matmul (cuBLAS) on config-5 shapes."""
import ctypes as C, json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lsqfit_b200 import _cabi
dev = torch.device("cuda", 0)
def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
res = []
for (tA, tB, M, N, K, what) in [(0, 0, 5000, 2001, 5000, "W.[G|delta]"), (1, 0, 2000, 2000, 7000, "J^T J"),
                                (0, 0, 2000, 7000, 7000, "D.C"), (0, 1, 2000, 2000, 7000, "(DC).D^T"),
                                (0, 0, 8192, 8192, 8192, "square 8192"), (0, 0, 5000, 64, 64, "Jacobi column update")]:
    A = torch.randn((K, M) if tA else (M, K), device=dev, dtype=torch.float64)
    B = torch.randn((N, K) if tB else (K, N), device=dev, dtype=torch.float64)
    Cm = torch.empty((M, N), device=dev, dtype=torch.float64)
    st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    f1 = lambda: _cabi.check(_cabi.lib.b200lm_dgemm(0, tA, tB, 1, M, N, K, 1.0, A.data_ptr(), 0, A.stride(0), B.data_ptr(), 0, B.stride(0), 0.0, Cm.data_ptr(), 0, Cm.stride(0), st))
    f2 = lambda: torch.matmul(A.T if tA else A, B.T if tB else B, out=Cm)
    t1, t2 = timed(f1), timed(f2)
    fl = 2.0 * M * N * K
    res.append(dict(what=what, M=M, N=N, K=K, transA=tA, transB=tB, ms=t1, tflops=fl / t1 / 1e9, cublas_ms=t2, cublas_tflops=fl / t2 / 1e9))
    print("%-22s %5d x %5d x %5d  ours %8.3f ms %6.2f TF/s | cuBLAS %8.3f ms %6.2f TF/s" % (what, M, N, K, t1, fl / t1 / 1e9, t2, fl / t2 / 1e9))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/dgemm_r01.json", "w"), indent=1)
