import sys, os; sys.path.insert(0, os.getcwd())
import torch, numpy as np, ctypes as C
from lsqfit_b200 import _cabi
from lsqfit_b200.dense import _LA
la = _LA(0)
for n in (512, 2000, 4096):
    a = torch.randn(n, n, dtype=torch.float64, device='cuda'); A = a @ a.T + n * torch.eye(n, dtype=torch.float64, device='cuda')
    L = la.empty(n, n); linv = la.empty((n + 63)//64, 64, 64); info = torch.zeros(1, dtype=torch.int32, device='cuda')
    for _ in range(2): ok = la.potrf(A, 0.0, L, linv, info)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        _cabi.check(_cabi.lib.b200lm_potrf(0, n, A.data_ptr(), A.stride(0), 0.0, L.data_ptr(), L.stride(0), linv.data_ptr(), info.data_ptr(), la.stream()))
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    Lr = torch.linalg.cholesky(A)
    err = float((torch.tril(L) - Lr).abs().max() / Lr.abs().max())
    print(dict(n=n, ms=round(ms, 3), tflops=round(n**3 / 3 / ms / 1e9, 3), ok=ok, err=err))
