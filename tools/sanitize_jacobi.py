"""Small Jacobi / Cholesky / tc3 GEMM launches for compute-sanitizer (racecheck / memcheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from basd_b200 import _engine as eng
from basd_b200._native import call, ptr, stream

torch.manual_seed(0)
dev = "cuda"
for n, m in [(20, 36), (196, 196), (131, 196)]:
    g = torch.randn(2, n, m, device=dev)
    sw = torch.zeros(2, dtype=torch.int32, device=dev)
    call("basd_jacobi_rows", ptr(g), n, m, m, n * m, 2, None, 1e-6, 4, ptr(sw), stream())
    torch.cuda.synchronize()
    print("jacobi", n, m, sw.tolist(), bool(torch.isfinite(g).all()))
a = torch.randn(2, 196, 260, device=dev)
k = (a @ a.transpose(1, 2)).contiguous()
lt = torch.empty(2, 196, 196, device=dev)
eng.pivoted_cholesky(k, lt, 1e-5)
torch.cuda.synchronize()
print("cholesky", float((lt.transpose(1, 2) @ lt - a @ a.transpose(1, 2)).abs().max()))
x = torch.randn(2, 196, 100, device=dev)
y = torch.randn(2, 100, 196, device=dev)
c = torch.empty(2, 196, 196, device=dev)
eng.sgemm(0, 0, 196, 196, 100, x, 100, 196 * 100, y, 196, 196 * 100, c, 196, 196 * 196, 2, tc=True)
torch.cuda.synchronize()
print("tc3", float((c - x @ y).abs().max()))
