"""A/B of the Jacobi paths on hard matrices (GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from basd_b200._native import call, ptr, stream

torch.manual_seed(0)
dev = "cuda"
for n, decades, zero_rows in [(196, 2, 0), (196, 4, 0), (196, 4, 1), (196, 6, 3), (64, 4, 0)]:
    batch = 8
    u, _ = torch.linalg.qr(torch.randn(batch, n, n, dtype=torch.float64))
    v, _ = torch.linalg.qr(torch.randn(batch, n, n, dtype=torch.float64))
    s = torch.logspace(0, -decades, n, dtype=torch.float64)
    g0 = (u * s) @ v.transpose(1, 2)
    if zero_rows:
        g0[:, -zero_rows:, :] = 0
    g = g0.float().to(dev).contiguous()
    ref = torch.linalg.svdvals(g.double())
    work = g.clone()
    sweeps = torch.zeros(batch, dtype=torch.int32, device=dev)
    call("basd_jacobi_rows", ptr(work), n, n, n, n * n, batch, None, 1e-6, 18, ptr(sweeps), stream())
    torch.cuda.synchronize()
    w = work.double()
    gram0 = g.double().transpose(1, 2) @ g.double()
    gram1 = w.transpose(1, 2) @ w
    inv = float((gram1 - gram0).abs().max() / gram0.abs().max())
    rr = w @ w.transpose(1, 2)
    nrm = rr.diagonal(dim1=1, dim2=2).sqrt()
    cosm = rr / (nrm.unsqueeze(1) * nrm.unsqueeze(2)).clamp(min=1e-300)
    big = (nrm > 1e-6 * nrm.max(dim=1, keepdim=True).values)
    mask = big.unsqueeze(1) & big.unsqueeze(2) & ~torch.eye(n, dtype=torch.bool, device=dev)
    off = float((cosm.abs() * mask).max())
    sv = nrm.sort(dim=1, descending=True).values
    sverr = float(((sv - ref).abs().max(dim=1).values / ref[:, 0]).max())
    print(f"n={n} decades={decades} zero_rows={zero_rows} legacy={os.environ.get('BASD_JACOBI_LEGACY')}: "
          f"gram-invariance {inv:.2e} max off-cos {off:.2e} sv err {sverr:.2e} sweeps {sweeps.tolist()}")
