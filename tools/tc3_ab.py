"""A/B of the two tc3 GEMM kernels (TMA-fed persistent vs register-staged) on the batched shapes of a C2
step: bitwise comparison of the outputs and CUDA-event timing.  python tools/tc3_ab.py [batch]"""
import os
import sys
import subprocess

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

SHAPES = [  # ta, tb, m, n, k   (per-sample products of the Procrustes stage at C2, and two selector shapes)
    (0, 1, 196, 196, 384), (0, 1, 196, 196, 768), (1, 0, 196, 196, 196), (0, 0, 196, 196, 196),
    (0, 1, 196, 196, 196), (1, 1, 196, 196, 196), (0, 0, 196, 384, 196), (0, 0, 196, 768, 196),
    (0, 0, 384, 384, 384), (1, 0, 384, 384, 384),
]


def run(batch):
    import torch
    import basd_b200._engine as eng
    dev = torch.device("cuda", 0)
    out = {}
    for (ta, tb, m, n, k) in SHAPES:
        b = batch if m < 300 else 12
        torch.manual_seed(m * 7 + n * 3 + k + ta * 2 + tb)
        a = torch.randn(b, *((k, m) if ta else (m, k)), device=dev)
        bb = torch.randn(b, *((n, k) if tb else (k, n)), device=dev)
        c = torch.full((b, m, n), float("nan"), device=dev)
        args = (ta, tb, m, n, k, a, a.shape[2], a[0].numel(), bb, bb.shape[2], bb[0].numel(), c, n, m * n, b)
        for _ in range(2):
            eng.sgemm(*args, tc=True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            eng.sgemm(*args, tc=True)
        e1.record()
        torch.cuda.synchronize()
        ref = (a.transpose(1, 2) if ta else a).double() @ (bb.transpose(1, 2) if tb else bb).double()
        err = float((c.double() - ref).abs().max() / ref.abs().max())
        out[(ta, tb, m, n, k)] = (e0.elapsed_time(e1) / 5, c.cpu(), err)
    # the selector backward product: bf16 tokens (A) x fp32 W, column shift in the epilogue, bf16 output
    m, n, k = 50176, 384, 384
    torch.manual_seed(5)
    a = (torch.randn(m, k, device=dev) + 0.5).bfloat16()
    w = torch.randn(k, n, device=dev) / k ** 0.5
    shift = (a.float().mean(0).double() @ w.double()).float()
    c = torch.full((m, n), float("nan"), device=dev, dtype=torch.bfloat16)
    args = (0, 0, m, n, k, a, k, 0, w, n, 0, c, n, 0, 1)
    for _ in range(2):
        assert eng.gemm_tc_ex(*args, alpha=2.0, col_sub=shift)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        eng.gemm_tc_ex(*args, alpha=2.0, col_sub=shift)
    e1.record()
    torch.cuda.synchronize()
    ref = 2.0 * ((a.double() - a.double().mean(0)) @ w.double())
    err = float((c.double() - ref).abs().max() / ref.abs().max())
    out[("bf16A", 0, m, n, k)] = (e0.elapsed_time(e1) / 5, c.float().cpu(), err)
    return out


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[2] == "child":
        import torch
        res = run(int(sys.argv[1]))
        torch.save({str(k): v for k, v in res.items()}, sys.argv[3])
        sys.exit(0)
    batch = sys.argv[1] if len(sys.argv) > 1 else "1024"
    paths = {}
    for name, env in (("tma", {}), ("staged", {"BASD_TC3_NO_TMA": "1"})):
        paths[name] = f"/tmp/tc3_ab_{name}.pt"
        e = dict(os.environ, **env)
        rc = subprocess.run([sys.executable, __file__, batch, "child", paths[name]], env=e, timeout=600).returncode
        print(name, "rc", rc, flush=True)
        if rc:
            sys.exit(rc)
    import torch
    x, y = torch.load(paths["tma"]), torch.load(paths["staged"])
    tot_x = tot_y = 0.0
    for key in x:
        tx, cx, ex = x[key]
        ty, cy, ey = y[key]
        same = torch.equal(cx, cy)
        tot_x += tx
        tot_y += ty
        print(f"{key}: tma {tx:.3f} ms  staged {ty:.3f} ms  bitwise_equal {same}  rel err {ex:.2e} / {ey:.2e}", flush=True)
    print(f"total: tma {tot_x:.3f} ms  staged {tot_y:.3f} ms")
