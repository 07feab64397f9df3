"""A/B of the Jacobi step synchronisation (GPU box): block-wide barriers (default) against
the neighbour-only mbarrier variant (BASD_JACOBI_PAIRSYNC=1, jacobi_oe8.cu).

    python tools/pairsync_ab.py                 # parent: runs both modes in child processes, diffs
    compute-sanitizer --tool racecheck python tools/pairsync_ab.py child small     # race check

The sweep's arithmetic does not depend on the synchronisation, so the two modes must agree
BITWISE (rows, sweep counts); the parent prints the timing of each mode on the shapes of the
per-sample Procrustes launch (1,024 x 196 x 196), the k x k SVDs and small / odd sizes.
The launcher reads the switch once per process, hence the child processes."""
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

CASES = [  # (batch, n, m, decades)
    (1024, 196, 196, 3), (1024, 196, 196, 5), (48, 174, 174, 2), (64, 131, 196, 3), (8, 20, 36, 1),
    (256, 64, 64, 3), (32, 255, 256, 4),
]
SMALL = [(2, 20, 36, 1), (2, 196, 196, 3), (2, 131, 196, 3)]


def child(which, out_path):
    import torch
    from basd_b200._native import call, ptr, stream
    dev = "cuda"
    results = {}
    for batch, n, m, decades in (SMALL if which == "small" else CASES):
        gen = torch.Generator().manual_seed(n * 1000 + m)
        u, _ = torch.linalg.qr(torch.randn(n, n, generator=gen, dtype=torch.float64))
        v, _ = torch.linalg.qr(torch.randn(m, m, generator=gen, dtype=torch.float64))
        s = torch.logspace(0, -decades, min(n, m), dtype=torch.float64)
        base = (u[:, : len(s)] * s) @ v[: len(s)]
        g0 = (base.unsqueeze(0) * (1 + 0.01 * torch.randn(batch, 1, 1, generator=gen, dtype=torch.float64)))
        g0 = g0 + 1e-3 * s[0] * torch.randn(batch, n, m, generator=gen, dtype=torch.float64) * s.mean()
        g = g0.float().to(dev).contiguous()
        times = []
        for rep in range(4):
            work = g.clone()
            sweeps = torch.zeros(batch, dtype=torch.int32, device=dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            call("basd_jacobi_rows", ptr(work), n, m, m, n * m, batch, None, 1e-6, 18, ptr(sweeps), stream())
            e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
        results[(batch, n, m, decades)] = dict(rows=work.cpu(), sweeps=sweeps.cpu(), ms=min(times[1:]))
        print(f"  {which:5s} pairsync={os.environ.get('BASD_JACOBI_PAIRSYNC')} batch {batch} n {n} m {m}: "
              f"{min(times[1:]):.3f} ms, sweeps mean {float(sweeps.float().mean()):.2f}", flush=True)
    if out_path:
        torch.save(results, out_path)


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        child(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)
        return
    import torch
    outs = {}
    for mode in ("block", "pair"):
        env = dict(os.environ)
        env.pop("BASD_JACOBI_PAIRSYNC", None)
        if mode == "pair":
            env["BASD_JACOBI_PAIRSYNC"] = "1"
        path = f"/tmp/pairsync_{mode}.pt"
        subprocess.run([sys.executable, os.path.abspath(__file__), "child", "full", path], check=True, env=env,
                       timeout=900)
        outs[mode] = torch.load(path, weights_only=False)
    ok = True
    for key, blk in outs["block"].items():
        par = outs["pair"][key]
        same = torch.equal(blk["rows"], par["rows"]) and torch.equal(blk["sweeps"], par["sweeps"])
        ok &= same
        print(f"{key}: bitwise equal {same}; block {blk['ms']:.3f} ms, pair {par['ms']:.3f} ms "
              f"({blk['ms'] / par['ms']:.2f}x)")
    print("ALL EQUAL" if ok else "MISMATCH")
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
