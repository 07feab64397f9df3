"""CPU study of the Procrustes Jacobi's sweep count (no GPU needed): fp32 emulation of the
odd-even one-sided row Jacobi that jacobi_oe8.cu runs, against candidate re-orderings /
block variants, on the G = F_q^T F_p matrices of real C2-shaped problems (random-init
backbone features -> selector weights -> mixed teacher -> weighted centring -> pivoted
Cholesky, all through oracle/kernel_model.py).

    python tests/tools/jacobi_sweep_study.py [batch=4] [features=backbone|spectral]

Prints, per variant: mean / max sweeps, rotations (or block visits), the relative error of
the nuclear norm and the orthogonality defect of the result.  Results are recorded in
DESIGN.md section 10 (next round).
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch

import basd_b200.synthetic as syn
from oracle import kernel_model as km
from tests import _cases as cs

TOL = 1e-6


def problems(batch, features):
    work = cs.workload("c2", batch)
    if features == "backbone":
        from basd_b200 import backbone_features as bf
        logits, targets, st, te, at = bf.workload_inputs("c2", work, seed=0, device="cpu")
    else:
        logits, targets, st, te, at = syn.make_inputs(work, seed=0)
    proj_s, proj_t, logt = cs.selector_state(work)
    layers = sorted(st)
    t_keys = sorted(te)
    t_stack = torch.stack([te[k].float() for k in t_keys])
    bsz, n_t = t_stack.shape[1], t_stack.shape[2]
    sel = km.selector_model([km.token_stats(st[l]) for l in layers], [km.token_stats(te[k]) for k in t_keys],
                            bsz * work.n_student, bsz * n_t, proj_s, proj_t, logt)
    rows = torch.stack([km.attn_rows(at[k].float(), work.has_cls) for k in t_keys])
    out = []
    for i, layer in enumerate(layers):
        aligned = km.mix_and_align(sel["weights"][i], t_stack, work.n_student)
        imp, _ = km.mix_importance(sel["weights"][i], rows, work.n_student)
        for b in range(bsz):
            w = imp[b]
            root = w.sqrt().unsqueeze(1)
            s_tok, t_tok = st[layer][b].float(), aligned[b].float()
            a = root * (s_tok - (w.unsqueeze(1) * s_tok).sum(0, keepdim=True))
            bm = root * (t_tok - (w.unsqueeze(1) * t_tok).sum(0, keepdim=True))
            f_s = km.pivoted_cholesky(a @ a.T)
            f_t = km.pivoted_cholesky(bm @ bm.T)
            out.append((f_s.numpy(), f_t.numpy()))
    return out


# ------------------------------------------------------------------ scalar odd-even (the kernel's ordering)
def oe_scalar(g, tol=TOL, max_sweeps=30):
    """Rows of g (n x m) orthogonalised by odd-even transposition sweeps; rows trade places
    after every pair visit (jacobi_oe8.cu).  Returns (rows, sweeps, rotations)."""
    g = g.astype(np.float32).copy()
    n = g.shape[0]
    nn = n + (n & 1)
    if nn != n:
        g = np.vstack([g, np.zeros((1, g.shape[1]), np.float32)])
    rot_total = 0
    for sweep in range(1, max_sweeps + 1):
        worst = 0.0
        zero_thr = 1e-14 * float((g * g).sum(1).max())
        for step in range(nn):
            lo = step & 1
            i = np.arange(lo, nn - 1, 2)
            x, y = g[i], g[i + 1]
            ga = (x * y).sum(1)
            nx, ny = (x * x).sum(1), (y * y).sum(1)
            gg, nxy = ga * ga, nx * ny
            rot = (gg > tol * tol * nxy) & (nx > zero_thr) & (ny > zero_thr)
            if rot.any():
                worst = max(worst, float((gg[rot] / nxy[rot]).max()))
            d = ny - nx
            root = np.sqrt(d * d + 4 * gg)
            t = 2 * np.abs(ga) / np.maximum(np.abs(d) + root, 1e-37)
            t = np.where((d < 0) != (ga < 0), -t, t)
            t = np.where(rot, t, 0).astype(np.float32)
            c = (1 / np.sqrt(1 + t * t)).astype(np.float32)
            s = c * t
            xn = c[:, None] * x - s[:, None] * y
            yn = s[:, None] * x + c[:, None] * y
            g[i], g[i + 1] = yn, xn                       # trade places
            rot_total += int(rot.sum())
        if worst < tol:
            break
    return g[:n] if nn == n else g, sweep, rot_total


# ------------------------------------------------------------------ block odd-even
def oe_block(g, bs=4, tol=TOL, max_sweeps=30, sort="near"):
    """Block one-sided Jacobi: blocks of bs rows, odd-even transposition at block level; a
    visit of blocks (I, J) forms the 2bs x 2bs Gram of their rows, diagonalises it (eigh,
    fp32 data / fp64 solve as the in-register solver would reach) and applies V^T to the
    rows; blocks trade places.  Returns (rows, sweeps, visits)."""
    g = g.astype(np.float32).copy()
    n, m = g.shape
    nb = -(-n // bs)
    nb += nb & 1
    pad = nb * bs - n
    if pad:
        g = np.vstack([g, np.zeros((pad, m), np.float32)])
    blocks = g.reshape(nb, bs, m)
    visits = 0
    for sweep in range(1, max_sweeps + 1):
        worst = 0.0
        for step in range(nb):
            lo = step & 1
            i = np.arange(lo, nb - 1, 2)
            r = np.concatenate([blocks[i], blocks[i + 1]], axis=1)          # (pairs, 2bs, m)
            gram = np.einsum("pim,pjm->pij", r, r).astype(np.float32)
            dg = np.sqrt(np.maximum(np.einsum("pii->pi", gram), 1e-37))
            cos2 = (gram / (dg[:, :, None] * dg[:, None, :])) ** 2
            live = (dg[:, :, None] * dg[:, None, :]) > 1e-14 * float(dg.max()) ** 2
            off = np.where(live & ~np.eye(2 * bs, dtype=bool)[None], cos2, 0.0)
            wmax = off.reshape(len(i), -1).max(1)
            todo = wmax > tol * tol
            worst = max(worst, float(wmax.max()))
            lam, v = np.linalg.eigh(gram.astype(np.float64))
            if sort == "desc":
                v = v[:, :, ::-1]                                            # descending eigenvalues
            elif sort == "near":
                # the eigenvector matrix closest to the identity (what an inner small-angle Jacobi
                # produces): column j goes to the position where it is largest, sign made positive
                from scipy.optimize import linear_sum_assignment
                v = v.copy()
                for q in range(len(i)):
                    rr, cc = linear_sum_assignment(-np.abs(v[q]))
                    vq = np.empty_like(v[q])
                    vq[:, rr] = v[q][:, cc] * np.sign(v[q][rr, cc])[None, :]
                    v[q] = vq
            v = v.astype(np.float32)
            rn = np.einsum("pji,pjm->pim", v, r).astype(np.float32)
            rn = np.where(todo[:, None, None], rn, r)
            blocks[i], blocks[i + 1] = rn[:, bs:], rn[:, :bs]                # trade places
            visits += int(todo.sum())
        if worst < tol:
            break
    return blocks.reshape(nb * bs, m)[: n + pad], sweep, visits


def quality(rows, g):
    sv = np.linalg.svd(g.astype(np.float64), compute_uv=False)
    nrm = np.sqrt((rows.astype(np.float64) ** 2).sum(1))
    keep = nrm > 1e-6 * nrm.max()
    u = rows[keep].astype(np.float64) / nrm[keep, None]
    gram = u @ u.T
    defect = np.abs(gram - np.eye(len(gram))).max()
    return abs(nrm.sum() - sv.sum()) / sv.sum(), defect


def main():
    batch = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    features = sys.argv[2] if len(sys.argv) > 2 else "backbone"
    t0 = time.time()
    probs = problems(batch, features)
    print(f"{len(probs)} problems ({features}, C2 shapes, batch {batch}) built in {time.time() - t0:.1f} s")
    variants = {
        "scalar q=student (runs now)": lambda fs, ft: oe_scalar(fs.T @ ft),
        "scalar q=teacher": lambda fs, ft: oe_scalar(ft.T @ fs),
        "block 4+4 q=student near-identity V": lambda fs, ft: oe_block(fs.T @ ft, 4),
        "block 4+4 q=student sorted V": lambda fs, ft: oe_block(fs.T @ ft, 4, sort="desc"),
        "block 8+8 q=student near-identity V": lambda fs, ft: oe_block(fs.T @ ft, 8),
        "block 2+2 q=student near-identity V": lambda fs, ft: oe_block(fs.T @ ft, 2),
    }
    for name, fn in variants.items():
        sw, cnt, err, dfc = [], [], [], []
        t0 = time.time()
        for fs, ft in probs:
            g = (fs.T @ ft) if "q=student" in name else (ft.T @ fs)
            rows, s, c = fn(fs, ft)
            e, d = quality(rows, g)
            sw.append(s); cnt.append(c); err.append(e); dfc.append(d)
        print(f"{name:38s} sweeps mean {np.mean(sw):5.2f} max {max(sw):2d}  visits/rotations mean {np.mean(cnt):9.0f}  "
              f"nuc rel err {max(err):.1e}  orth defect {max(dfc):.1e}  ({time.time() - t0:.0f} s)")


if __name__ == "__main__":
    main()
