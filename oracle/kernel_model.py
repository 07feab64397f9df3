"""CPU model of the *reformulated* algorithm the CUDA kernels implement (TEST INFRA ONLY).

``oracle/ref_port.py`` restates the reference (tall SVDs, 384x768 per-sample SVDs,
autograd).  The CUDA path computes the same numbers through cheaper, exact
reformulations: token-space Grams, small symmetric eigenproblems, an
N x N per-sample Procrustes problem and closed-form gradients.  This file states that
reformulated algorithm, stage by stage, in plain torch so that

  * each CUDA kernel has a stage-level expected output (tests/test_kernels_gpu.py), and
  * the reformulation itself is proven against the reference's autograd on CPU
    (tests/test_kernel_model.py) -- no GPU needed.

Nothing under the shipped package imports this file.
"""
from __future__ import annotations

import math

import torch

EPS32 = float(torch.finfo(torch.float32).eps)


# ------------------------------------------------------------------ stage: statistics
def token_stats(tokens: torch.Tensor):
    """Gram and column sum in *token space* (kernel: basd_gram_colsum).
    bf16 x bf16 products are exact in fp32, so this is exact up to accumulation order."""
    x = tokens.reshape(-1, tokens.shape[-1]).float()
    return x.T @ x, x.sum(dim=0)


def rotate_stats(gram, colsum, proj):
    """P G P^T and P c in double precision (kernel: basd_rotate_stats_f64) -- reference:
    layer_selector.py:72,88.  The fp32 version of this product is what limited the selector's
    gradient on ill-conditioned tokens (cosine 0.9989 -> 0.99999 at condition number 2e3)."""
    p64 = proj.double()
    return p64 @ gram.double() @ p64.T, p64 @ colsum.double()


def centred(g64, c64, rows):
    """fp32( sym(P G P^T) - (P c)(P c)^T / M ): one rounding per entry, as the kernel does."""
    return (0.5 * (g64 + g64.T) - torch.outer(c64, c64) / rows).float()


def sym_eig_desc(mat: torch.Tensor):
    """Symmetric eigendecomposition, eigenvalues descending (kernel: jacobi_eig)."""
    lam, vec = torch.linalg.eigh(mat.double())
    lam = lam.flip(0).float()
    vec = vec.flip(1).float()
    return lam, vec


def mp_rank_from_spectrum(lam_desc: torch.Tensor, rows: int, dim_cap: int) -> int:
    """MP rank from the spectrum of the uncentred second moment (kernel: mp_rank).
    reference: layer_selector.py:8-20,74. Scale-free, so the 1/M factor is dropped."""
    n = lam_desc.numel()
    asc = lam_desc.flip(0)
    median = asc[(n - 1) // 2]                      # torch.median: lower middle
    edge = median * (1.0 + math.sqrt(n / rows)) ** 2
    return min(int((lam_desc > edge).sum()), dim_cap)


# ------------------------------------------------------------------ stage: selector
def selector_model(student_stats, teacher_stats, rows_s, rows_t, proj_s, proj_t, log_temps):
    """student_stats/teacher_stats: lists of (gram, colsum) in token space.
    Returns dict with ranks, dist (E,L), weights (E,L) and what backward needs."""
    d_s = proj_s.shape[0]
    teach = []
    for gram, col in teacher_stats:
        g, c = rotate_stats(gram, col, proj_t)
        lam_u, _ = sym_eig_desc(g)
        k = mp_rank_from_spectrum(lam_u, rows_t, d_s - 1)
        lam, vec = sym_eig_desc(centred(g, c, rows_t))
        teach.append(dict(k=k, basis=vec[:, :k], sw=lam[:k].clamp(min=0).sqrt()))
    temps = torch.nn.functional.softplus(log_temps)
    out = dict(ranks=[t["k"] for t in teach], dist=[], weights=[], saved=[])
    for i, (gram, col) in enumerate(student_stats):
        g, c = rotate_stats(gram, col, proj_s)
        lam, vec = sym_eig_desc(centred(g, c, rows_s))
        dist = torch.zeros(len(teach))
        per = []
        for j, t in enumerate(teach):
            k = t["k"]
            full = vec.T @ t["basis"]                       # (D,k) = V_s^T U_t
            ux, sig, vxt = torch.linalg.svd(full[:k].double())
            ux, sig, vxt = ux.float(), sig.float(), vxt.float()
            theta = torch.acos(sig.clamp(max=1.0 - EPS32))
            dist[j] = (t["sw"] * theta ** 2).sum() / t["sw"].sum()
            per.append(dict(full=full, ux=ux, sig=sig, vx=vxt.T, theta=theta))
        w = torch.softmax(-dist / temps[i], dim=0)
        out["dist"].append(dist)
        out["weights"].append(w)
        out["saved"].append(dict(lam=lam, vec=vec, per=per, colsum_tok=col))
    out["teach"] = teach
    out["temps"] = temps
    return out


def selector_backward_model(sel, d_weights, rows_s, proj_s, log_temps):
    """Given dL/dweights (E,L): returns dL/dlog_temps (E,) and, per student layer, the
    token-space matrix W' (D_s x D_s) with dL/dS += (S - 1 mean^T) W'.
    Closed form of the SVD/eig backward: only cross-block terms."""
    d_logt = torch.zeros_like(log_temps)
    w_primes = []
    temps = sel["temps"]
    for i, saved in enumerate(sel["saved"]):
        w = sel["weights"][i]
        dist = sel["dist"][i]
        dy = w * (d_weights[i] - (w * d_weights[i]).sum())          # softmax backward
        d_dist = -dy / temps[i]
        d_tau = (dy * dist).sum() / temps[i] ** 2
        d_logt[i] = d_tau * torch.sigmoid(log_temps[i])             # softplus backward
        lam, vec = saved["lam"], saved["vec"]
        dim = lam.numel()
        omega = torch.zeros(dim, dim)
        for j, t in enumerate(sel["teach"]):
            k, p = t["k"], saved["per"][j]
            sw = t["sw"]
            sig, theta = p["sig"], p["theta"]
            live = (sig < 1.0 - EPS32).float()                      # clamp kills the grad
            d_sig = d_dist[j] * (sw / sw.sum()) * 2 * theta * (-1.0 / (1 - sig ** 2).clamp(min=1e-30).sqrt()) * live
            lower = p["full"][k:]                                   # V_perp^T U_t  (D-k, k)
            block = (lower @ p["vx"]) * d_sig @ p["ux"].T           # (D-k, k)
            gap = lam[:k].unsqueeze(0) - lam[k:].unsqueeze(1)       # lam_i - lam_j
            omega[k:, :k] += block / gap
        d_gram = vec @ omega @ vec.T
        w_sym = d_gram + d_gram.T
        w_primes.append(proj_s.T @ w_sym @ proj_s)
        saved["omega"] = omega
        saved["d_dist"] = d_dist
    return d_logt, w_primes


# ------------------------------------------------------------------ stage: mixing
def _round_like(x, dtype):
    return x.to(dtype).float()


def interp_taps(n_src: int, n_dst: int):
    """1-D linear, align_corners=False (reference: combined.py:12; relational.py:30)."""
    idx = torch.arange(n_dst, dtype=torch.float32)
    src = ((idx + 0.5) * (n_src / n_dst) - 0.5).clamp(min=0)
    lo = src.floor().long().clamp(max=n_src - 1)
    hi = (lo + 1).clamp(max=n_src - 1)
    frac = src - lo.float()
    return lo, hi, frac


def mix_and_align(weights_row, teacher_stack, n_student):
    """Mixed + aligned teacher tokens for one student layer (kernel: mix_interp).
    fp32 arithmetic on the (exactly upcast) tokens -- layer_selector.py:110-111 then
    combined.py:12.  The kernel stores bf16 only when the tokens are bf16 AND no resampling
    happens; resampled tokens are rank-deficient and their Procrustes term is sensitive to
    rounding noise (1.5e-3 on the loss, gradient cosine 0.97), so they stay fp32."""
    dt = teacher_stack.dtype
    mixed = (weights_row.view(-1, 1, 1, 1) * teacher_stack.float()).sum(dim=0)   # (B,N_t,D_t)
    n_t = mixed.shape[1]
    if n_t != n_student:
        lo, hi, frac = interp_taps(n_t, n_student)
        return mixed[:, lo] * (1 - frac).view(1, -1, 1) + mixed[:, hi] * frac.view(1, -1, 1)
    return _round_like(mixed, dt)


def attn_rows(attn: torch.Tensor, has_cls: bool):
    """Per-layer importance row (kernel: attn_rows) -- relational.py:22-27, before mixing.
    fp32 attention maps only in the model (bf16 maps are handled in the kernel test)."""
    a = attn.float()
    return a[:, :, 0, 1:].mean(dim=1) if has_cls else a.mean(dim=(1, 2))


def mix_importance(weights_row, rows_stack, n_student):
    """(L,B,N_t) rows -> normalised (B,N_s) importance + pre-normalisation sum."""
    mixed = (weights_row.view(-1, 1, 1) * rows_stack).sum(dim=0)
    n_t = mixed.shape[1]
    if n_t != n_student:
        lo, hi, frac = interp_taps(n_t, n_student)
        mixed = mixed[:, lo] * (1 - frac) + mixed[:, hi] * frac
    total = mixed.sum(dim=1, keepdim=True)
    return mixed / total, total


# ------------------------------------------------------------------ stage: Procrustes
def pivoted_cholesky(k: torch.Tensor, rel_tol: float = 1e-5):
    """Diagonally pivoted (rank-revealing) Cholesky of a PSD matrix, stopping when the
    largest remaining pivot drops below rel_tol * largest initial diagonal
    (kernel: procrustes_factor).  Returns L (N x N, zero columns beyond the rank) with
    L L^T ~= K."""
    n = k.shape[0]
    a = k.clone()
    low = torch.zeros_like(k)
    floor = rel_tol * a.diagonal().max()
    for j in range(n):
        diag = a.diagonal()
        p = int(diag.argmax())
        if not diag[p] > floor:
            break
        col = a[:, p] / diag[p].sqrt()
        low[:, j] = col
        a -= torch.outer(col, col)
        a[p, :] = 0                      # exact zero of the eliminated row/col
        a[:, p] = 0
    return low


def procrustes_sample(s_tok, t_tok, w, *, sv_floor: float = 1e-5, rel_tol: float = 1e-5,
                      direct_sv_floor: float = 1e-5):
    """One sample. s_tok (N,Ds), t_tok (N,Dt) fp32, w (N,) normalised.
    Returns value f = tr_s + tr_t - 2 nuc and the closed-form pieces of its gradient.

    Each side gets a factor F with F F^T = K (its N x N Gram):  when D <= N the weighted,
    centred tokens themselves (F = A, r = D columns: no Gram, no squared condition number);
    otherwise the pivoted-Cholesky factor of K (r = N).  X = F_s^T F_t (r_s x r_t) = U S V^T has
    the singular values of the cross-covariance A^T B, and
      d nuc/dA = Y_A A,  Y_A = (F_t V) S^+ (F_t V)^T ;   d nuc/dB = Y_B B,  Y_B = (F_s U) S^+ (F_s U)^T
      diag(A polar(A^T B) B^T) = rowdot(F_s U, F_t V).
    The one-sided Jacobi sweep orthogonalises the rows of the orientation of X with fewer
    rows (r_q <= r_p) and yields the p-side vectors; the q-side ones are the normalised rows
    of P^T X (no inverse of any factor is ever formed)."""
    n = s_tok.shape[0]
    root = w.sqrt().unsqueeze(1)
    a = root * (s_tok - (w.unsqueeze(1) * s_tok).sum(0, keepdim=True))
    b = root * (t_tok - (w.unsqueeze(1) * t_tok).sum(0, keepdim=True))
    diag_s, diag_t = (a * a).sum(dim=1), (b * b).sum(dim=1)
    direct_s, direct_t = a.shape[1] <= n, b.shape[1] <= n
    if direct_s != direct_t and (b.shape[1] if direct_s else a.shape[1]) <= 2 * n:
        direct_s = direct_t = True                        # see _engine.procrustes_forward
    f_s = a if direct_s else pivoted_cholesky(a @ a.T, rel_tol)
    f_t = b if direct_t else pivoted_cholesky(b @ b.T, rel_tol)
    floor = direct_sv_floor if (direct_s and direct_t) else sv_floor
    swap = f_t.shape[1] > f_s.shape[1]                    # q = the side with fewer columns
    f_p, f_q = (f_t, f_s) if swap else (f_s, f_t)
    g = f_q.T @ f_p                                       # (r_q, r_p); rows -> sigma_j p_j^T
    _, sig, pt = torch.linalg.svd(g.double(), full_matrices=False)
    sig, pt = sig.float(), pt.float()                     # pt rows = p-side singular vectors
    keep = sig > floor * sig.max()
    rows = pt @ g.T                                       # S Q^T, row j has norm sig_j
    norms = rows.norm(dim=1)
    qt = torch.where(keep.unsqueeze(1), rows / norms.clamp(min=1e-30).unsqueeze(1), torch.zeros_like(rows))
    inv_sig = torch.where(keep, 1.0 / sig.clamp(min=1e-30), torch.zeros_like(sig))
    nuc = sig.sum()
    fq = f_q @ qt.T                                       # F_q Q   (N, r_q)
    fp = (f_p @ pt.T) * keep                              # F_p P
    pi_diag = (fp * fq).sum(dim=1)
    f = diag_s.sum() + diag_t.sum() - 2 * nuc

    def side_grad(tok, direct, mine, other, my_vecs):
        """d nuc / d tok.  Direct side: (F_other W)(own D-space vectors)^T -- unit vectors only,
        no 1/sigma.  Gram side: Y tok with Y = (F_other W) S^+ (F_other W)^T."""
        if direct:
            return other @ my_vecs
        return ((other * inv_sig) @ other.T) @ tok

    # p side: own vectors pt (rows), other-side image fq; q side: own vectors qt, image fp
    d_p = side_grad(b if swap else a, direct_t if swap else direct_s, fp, fq, pt * keep.unsqueeze(1))
    d_q = side_grad(a if swap else b, direct_s if swap else direct_t, fq, fp, qt)
    d_a, d_b = (d_q, d_p) if swap else (d_p, d_q)
    grad_s = 2 * root * (a - d_a)                         # df/dS      (w held fixed)
    grad_t = 2 * root * (b - d_b)                         # df/dR'
    grad_w = (diag_s + diag_t - 2 * pi_diag) / w          # df/dw (normalised w)
    return f, grad_s, grad_t, grad_w


# ------------------------------------------------------------------ whole step
def full_step_model(logits, targets, students, teachers, attns, *, layers, proj_s, proj_t,
                    log_temps, n_student, has_cls, criterion):
    """Forward + hand-derived backward of the whole loss, assembled from the stages
    above exactly as the CUDA host code assembles the kernels.
    Returns dict(loss, ce, geo, ranks, weights, grad_students{layer}, grad_log_temps,
    grad_logits)."""
    t_keys = sorted(teachers.keys())
    t_stack = torch.stack([teachers[k] for k in t_keys])              # (L,B,Nt,Dt)
    n_l, bsz, n_t, d_t = t_stack.shape
    rows_t = bsz * n_t
    rows_s = bsz * students[layers[0]].shape[1]
    sel = selector_model([token_stats(students[l]) for l in layers],
                         [token_stats(teachers[k]) for k in t_keys],
                         rows_s, rows_t, proj_s, proj_t, log_temps.detach())
    rows = torch.stack([attn_rows(attns[k], has_cls) for k in t_keys])  # (L,B,Nt)

    geo_terms, g_s, g_t, g_wt = [], {}, {}, {}
    for i, layer in enumerate(layers):
        w_mix = sel["weights"][i]
        aligned = mix_and_align(w_mix, t_stack, n_student)
        imp, total = mix_importance(w_mix, rows, n_student)
        vals = []
        gs = torch.zeros(bsz, n_student, students[layer].shape[2])
        gt = torch.zeros(bsz, n_student, d_t)
        gw = torch.zeros(bsz, n_student)
        for b in range(bsz):
            f, a_, b_, c_ = procrustes_sample(students[layer][b].float(), aligned[b].float(), imp[b])
            vals.append(f)
            gs[b], gt[b] = a_, b_
            gw[b] = (c_ - f) / total[b]              # through w = w~/sum(w~): sum_n w_n df/dw_n = f
        geo_terms.append(torch.stack(vals).mean())
        g_s[layer], g_t[layer], g_wt[layer] = gs, gt, gw

    logits = logits.detach().requires_grad_(True)
    ce = criterion(logits, targets)
    geo = torch.stack(geo_terms).mean()
    inv = torch.stack([1.0 / ce.detach().clamp(min=EPS32), 1.0 / geo.clamp(min=EPS32)])
    share = inv / inv.sum()
    loss = share[0] * ce.detach() + share[1] * geo
    (share[0] * ce).backward()

    scale = share[1] / (len(layers) * bsz)           # dloss/df_{i,b}
    d_weights = torch.zeros(len(layers), n_l)
    grad_students = {}
    for i, layer in enumerate(layers):
        # teacher-token path: <dL/dR', A_interp T_l>; importance path: <dL/dw~, A_interp a_l>
        if n_t != n_student:
            lo, hi, frac = interp_taps(n_t, n_student)
            up_tok = lambda z: z[:, lo] * (1 - frac).view(1, -1, 1) + z[:, hi] * frac.view(1, -1, 1)
            up_row = lambda z: z[:, lo] * (1 - frac) + z[:, hi] * frac
        else:
            up_tok = up_row = lambda z: z
        for j in range(n_l):
            d_weights[i, j] = scale * ((g_t[layer] * up_tok(t_stack[j].float())).sum()
                                       + (g_wt[layer] * up_row(rows[j])).sum())
        grad_students[layer] = scale * g_s[layer]
    d_logt, w_primes = selector_backward_model(sel, d_weights, rows_s, proj_s, log_temps.detach())
    grad_direct = {l: g.clone() for l, g in grad_students.items()}
    for i, layer in enumerate(layers):
        x = students[layer].float().reshape(rows_s, -1)
        centred = x - x.mean(dim=0, keepdim=True)
        grad_students[layer] = grad_students[layer] + (centred @ w_primes[i]).reshape(bsz, n_student, -1)
    return dict(loss=loss.detach(), ce=ce.detach(), geo=geo.detach(), ranks=sel["ranks"],
                weights=torch.stack(sel["weights"]), dist=torch.stack(sel["dist"]),
                grad_students=grad_students, grad_log_temps=d_logt, grad_logits=logits.grad,
                d_weights=d_weights, grad_direct=grad_direct, w_primes=w_primes, sel=sel)
