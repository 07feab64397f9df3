"""Host-side orchestration of the BASD kernels.

Four device phases, each a fixed sequence of C-ABI calls enqueued on the current CUDA
stream with no host synchronisation in between:

  statistics  -> (all-reduce across data-parallel ranks) -> selector -> Procrustes forward
  backward: Procrustes grads -> mixing-weight grads -> (all-reduce) -> selector backward

torch is used for allocation, streams and collectives only.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import torch
import torch.distributed as dist

from . import _native as nat
from ._native import call, ptr, stream

import os as _os

# Numerical constants of the path (how each was chosen: DESIGN.md section 3.3; the sweeps that
# measured them: tests/tools/floor_sweep.py, tests/tools/parity_report.py).
JACOBI_TOL = 1e-6        # |cos| between two rows below which they count as orthogonal
JACOBI_SWEEPS = 18       # sweep cap (C2 backbone features converge in 8-11)
# Symmetric eigenproblems of the selector Grams: the sweeps end only when every rotated pair had |cos| below
# this.  The default sqrt(tol) = 1e-3 leaves rows of nearly equal norm mixed by ~cos / (relative gap); with
# 50,176 token rows the spectrum is dense (gaps ~1e-3 at the MP rank boundary) and the selector gradient came
# out with cosine 0.87 against autograd through the reference (C2, B = 256) while every small-batch case passed.
EIG_STOP_COS = 1e-3
PROC_JACOBI_TOL = 1e-6   # per-sample Procrustes SVDs: loosening costs gradient parity before it saves a sweep
CHOL_TOL = 1e-5          # pivoted-Cholesky rank cut (per-sample N x N Grams), relative to the
                         # largest diagonal: just above the fp32 accumulation noise of K
GRAM_CHOL_TOL = 1e-7     # same for the D x D selector Grams (after the diagonal shift below it never triggers)
# sym_eig factors K + GRAM_SHIFT * max(diag K) * I: identical eigenvectors, and no noise-sized pivots when
# fp32 rounding has pushed the smallest eigenvalues of a nearly singular Gram to zero or below.  Measured
# without it (C2, B = 256): one eigenvalue at -5e-5 lambda_max made the factorisation lose 12 rows and rotated
# the top-k student subspace by sin 0.6; stopping the factorisation at a higher tolerance instead drops up
# to (D - r) pivots of mass, more than the eigenvalue gap at the rank boundary (C4: selector gradient 0.88).
# The eigenvalues themselves come from Rayleigh quotients against the unshifted matrix.
GRAM_SHIFT = 1e-5        # ~10x the fp32 noise of the Schur complement (sqrt(D) eps max diag)
SV_FLOOR = 1e-6          # k x k principal-angle SVD: directions below this are dropped
ROW_FLOOR = 1e-7         # rows this far below the largest row norm are numerically zero in fp32
# Procrustes with a Gram side: singular directions of G = F_q^T F_p below this fraction of sigma_max
# leave the polar factor.  The factors are already cut at sqrt(CHOL_TOL) = 3e-3 of their own largest
# direction, so G's spectrum ends near 1e-5 by itself.  2.5e-4 (the first choice) dropped directions the
# reference keeps: per-sample tokens with condition number 2e3 gave gradient cosine 0.99909, 0.99999
# at 1e-5, with every well-conditioned case unchanged (measured on B200).
PROC_SV_FLOOR = 1e-5
# ... except for the side whose vectors are derived (q_j = normalise(G p_j), direction error
# eps sigma_max / sigma_j) when they enter a Gram-side operator sum_j (F_q q_j)(F_q q_j)^T / sigma_j:
# at 1e-5 the teacher-token gradient of ill-conditioned tokens came out 41x too large (B200)
PROC_SV_FLOOR_DERIVED = 2.5e-4
# both sides direct (no Gram): q_j = normalise(G p_j) is recovered with noise ~eps*sqrt(K)*sigma_max/sigma_j;
# measured on B200: 1e-6 -> cosine 0.998, 1e-5 -> 0.9998, 3e-5 -> 0.9999, 1e-4 -> 0.9996
PROC_SV_FLOOR_DIRECT = 3e-5
MIXED_DIRECT_RATIO = 2.0
# Equal factor widths (both sides Gram factors): the side with the smaller token dimension indexes the
# Jacobi rows (see procrustes_forward)
PROC_TIE_Q = "narrow"
# problems of one selector launch are dealt to the data-parallel ranks once they exceed one wave
SHARD_SELECTOR_EIG = _os.environ.get("BASD_NO_SELECTOR_SHARDING") is None
SHARD_WAVE_CTAS = 148


def _f32(*shape, device):
    return torch.empty(*shape, dtype=torch.float32, device=device)


def _pad4(x: int) -> int:
    return (x + 3) // 4 * 4


def _mat(p: int, rows: int, cols: int, device):
    """(p, rows, pitch) fp32 with the pitch rounded up to a multiple of four floats: the factor
    kernels read rows in 128-bit pieces, so a token count like 49, 225 or 577 needs padded rows.
    The pad columns are zero and stay zero (rotations and products of zeros); the logical width is
    passed to every kernel next to the pitch."""
    pitch = _pad4(cols)
    if pitch == cols:
        return torch.empty(p, rows, cols, dtype=torch.float32, device=device)
    return torch.zeros(p, rows, pitch, dtype=torch.float32, device=device)


# ---------------------------------------------------------------------------- op wrappers
# the Procrustes products run on the tensor cores (3xTF32, gemm_tc3.cu) unless this is set
# (A/B measurements and the SIMT-vs-tensor parity test)
TC_GEMM = _os.environ.get("BASD_NO_TC_GEMM") is None
ROTATE_F64 = _os.environ.get("BASD_ROTATE_F32") is None    # A/B knob: fp32 projection of the statistics


def sgemm(ta, tb, m, n, k, a, lda, sa, b, ldb, sb, c, ldc, sc, batch=1, alpha=1.0,
          alpha_dev=None, beta=0.0, a_shift=None, tc=False):
    """C = alpha op(A) op(B) + beta C, batched.  `tc=True` asks for the tcgen05 3xTF32 kernel
    (fp32 operands, beta = 0, 4-float aligned pitches); anything it does not accept runs on the
    SIMT kernel."""
    if (tc and TC_GEMM and beta == 0.0 and a_shift is None and a.dtype == torch.float32
            and min(m, n) >= 32 and k >= 16 and batch <= 65535
            and not ((a.data_ptr() | b.data_ptr() | c.data_ptr()) & 15)
            and nat.load().basd_gemm_tc3_supported(m, n, k, lda, ldb, ldc, sa, sb, sc)):
        call("basd_gemm_tc3_batched", int(ta), int(tb), m, n, k, ptr(a), lda, sa, ptr(b), ldb, sb,
             ptr(c), ldc, sc, batch, float(alpha), ptr(alpha_dev), stream())
        return
    call("basd_sgemm_batched", int(ta), int(tb), m, n, k, ptr(a), nat.dtype_code(a), lda, sa,
         ptr(a_shift), ptr(b), ldb, sb, ptr(c), ldc, sc, batch, float(alpha), ptr(alpha_dev),
         float(beta), stream())


def gemm_tc_ex(ta, tb, m, n, k, a, lda, sa, b, ldb, sb, c, ldc, sc, batch=1, alpha=1.0, alpha_dev=None,
               col_sub=None) -> bool:
    """Extended tcgen05 GEMM: A may be bf16 (ta = 0), C may be bf16, optional column shift
    (C = alpha (A B - 1 col_sub^T)).  Returns False (nothing launched) when the shape or the
    alignment is not accepted, so that the caller takes its SIMT route."""
    a_bf16 = a.dtype == torch.bfloat16
    if (not TC_GEMM or min(m, n) < 32 or k < 16 or batch > 65535
            or ((a.data_ptr() | b.data_ptr() | c.data_ptr()) & 15)
            or not nat.load().basd_gemm_tc3_supported(m, n, k, lda, ldb, ldc, sa, sb, sc)
            or (a_bf16 and (ta or k % 8 or lda % 8 or sa % 8))):
        return False
    call("basd_gemm_tc3_batched_ex", int(ta), int(tb), m, n, k, ptr(a), nat.dtype_code(a), lda, sa,
         ptr(b), ldb, sb, ptr(c), nat.dtype_code(c), ldc, sc, batch, float(alpha), ptr(alpha_dev),
         ptr(col_sub), stream())
    return True


def token_gram(tokens: torch.Tensor, gram: torch.Tensor, colsum: torch.Tensor | None, mu0=None):
    """gram (D,D) = X'^T X', colsum (D) = X'^T 1 (skipped when None) for X' = tokens.reshape(-1, D) - mu0
    (mu0: (D,) fp32 with bf16-representable values, or None for no shift).  The tokens themselves are
    never modified: the kernels work the shift into the accumulation (gram_tc.cu / gemm_simt.cu)."""
    d = tokens.shape[-1]
    rows = tokens.numel() // d
    x = tokens if tokens.is_contiguous() else tokens.contiguous()
    if (x.dtype == torch.bfloat16 and nat.has("basd_token_gram_tc") and d % 128 == 0
            and x.data_ptr() % 16 == 0 and (mu0 is None or colsum is not None)):
        nbytes = nat.load().basd_token_gram_tc_workspace_bytes(rows, d)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
        call("basd_token_gram_tc", ptr(x), rows, d, ptr(mu0), ptr(gram), ptr(colsum), ptr(ws), stream())
        return
    nfl = nat.load().basd_token_gram_simt_workspace_floats(rows, d)
    ws = _f32(nfl, device=x.device)
    call("basd_token_gram_simt", ptr(x), nat.dtype_code(x), rows, d, ptr(mu0), ptr(gram), ptr(colsum),
         ptr(ws), stream())


def pivoted_cholesky(k, lt, rel_tol, dims=None, rank_out=None):
    """k, lt: (batch, n, pitch) with pitch >= n (see _mat)."""
    batch, n, ldk = k.shape
    ldl = lt.shape[2]
    call("basd_pivoted_cholesky", ptr(k), n, ldk, n * ldk, ptr(lt), ldl, n * ldl, batch, rel_tol,
         ptr(rank_out), ptr(dims), stream())


# (name, n, m, sweeps, rotations) of every Jacobi launch of the last step, filled only while
# `jacobi_log` is a list (bench.py's roofline leg turns it on for one step)
jacobi_log = None


def jacobi_rows(g, dims=None, sweeps_out=None, tag="jacobi", tol=None, row_dims=None, cols=None, stop_cos=None):
    """dims: active leading size of square problems (rows and columns); row_dims: number of
    leading non-zero rows (rank of the factor the rows come from), all columns active.
    cols: logical row length when the rows are padded (g.shape[2] is the pitch).
    stop_cos: the sweeps end after one whose rotated pairs all had |cos| below it (default sqrt(tol):
    quadratic convergence; eigenvector problems pass EIG_STOP_COS)."""
    batch, n, ld = g.shape
    m = ld if cols is None else cols
    tol = JACOBI_TOL if tol is None else tol
    stop_cos = tol ** 0.5 if stop_cos is None else stop_cos
    rot = None
    if jacobi_log is not None:
        sweeps_out = sweeps_out if sweeps_out is not None else torch.zeros(batch, dtype=torch.int32, device=g.device)
        rot = torch.zeros(batch, dtype=torch.int32, device=g.device)
        jacobi_log.append((tag, n, m, dims, sweeps_out, rot, row_dims))
    call("basd_jacobi_rows_ex", ptr(g), n, m, ld, n * ld, batch, ptr(dims), ptr(row_dims), tol, stop_cos,
         JACOBI_SWEEPS, ptr(sweeps_out), ptr(rot), stream())


def rows_normalize(g, out, vals, *, sort, square, rel_floor, dims=None, cols=None):
    batch, n, ld = g.shape
    m = ld if cols is None else cols
    ldo = out.shape[2]
    call("basd_rows_normalize", ptr(g), n, m, ld, n * ld, ptr(out), ldo, n * ldo, ptr(vals), batch,
         int(sort), int(square), rel_floor, ptr(dims), stream())


last_eig_sweeps = []      # per sym_eig call of the current step (diagnostics: SelectorState.sweeps["eig"])


def sym_eig(kmats: torch.Tensor):
    """Batched symmetric PSD eigendecomposition: returns (lam (b,D) descending, Vt (b,D,D)
    with eigenvectors as rows). Pivoted Cholesky -> row-Jacobi on the factor -> sort -> completion of
    the directions behind the rank cut -> Rayleigh-quotient refinement against the untouched matrix
    (the completed rows are a basis of the numerical null space, their lam are its Rayleigh quotients)."""
    batch, d, _ = kmats.shape
    dev = kmats.device
    work = kmats.clone()
    call("basd_shift_diag", ptr(work), d, GRAM_SHIFT, batch, stream())
    lt = _f32(batch, d, d, device=dev)
    pivoted_cholesky(work, lt, GRAM_CHOL_TOL)
    sweeps = torch.zeros(batch, dtype=torch.int32, device=dev)
    jacobi_rows(lt, tag="eig", stop_cos=EIG_STOP_COS, sweeps_out=sweeps)
    last_eig_sweeps.append(sweeps)
    vt = work                         # reuse: the Schur complement is dead
    coarse = _f32(batch, d, device=dev)
    rows_normalize(lt, vt, coarse, sort=True, square=True, rel_floor=ROW_FLOOR)
    complete_null_space(vt)           # rows behind the rank cut: an orthonormal basis of the complement
    kv = lt                           # reuse
    sgemm(0, 0, d, d, d, vt, d, d * d, kmats, d, d * d, kv, d, d * d, batch)
    lam = _f32(batch, d, device=dev)
    call("basd_rowdot", ptr(kv), d, d * d, ptr(vt), d, d * d, d, d, batch, ptr(lam), stream())
    return lam, vt


def complete_null_space(vt: torch.Tensor):
    """In place: the zero rows sym_eig leaves beyond the numerical rank of a Gram become an orthonormal
    basis of its null space (selector.cu: projector + pivoted Cholesky).  The thin-SVD backward's
    (I - V V^T) term lives there: with fewer token rows than dimensions (layer_selector.py:14-15), and
    when a spectrum spanning more than 1/eps pushes an eigenvalue under the row floor (its eigenvector is
    then the complement of all the others).  Problems whose basis is complete exit at once (device-side
    `dims`), so the common case costs four small launches."""
    batch, d, _ = vt.shape
    dev = vt.device
    vtv = _f32(batch, d, d, device=dev)
    sgemm(1, 0, d, d, d, vt, d, d * d, vt, d, d * d, vtv, d, d * d, batch, tc=True)
    proj = _f32(batch, d, d, device=dev)
    active = torch.empty(batch, dtype=torch.int32, device=dev)
    call("basd_projector_complement", ptr(vtv), d, ptr(proj), ptr(active), batch, stream())
    rank_p = torch.empty(batch, dtype=torch.int32, device=dev)
    pivoted_cholesky(proj, vtv, 1e-3, dims=active, rank_out=rank_p)      # projector eigenvalues are 0 or 1
    call("basd_place_complement", ptr(vt), ptr(vtv), ptr(rank_p), d, batch, stream())


def sharded_sym_eig(kmats: torch.Tensor, group, world: int, solver=None):
    """sym_eig with the problems dealt round-robin to the data-parallel ranks.  After the
    statistics all-reduce every rank holds bit-identical Grams, and without this every rank would
    repeat the same L_t + E eigenproblems -- latency-bound work that occupies a few SMs (9 ms of the
    44 ms C2 step, 110 of 212 ms at C4).  Rank r factors problems r, r + W, ...; one all-gather over
    NVLink (16 x (D^2 + D) floats = 9.4 MB at C2, 66 MB at C4) hands everyone the full set.  Every
    problem is solved by one CTA / cluster with a fixed schedule, so the result does not depend on
    which rank solved it."""
    solver = sym_eig if solver is None else solver
    # the problems of one launch run concurrently (one CTA cluster each): dealing them out only
    # pays once a launch exceeds one wave of the GPU (C4: 28 x 768^2; at C2 the 16 x 384^2 problems
    # fill 64 of 148 SMs and sharding measured 44.4 vs 44.0 ms per step on 2 GPUs)
    q_all, d_all = kmats.shape[0], kmats.shape[1]
    one_wave = q_all * max(1, d_all // 96) <= SHARD_WAVE_CTAS
    if world <= 1 or SHARD_SELECTOR_EIG is False or (one_wave and solver is sym_eig):
        return solver(kmats)
    rank = dist.get_rank(group)
    q, d, _ = kmats.shape
    per = (q + world - 1) // world
    own = torch.arange(per, device=kmats.device) * world + rank
    own = own.clamp(max=q - 1)                            # ranks short of a problem redo the last one
    lam_l, vt_l = solver(kmats.index_select(0, own).contiguous())
    packed = torch.cat([lam_l.reshape(per, -1), vt_l.reshape(per, -1)], dim=1).contiguous()
    gathered = torch.empty(world * per, d + d * d, dtype=torch.float32, device=kmats.device)
    dist.all_gather_into_tensor(gathered, packed, group=group)
    # gathered row (r * per + s) is problem s * world + r
    order = (torch.arange(q, device=kmats.device) % world) * per + torch.arange(q, device=kmats.device) // world
    full = gathered.index_select(0, order)
    return full[:, :d].contiguous(), full[:, d:].reshape(q, d, d).contiguous()


# ---------------------------------------------------------------------------- phases
@dataclass
class Stats:
    gram_s: torch.Tensor      # (E, Ds, Ds) token-space
    col_s: torch.Tensor       # (E, Ds)  column sums in the same (mean-shifted) frame as the Grams
    gram_t: torch.Tensor      # (L, Dt, Dt)
    col_t: torch.Tensor       # (L, Dt)
    rows: torch.Tensor        # (L, B, Nt) attention importance rows (local shard)
    mu0_s: torch.Tensor | None = None     # (E, Ds) shift applied to the student tokens before the Gram
    mu0_t: torch.Tensor | None = None     # (L, Dt)


def _all_reduce(flat: torch.Tensor, group):
    if group is not None or (dist.is_available() and dist.is_initialized()
                             and dist.get_world_size() > 1):
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)


def importance_rows(attns, b: int, has_cls: bool, dev) -> torch.Tensor:
    """(L, B, n_rows) fp32 importance rows: the CLS query row averaged over heads, or the mean over
    heads and queries without a CLS token (relational.py:22-27).  n_rows is the ATTENTION map's own
    token count (side minus the CLS column); it equals the teacher token count inside BASDLoss and
    may differ from it in the stand-alone loss, whose tokens arrive already resampled
    (relational.py:29-32) -- the mixing and gradient kernels resample the row to the student grid."""
    widths = set()
    for a in attns:
        if a.shape[0] != b:
            raise ValueError(f"attention batch {a.shape[0]} differs from the token batch {b}")
        if a.dim() == 2:
            widths.add(a.shape[1])
        elif a.dim() in (3, 4):
            widths.add(a.shape[-1] - (1 if has_cls else 0))
        else:
            raise ValueError(f"attention must be (B,H,Q,K), (B,H,K) or a (B,N) row, got {tuple(a.shape)}")
    if len(widths) != 1:
        raise ValueError(f"teacher attention maps disagree on the token count: {sorted(widths)}")
    n_rows = widths.pop()
    if n_rows < 1:
        raise ValueError("attention map has no token columns")
    rows = _f32(len(attns), b, n_rows, device=dev)
    for j, a in enumerate(attns):
        a = a if a.is_contiguous() else a.contiguous()
        if a.dim() == 2:      # already an importance row (B, n_rows): "next" row f1 of SURVEY §8
            rows[j].copy_(a)
            continue
        if a.dim() == 3:      # (B, H, side): only the CLS query row was handed over
            _, h, side = a.shape
            q_rows = 1
        else:
            _, h, q_rows, side = a.shape
        call("basd_attn_rows", ptr(a), nat.dtype_code(a), b, h, side, q_rows, int(has_cls),
             ptr(rows[j]), stream())
    return rows


ROUGH_MEAN_ROWS = 1024        # evenly spaced token rows behind the rough mean the statistics are shifted by


def _ptr_array(tensors):
    return (nat.C.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])


def statistics(students, teachers, attns, has_cls, group=None, world: int = 1):
    """Token-space Grams and column sums of every student / teacher tensor, in a MEAN-SHIFTED frame
    (the Gram kernels subtract a rough mean inside the accumulation so that the accumulators never
    hold the M mu mu^T term the centring would otherwise have to cancel), plus the importance rows.

    Data parallel (world > 1): every rank accumulates in its OWN frame (the rough mean of its shard) and
    ONE all-reduce of a flat buffer sums the Grams and carries each rank's column sums and shift in a slot
    of its own; `basd_merge_shifted_stats` then moves everything to the common frame.  Returns
    (Stats of the global batch, the flat buffer that was exchanged)."""
    dev = students[0].device
    e, l = len(students), len(teachers)
    b, n_t, d_t = teachers[0].shape
    n_s, d_s = students[0].shape[1], students[0].shape[2]
    slot = e * d_s + l * d_t
    sizes = [e * d_s * d_s, e * d_s, l * d_t * d_t, l * d_t]
    extra = 2 * world * slot if world > 1 else 0
    flat = _f32(sum(sizes) + extra, device=dev)
    offs = [0]
    for s in sizes:
        offs.append(offs[-1] + s)
    gram_s = flat[offs[0]:offs[1]].view(e, d_s, d_s)
    col_s = flat[offs[1]:offs[2]].view(e, d_s)
    gram_t = flat[offs[2]:offs[3]].view(l, d_t, d_t)
    col_t = flat[offs[3]:offs[4]].view(l, d_t)
    mu0 = _f32(slot, device=dev)
    mu0_s, mu0_t = mu0[:e * d_s].view(e, d_s), mu0[e * d_s:].view(l, d_t)
    for tensors, nrows, d, m0 in ((students, b * n_s, d_s, mu0_s), (teachers, b * n_t, d_t, mu0_t)):
        call("basd_rough_means", _ptr_array(tensors), len(tensors), nat.dtype_code(tensors[0]), nrows, d,
             min(nrows, ROUGH_MEAN_ROWS), ptr(m0), stream())
    # the correction operand of the tensor-core Gram holds 8 mu0 in bf16: it must be exact
    mu0 = mu0.to(torch.bfloat16).to(torch.float32)
    mu0_s, mu0_t = mu0[:e * d_s].view(e, d_s), mu0[e * d_s:].view(l, d_t)
    for i, s in enumerate(students):
        token_gram(s, gram_s[i], col_s[i], mu0_s[i])
    for j, t in enumerate(teachers):
        token_gram(t, gram_t[j], col_t[j], mu0_t[j])
    if world > 1:
        # slots (world, E, D_s) / (world, L, D_t) for the column sums and for the shifts: this rank fills its
        # own, the all-reduce (sum) delivers everybody's
        rank = dist.get_rank(group)
        tail = flat[offs[4]:]
        tail.zero_()
        d_s_slots = tail[:world * e * d_s].view(world, e, d_s)
        m_s_slots = tail[world * e * d_s:2 * world * e * d_s].view(world, e, d_s)
        rest = tail[2 * world * e * d_s:]
        d_t_slots = rest[:world * l * d_t].view(world, l, d_t)
        m_t_slots = rest[world * l * d_t:].view(world, l, d_t)
        d_s_slots[rank].copy_(col_s)
        m_s_slots[rank].copy_(mu0_s)
        d_t_slots[rank].copy_(col_t)
        m_t_slots[rank].copy_(mu0_t)
        _all_reduce(flat, group)
        mu0 = _f32(slot, device=dev)
        mu0_s, mu0_t = mu0[:e * d_s].view(e, d_s), mu0[e * d_s:].view(l, d_t)
        call("basd_merge_shifted_stats", ptr(gram_s), ptr(col_s), ptr(mu0_s), ptr(d_s_slots), ptr(m_s_slots), world,
             e, d_s, b * n_s, stream())
        call("basd_merge_shifted_stats", ptr(gram_t), ptr(col_t), ptr(mu0_t), ptr(d_t_slots), ptr(m_t_slots), world,
             l, d_t, b * n_t, stream())
    rows = importance_rows(attns, b, has_cls, dev)
    return Stats(gram_s, col_s, gram_t, col_t, rows, mu0_s, mu0_t), flat


def attention_only_stats(teachers, attns, has_cls):
    """Importance rows only (stand-alone Procrustes use: no selector statistics needed)."""
    b = teachers[0].shape[0]
    return Stats(None, None, None, None, importance_rows(attns, b, has_cls, teachers[0].device))


@dataclass
class SelectorState:
    ranks: torch.Tensor       # (L,) int32
    edges: torch.Tensor       # (L,3) median, lambda_plus, tie flag (an eigenvalue within 1e-4 of the edge)
    lam_t: torch.Tensor       # (L, Ds) centred teacher eigenvalues
    lam_s: torch.Tensor       # (E, Ds)
    vt_s: torch.Tensor        # (E, Ds, Ds)
    wfull: torch.Tensor       # (E*L, Ds, Ds)  V_s^T V_t
    uxt: torch.Tensor         # (E*L, Ds, Ds)
    vxt: torch.Tensor         # (E*L, Ds, Ds)
    sig: torch.Tensor         # (E*L, Ds)
    dist: torch.Tensor        # (E, L)
    weights: torch.Tensor     # (E, L)
    temps: torch.Tensor       # (E,)
    mean_s: torch.Tensor      # (E, Ds) global token means
    sweeps: dict = field(default_factory=dict)


def selector_forward(stats: Stats, rows_s: int, rows_t: int, proj_s, proj_t, log_temps, group=None,
                     world: int = 1):
    dev = proj_s.device
    e, d_s, _ = stats.gram_s.shape
    l, d_t, _ = stats.gram_t.shape
    # rows < D_s (the reference's M < D branch, layer_selector.py:14-15): the M x M Gram it switches
    # to has the M largest eigenvalues of the D_s x D_s one; the MP kernels take the median over those
    # --- rotate the token-space statistics by the fixed projections and centre them
    kall = _f32(l + e, d_s, d_s, device=dev)
    chat_t = _f32(l, d_s, device=dev)
    chat_s = _f32(e, d_s, device=dev)
    if ROTATE_F64:
        # fp64 products, one rounding per entry of K (rotate_f64.cu): the fp32 rotation of a spectrum
        # spanning cond(X)^2 is what limits the selector's gradient on ill-conditioned tokens
        for proj, gram, col, rows, k_out, c_out in ((proj_t, stats.gram_t, stats.col_t, rows_t, kall[:l], chat_t),
                                                    (proj_s, stats.gram_s, stats.col_s, rows_s, kall[l:], chat_s)):
            nb, d_in = gram.shape[0], gram.shape[1]
            ws = torch.empty(nat.load().basd_rotate_stats_f64_workspace_bytes(d_s, d_in, nb) // 8,
                             dtype=torch.float64, device=dev)
            call("basd_rotate_stats_f64", ptr(proj), d_s, d_in, ptr(gram), ptr(col), nb, 1.0 / rows,
                 ptr(ws), ptr(k_out), ptr(c_out), stream())
    else:
        tmp_t = _f32(l, d_s, d_t, device=dev)
        sgemm(0, 0, d_s, d_t, d_t, proj_t, d_t, 0, stats.gram_t, d_t, d_t * d_t, tmp_t, d_t, d_s * d_t, l)
        ghat_t = _f32(l, d_s, d_s, device=dev)
        sgemm(0, 1, d_s, d_s, d_t, tmp_t, d_t, d_s * d_t, proj_t, d_t, 0, ghat_t, d_s, d_s * d_s, l)
        sgemm(0, 1, l, d_s, d_t, stats.col_t, d_t, 0, proj_t, d_t, 0, chat_t, d_s, 0, 1)
        tmp_s = _f32(e, d_s, d_s, device=dev)
        sgemm(0, 0, d_s, d_s, d_s, proj_s, d_s, 0, stats.gram_s, d_s, d_s * d_s, tmp_s, d_s, d_s * d_s, e)
        ghat_s = _f32(e, d_s, d_s, device=dev)
        sgemm(0, 1, d_s, d_s, d_s, tmp_s, d_s, d_s * d_s, proj_s, d_s, 0, ghat_s, d_s, d_s * d_s, e)
        sgemm(0, 1, e, d_s, d_s, stats.col_s, d_s, 0, proj_s, d_s, 0, chat_s, d_s, 0, 1)
        call("basd_center_gram", ptr(ghat_t), ptr(chat_t), d_s, 1.0 / rows_t, ptr(kall[:l]), l, stream())
        call("basd_center_gram", ptr(ghat_s), ptr(chat_s), d_s, 1.0 / rows_s, ptr(kall[l:]), e, stream())
    # the statistics live in a mean-shifted frame (statistics()): K is shift invariant, the projected column
    # SUMS get the shift back (chat = P c' + M P mu0) -- the uncentred spectrum behind the MP rank needs them
    if stats.mu0_t is not None:
        sgemm(0, 1, l, d_s, d_t, stats.mu0_t, d_t, 0, proj_t, d_t, 0, chat_t, d_s, 0, 1, alpha=float(rows_t),
              beta=1.0)
        sgemm(0, 1, e, d_s, d_s, stats.mu0_s, d_s, 0, proj_s, d_s, 0, chat_s, d_s, 0, 1, alpha=float(rows_s),
              beta=1.0)
    # --- [centred teacher | centred student] eigenproblems in one batch.  The uncentred teacher
    # spectrum (only needed for the MP rank) follows from the centred one by the rank-one
    # secular equation, so it costs no eigenproblem of its own.
    last_eig_sweeps.clear()
    lam, vt = sharded_sym_eig(kall, group, world)
    eig_sweeps = torch.cat(last_eig_sweeps) if last_eig_sweeps else None
    lam_t, lam_s = lam[:l], lam[l:]
    vt_t, vt_s = vt[:l], vt[l:]
    y_t = _f32(l, d_s, device=dev)                        # y = V^T c_hat per teacher layer
    sgemm(0, 0, d_s, 1, d_s, vt_t, d_s, d_s * d_s, chat_t, 1, d_s, y_t, 1, d_s, l)
    ranks = torch.empty(l, dtype=torch.int32, device=dev)
    edges = _f32(l, 3, device=dev)
    call("basd_mp_rank_secular", ptr(lam_t), ptr(y_t), d_s, rows_t, d_s - 1, ptr(ranks), ptr(edges), l,
         stream())
    dims = torch.empty(e * l, dtype=torch.int32, device=dev)
    call("basd_expand_ranks", ptr(ranks), e, l, ptr(dims), stream())

    # --- principal angles between span(V_s[:k]) and span(V_t[:k])
    dd = d_s * d_s
    wfull = _f32(e * l, d_s, d_s, device=dev)
    gx = _f32(e * l, d_s, d_s, device=dev)
    for i in range(e):
        sgemm(0, 1, d_s, d_s, d_s, vt_s[i], d_s, 0, vt_t, d_s, dd, wfull[i * l:], d_s, dd, l, tc=True)
        sgemm(0, 1, d_s, d_s, d_s, vt_t, d_s, dd, vt_s[i], d_s, 0, gx[i * l:], d_s, dd, l, tc=True)
    call("basd_mask_block", ptr(gx), ptr(gx), d_s, ptr(dims), e * l, stream())
    sweeps = torch.zeros(e * l, dtype=torch.int32, device=dev)
    jacobi_rows(gx, dims=dims, sweeps_out=sweeps, tag="kxk")
    uxt = _f32(e * l, d_s, d_s, device=dev)
    sig = _f32(e * l, d_s, device=dev)
    rows_normalize(gx, uxt, sig, sort=True, square=False, rel_floor=SV_FLOOR, dims=dims)
    vxt = gx                                             # reuse as rows2 = Uxt . X
    sgemm(0, 0, d_s, d_s, d_s, uxt, d_s, dd, wfull, d_s, dd, vxt, d_s, dd, e * l, tc=True)
    rows_normalize(vxt, vxt, sig, sort=False, square=False, rel_floor=SV_FLOOR, dims=dims)

    dist_el = _f32(e, l, device=dev)
    call("basd_angle_distance", ptr(sig), ptr(lam_t), ptr(ranks), d_s, e, l, ptr(dist_el), stream())
    weights = _f32(e, l, device=dev)
    temps = _f32(e, device=dev)
    call("basd_mix_weights", ptr(dist_el), ptr(log_temps), e, l, ptr(weights), ptr(temps), stream())
    mean_s = stats.col_s / float(rows_s)
    if stats.mu0_s is not None:
        mean_s = mean_s + stats.mu0_s
    return SelectorState(ranks, edges, lam_t, lam_s, vt_s, wfull, uxt, vxt, sig, dist_el,
                         weights, temps, mean_s, {"kxk": sweeps, "eig": eig_sweeps})


@dataclass
class ProcrustesState:
    a: torch.Tensor           # (E,B,N,Ds) fp32 weighted-centred student tokens
    bm: torch.Tensor          # (E,B,N,Dt)
    m_a: torch.Tensor | None  # (E*B,N,N)  2 sqrt(w) (I - Y_A)      (Gram side: D > N)
    m_b: torch.Tensor | None
    g_a: torch.Tensor | None  # (E*B,N,Ds) 2 sqrt(w) (A - (F_t V) U^T)   (direct side: D <= N)
    g_b: torch.Tensor | None
    gw: torch.Tensor | None   # (E,B,N)    d f / d w~
    f: torch.Tensor           # (E*B,)
    geo_terms: torch.Tensor   # (E,)
    geo: torch.Tensor         # ()
    aligned: torch.Tensor     # (E,B,N,Dt) mixed + aligned teacher tokens
    w: torch.Tensor           # (E,B,N)
    sweeps: torch.Tensor | None = None


class _Side:
    """One side (student or teacher) of the per-sample Procrustes problems: the weighted,
    centred tokens (p, N, D) and a factor F with F F^T = K = tok tok^T.  Gram side (D > N):
    pivoted Cholesky of K, stored transposed (r = N rows of length N).  Direct side (D <= N):
    F is the token matrix itself (r = D), no Gram and no squared condition number."""

    def __init__(self, tok: torch.Tensor, n: int, direct: bool | None = None):
        self.tok = tok
        self.p, self.n, self.d = tok.shape
        self.direct = (self.d <= n) if direct is None else direct
        self.r = self.d if self.direct else n
        dev = tok.device
        self.diag = _f32(self.p, n, device=dev)
        self.rank = None                                 # per-problem factor rank (Gram sides)
        if self.direct:
            call("basd_rowdot", ptr(tok), self.d, n * self.d, ptr(tok), self.d, n * self.d, n, self.d,
                 self.p, ptr(self.diag), stream())
            self.fac, self.ld = tok, self.d              # stored as F (N x r)
        else:
            ldn = _pad4(n)
            k = _mat(self.p, n, n, dev)
            sgemm(0, 1, n, n, self.d, tok, self.d, n * self.d, tok, self.d, n * self.d, k, ldn, n * ldn, self.p,
                  tc=True)
            call("basd_extract_diag", ptr(k), n, ldn, n * ldn, self.p, ptr(self.diag), stream())
            self.fac = _mat(self.p, n, n, dev)           # stored as F^T (r x N)
            self.rank = torch.empty(self.p, dtype=torch.int32, device=dev)
            pivoted_cholesky(k, self.fac, CHOL_TOL, rank_out=self.rank)
            self.ld = ldn
        self.stride = self.fac.shape[1] * self.fac.shape[2]

    # operand descriptors for sgemm: (transpose flag, tensor, ld, stride)
    def ft_left(self):        # F^T (r x N) as the left operand
        return (1 if self.direct else 0), self.fac, self.ld, self.stride

    def f_right(self):        # F (N x r) as the right operand
        return (0 if self.direct else 1), self.fac, self.ld, self.stride

    def ft_right(self):       # F^T (r x N) as the right operand
        return (1 if self.direct else 0), self.fac, self.ld, self.stride


def procrustes_forward(students, teachers, stats: Stats, weights, n_student, with_grad: bool):
    dev = students[0].device
    e, l = len(students), len(teachers)
    b, n_t, d_t = teachers[0].shape
    d_s = students[0].shape[2]
    n = n_student
    tdt = teachers[0].dtype
    out_dtype = torch.bfloat16 if (tdt == torch.bfloat16 and n_t == n) else torch.float32
    aligned = torch.empty(e, b, n, d_t, dtype=out_dtype, device=dev)
    tptrs = (nat.C.c_void_p * l)(*[t.data_ptr() for t in teachers])
    call("basd_mix_interp", tptrs, l, e, ptr(weights), nat.dtype_code(teachers[0]), b, n_t, n, d_t,
         ptr(aligned), nat.dtype_code(aligned), stream())
    w = _f32(e, b, n, device=dev)
    totals = _f32(e, b, device=dev)
    call("basd_mix_rows", ptr(stats.rows), ptr(weights), e, l, b, stats.rows.shape[2], n, ptr(w), ptr(totals),
         stream())

    a = _f32(e, b, n, d_s, device=dev)
    bm = _f32(e, b, n, d_t, device=dev)
    for i, s in enumerate(students):
        call("basd_weighted_center", ptr(s), nat.dtype_code(s), n * d_s, ptr(w[i]), n, n, d_s,
             ptr(a[i]), n * d_s, b, stream())
    call("basd_weighted_center", ptr(aligned), nat.dtype_code(aligned), n * d_t, ptr(w), n, n, d_t,
         ptr(bm), n * d_t, e * b, stream())

    p = e * b
    # A side is direct when D <= N.  When exactly one side qualifies, the other one is taken direct
    # as well if its width is at most MIXED_DIRECT_RATIO * N: a Gram factor on one side hides the
    # singular directions below sqrt(eps) from BOTH gradients (measured: N = 196, D_s = 192,
    # D_t = 384 gives gradient cosine 0.997 with a Gram teacher side, 0.99999 with X = A^T B), at
    # the price of Jacobi rows of length D instead of N.
    direct_s, direct_t = d_s <= n, d_t <= n
    if direct_s != direct_t:
        wide = d_t if direct_s else d_s
        if wide <= MIXED_DIRECT_RATIO * n:
            direct_s = direct_t = True
    side_s = _Side(a.view(p, n, d_s), n, direct_s)
    side_t = _Side(bm.view(p, n, d_t), n, direct_t)
    # q = the side with fewer factor columns: the Jacobi sweep orthogonalises the r_q rows of
    # G = F_q^T F_p (length r_p >= r_q).  With equal widths (two Gram factors) the sweep implicitly
    # diagonalises G G^T = L_q^T K_p L_q: the closer K_p is to a multiple of the identity, the closer
    # this is to L_q^T L_q, one LR step of K_q and already nearly diagonal.  The side with the wider
    # tokens averages more terms per Gram entry and is the better conditioned one, so it becomes p
    # (fp32 emulation of the sweep on C2 draws: 10 sweeps / 133k rotations with q = teacher, 7-8
    # sweeps / 103k with q = student).  A teacher resampled from fewer tokens is rank deficient and
    # stays q (its zero rows are skipped by the rank-aware sweep).
    swap = side_t.r > side_s.r
    if (side_t.r == side_s.r and not side_s.direct and not side_t.direct and PROC_TIE_Q == "narrow"
            and d_s < d_t and n_t >= n):
        swap = True
    sq, sp = (side_s, side_t) if swap else (side_t, side_s)
    rq, rp = sq.r, sp.r
    lrq, lrp, ldn = _pad4(rq), _pad4(rp), _pad4(n)        # pitches (a Gram side's width is the token count)
    g = _mat(p, rq, rp, dev)
    ta, fa_, lda, sa = sq.ft_left()
    tb, fb_, ldb, sb = sp.f_right()
    sgemm(ta, tb, rq, rp, n, fa_, lda, sa, fb_, ldb, sb, g, lrp, rq * lrp, p, tc=True)     # G = F_q^T F_p
    g0 = g.clone()
    sweeps = torch.zeros(p, dtype=torch.int32, device=dev)
    # rows -> sigma_j p_j^T; rows of G beyond the rank of a Gram-side F_q are exact zeros
    jacobi_rows(g, sweeps_out=sweeps, tag="procrustes", tol=PROC_JACOBI_TOL, row_dims=sq.rank, cols=rp)
    rows_normalize(g, g, None, sort=False, square=False, rel_floor=ROW_FLOOR, cols=rp)
    pt = g                                                                # (rq, rp) unit rows
    rows2 = _mat(p, rq, rq, dev)
    sgemm(0, 1, rq, rq, rp, pt, lrp, rq * lrp, g0, lrp, rq * lrp, rows2, lrq, rq * lrq, p, tc=True)  # P^T G^T = S Q^T
    del g0
    # scaling exponents (in halves) of the own-vector rows, see basd_procrustes_rows_finish
    e_s = -1 if not side_t.direct else 0
    e_t = -1 if not side_s.direct else 0
    if side_s.direct and not side_t.direct:
        e_t = 1
    if side_t.direct and not side_s.direct:
        e_s = 1
    eq, ep = (e_s, e_t) if swap else (e_t, e_s)
    floor = PROC_SV_FLOOR_DIRECT if (side_s.direct and side_t.direct) else PROC_SV_FLOOR
    # the q vectors are derived (normalised rows of P^T G^T); the Gram-side operator they enter
    # (Y_p, the p side's gradient) divides by sigma and takes the higher floor
    floor_q = floor if sp.direct else max(floor, PROC_SV_FLOOR_DERIVED)
    sig = _f32(p, rq, device=dev)
    pic = _f32(p, rq, device=dev)
    nuc = _f32(p, device=dev)
    call("basd_procrustes_rows_finish", ptr(rows2), rq, lrq, rq * lrq, ptr(pt), rp, lrp, rq * lrp, p,
         floor, floor_q, eq, ep, ptr(sig), ptr(nuc), ptr(pic), stream())
    f = _f32(p, device=dev)
    m_a = m_b = g_a = g_b = gw = None
    if with_grad:
        img_q = _mat(p, rq, n, dev)                       # rows (F_q q_j)^T (scaled)
        img_p = _mat(p, rq, n, dev)                       # rows (F_p p_j)^T (scaled)
        tb, fb_, ldb, sb = sq.ft_right()
        sgemm(0, tb, rq, n, rq, rows2, lrq, rq * lrq, fb_, ldb, sb, img_q, ldn, rq * ldn, p, tc=True)
        tb, fb_, ldb, sb = sp.ft_right()
        sgemm(0, tb, rq, n, rp, pt, lrp, rq * lrp, fb_, ldb, sb, img_p, ldn, rq * ldn, p, tc=True)

        def side_operator(side, img_other, own, own_ld):
            """Gram side: Y = I_o^T I_o (N x N).  Direct side: T = I_o^T own (N x D)."""
            if side.direct:
                t = _f32(p, n, side.d, device=dev)
                sgemm(1, 0, n, side.d, rq, img_other, ldn, rq * ldn, own, own_ld, rq * own_ld, t,
                      side.d, n * side.d, p, tc=True)
                return None, t
            y = _mat(p, n, n, dev)
            sgemm(1, 0, n, n, rq, img_other, ldn, rq * ldn, img_other, ldn, rq * ldn, y, ldn, n * ldn, p, tc=True)
            return y, None

        y_p, t_p = side_operator(sp, img_q, pt, lrp)
        y_q, t_q = side_operator(sq, img_p, rows2, lrq)
        (m_a, g_a), (m_b, g_b) = ((y_q, t_q), (y_p, t_p)) if swap else ((y_p, t_p), (y_q, t_q))
        gw = _f32(e, b, n, device=dev)
        call("basd_procrustes_grad_prep", ptr(m_a), ptr(m_b), ptr(img_q), ptr(img_p), n, rq, ldn, rq * ldn,
             ldn, n * ldn, p, ptr(pic), ptr(nuc), ptr(side_s.diag), ptr(side_t.diag), ptr(w), ptr(totals),
             ptr(f), ptr(gw), 1, stream())
        if g_a is not None:
            call("basd_procrustes_direct_grad", ptr(side_s.tok), ptr(g_a), ptr(w), n, d_s, p, stream())
        if g_b is not None:
            call("basd_procrustes_direct_grad", ptr(side_t.tok), ptr(g_b), ptr(w), n, d_t, p, stream())
    else:
        call("basd_procrustes_grad_prep", None, None, None, None, n, rq, ldn, rq * ldn, ldn, n * ldn, p, None,
             ptr(nuc), ptr(side_s.diag), ptr(side_t.diag), ptr(w), ptr(totals), ptr(f), None, 0, stream())
    geo_terms = _f32(e, device=dev)
    geo = _f32((), device=dev)
    call("basd_geo_reduce", ptr(f), e, b, ptr(weights), e * l, ptr(geo_terms), ptr(geo), stream())
    return ProcrustesState(a, bm, m_a, m_b, g_a, g_b, gw, f, geo_terms, geo, aligned, w, sweeps)


def procrustes_backward(students, teachers, stats: Stats, pro: ProcrustesState,
                        grad_out: torch.Tensor, n_student: int, want_teacher_grad: bool = False):
    """Returns (direct dL/dS_i list in the students' dtype, dL/dweights (E,L) for the local
    shard, optional dL/d(aligned teacher tokens))."""
    dev = students[0].device
    e, l = len(students), len(teachers)
    b, n_t, d_t = teachers[0].shape
    d_s = students[0].shape[2]
    n = n_student
    ldn = _pad4(n)
    p, nn = e * b, n * ldn
    go = grad_out.detach().to(torch.float32).reshape(1).contiguous()
    scale = 1.0 / (e * b)                                # d geo / d f_{i,b}
    if pro.m_a is not None:
        sdt = students[0].dtype
        grad_s = torch.empty(e, b, n, d_s, dtype=sdt, device=dev)
        if all(s.dtype == sdt for s in students) and gemm_tc_ex(
                0, 0, n, d_s, n, pro.m_a, ldn, nn, pro.a, d_s, n * d_s, grad_s, d_s, n * d_s, p,
                alpha=scale, alpha_dev=go):
            outs = [grad_s[i] for i in range(e)]           # written in the token dtype by the epilogue
        else:
            grad_s = _f32(e, b, n, d_s, device=dev)
            sgemm(0, 0, n, d_s, n, pro.m_a, ldn, nn, pro.a, d_s, n * d_s, grad_s, d_s, n * d_s, p,
                  alpha=scale, alpha_dev=go, tc=True)
            outs = [_cast_like(grad_s[i], s) for i, s in enumerate(students)]
    else:
        g_a = pro.g_a.view(e, b, n, d_s)
        outs = []
        for i, s in enumerate(students):
            out = torch.empty(s.shape, dtype=s.dtype, device=dev)
            call("basd_scale_out", ptr(g_a[i]), ptr(out), nat.dtype_code(out), out.numel(), scale,
                 ptr(go), stream())
            outs.append(out)
    z = _f32(e, b, n, d_t, device=dev)
    if pro.m_b is not None:
        sgemm(0, 0, n, d_t, n, pro.m_b, ldn, nn, pro.bm, d_t, n * d_t, z, d_t, n * d_t, p,
              alpha=scale, alpha_dev=go, tc=True)
    else:
        call("basd_scale_out", ptr(pro.g_b), ptr(z), nat.F32, z.numel(), scale, ptr(go), stream())
    # mixing-weight gradients: one pass over the teacher stack
    slices = nat.load().basd_weight_grad_slices()
    partial = _f32(slices * l * e, device=dev)
    d_weights = _f32(e, l, device=dev)
    tptrs = (nat.C.c_void_p * l)(*[t.data_ptr() for t in teachers])
    call("basd_weight_grad", tptrs, l, e, ptr(z), ptr(pro.gw), ptr(stats.rows),
         nat.dtype_code(teachers[0]), b, n_t, stats.rows.shape[2], n, d_t, scale, ptr(go), ptr(partial),
         ptr(d_weights), stream())
    return outs, d_weights, (z if want_teacher_grad else None)


def _cast_like(src32: torch.Tensor, like: torch.Tensor) -> torch.Tensor:
    if like.dtype == torch.float32:
        return src32.view(like.shape)
    out = torch.empty(like.shape, dtype=like.dtype, device=like.device)
    call("basd_cast_out", ptr(src32), ptr(out), nat.dtype_code(out), out.numel(), stream())
    return out


def selector_backward(students, sel: SelectorState, proj_s, log_temps, d_weights: torch.Tensor,
                      group, world: int):
    """d_weights (E,L): dL/dweights summed over this rank's samples. Returns
    (list of dL/dS_i through the selector, dL/dlog_temps). Closed form of the SVD/eigh
    backward: only cross-block terms survive."""
    dev = students[0].device
    e, l = sel.weights.shape
    b, n, d_s = students[0].shape
    d_weights = d_weights.detach().to(torch.float32).contiguous()
    if world > 1:
        d_weights = d_weights.clone()
        _all_reduce(d_weights, group)
    d_dist = _f32(e, l, device=dev)
    d_logt = _f32(e, device=dev)
    call("basd_mix_weights_bwd", ptr(d_weights), ptr(sel.weights), ptr(sel.dist), ptr(log_temps), e,
         l, 1.0 / world, ptr(d_dist), ptr(d_logt), stream())
    dd = d_s * d_s
    uxt = sel.uxt.clone()
    call("basd_scale_rows_dsigma", ptr(uxt), ptr(sel.sig), ptr(sel.lam_t), ptr(sel.ranks),
         ptr(d_dist), d_s, e, l, stream())
    t1 = _f32(e * l, d_s, d_s, device=dev)
    sgemm(0, 1, d_s, d_s, d_s, sel.wfull, d_s, dd, sel.vxt, d_s, dd, t1, d_s, dd, e * l, tc=True)
    blk = _f32(e * l, d_s, d_s, device=dev)
    sgemm(0, 0, d_s, d_s, d_s, t1, d_s, dd, uxt, d_s, dd, blk, d_s, dd, e * l, tc=True)
    omega = _f32(e, d_s, d_s, device=dev)
    call("basd_omega_accumulate", ptr(blk), ptr(sel.lam_s), ptr(sel.ranks), d_s, e, l, ptr(omega),
         stream())
    t2 = _f32(e, d_s, d_s, device=dev)
    sgemm(1, 0, d_s, d_s, d_s, sel.vt_s, d_s, dd, omega, d_s, dd, t2, d_s, dd, e, tc=True)     # V Omega
    d_gram = _f32(e, d_s, d_s, device=dev)
    sgemm(0, 0, d_s, d_s, d_s, t2, d_s, dd, sel.vt_s, d_s, dd, d_gram, d_s, dd, e, tc=True)    # . V^T
    w_sym = t2
    call("basd_symmetrize_add", ptr(d_gram), d_s, ptr(w_sym), e, stream())
    t3 = d_gram
    sgemm(1, 0, d_s, d_s, d_s, proj_s, d_s, 0, w_sym, d_s, dd, t3, d_s, dd, e, tc=True)        # P^T W
    w_prime = omega
    sgemm(0, 0, d_s, d_s, d_s, t3, d_s, dd, proj_s, d_s, 0, w_prime, d_s, dd, e, tc=True)      # . P
    outs = []
    shift_dot = _f32(e, d_s, device=dev)                  # mu_i^T W'_i: the centring folded into the epilogue
    for i, s in enumerate(students):
        sgemm(0, 0, 1, d_s, d_s, sel.mean_s[i], d_s, 0, w_prime[i], d_s, 0, shift_dot[i], d_s, 0, 1)
        out = torch.empty(s.shape, dtype=s.dtype, device=dev)
        if gemm_tc_ex(0, 0, b * n, d_s, d_s, s, d_s, 0, w_prime[i], d_s, 0, out, d_s, 0, 1,
                      col_sub=shift_dot[i]):
            outs.append(out)
            continue
        g32 = _f32(b, n, d_s, device=dev)
        sgemm(0, 0, b * n, d_s, d_s, s, d_s, 0, w_prime[i], d_s, 0, g32, d_s, 0, 1,
              a_shift=sel.mean_s[i])
        outs.append(_cast_like(g32, s))
    return outs, d_logt
