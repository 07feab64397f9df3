"""Per-kernel checks of the C-ABI entry points on a real GPU (fp64 torch as the yardstick)."""
import ctypes

import pytest
import torch

pytestmark = pytest.mark.gpu

DEV = "cuda"


def _eng():
    from basd_b200 import _engine as eng
    return eng


def _nat():
    from basd_b200 import _native as nat
    return nat


@pytest.mark.parametrize("ta,tb", [(0, 0), (0, 1), (1, 0), (1, 1)])
def test_sgemm_all_layouts(ta, tb):
    eng = _eng()
    torch.manual_seed(0)
    b, m, n, k = 3, 70, 45, 133
    a = torch.randn(b, *( (k, m) if ta else (m, k)), device=DEV)
    bb = torch.randn(b, *((n, k) if tb else (k, n)), device=DEV)
    c = torch.randn(b, m, n, device=DEV)
    c0 = c.clone()
    alpha_dev = torch.tensor([0.5], device=DEV)
    eng.sgemm(ta, tb, m, n, k, a, a.shape[2], a[0].numel(), bb, bb.shape[2], bb[0].numel(), c, n,
              m * n, b, alpha=2.0, alpha_dev=alpha_dev, beta=0.25)
    ao = a.transpose(1, 2) if ta else a
    bo = bb.transpose(1, 2) if tb else bb
    ref = (ao.double() @ bo.double()) * 1.0 + 0.25 * c0.double()
    assert (c.double() - ref).abs().max() < 1e-4


@pytest.mark.parametrize("ta,tb", [(0, 0), (0, 1), (1, 0), (1, 1)])
@pytest.mark.parametrize("m,n,k", [(196, 196, 196), (196, 196, 384), (196, 768, 196), (128, 128, 256),
                                   (68, 52, 36), (256, 320, 100), (192, 196, 256)])
def test_gemm_tc3_matches_fp64_at_fp32_accuracy(ta, tb, m, n, k):
    """tcgen05 3xTF32 batched GEMM: fp32-level accuracy (not TF32's 1e-3) against fp64,
    and no worse than 4x the SIMT fp32 kernel's own error on the same inputs."""
    eng = _eng()
    torch.manual_seed(3)
    b = 5
    a = torch.randn(b, *((k, m) if ta else (m, k)), device=DEV) * torch.logspace(0, -3, m, device=DEV).view(
        *((1, 1, m) if ta else (1, m, 1)))
    bb = torch.randn(b, *((n, k) if tb else (k, n)), device=DEV)
    alpha_dev = torch.tensor([0.5], device=DEV)
    c = torch.full((b, m, n), float("nan"), device=DEV)
    c_simt = torch.empty(b, m, n, device=DEV)
    args = (ta, tb, m, n, k, a, a.shape[2], a[0].numel(), bb, bb.shape[2], bb[0].numel())
    eng.sgemm(*args, c, n, m * n, b, alpha=2.0, alpha_dev=alpha_dev, tc=True)
    eng.sgemm(*args, c_simt, n, m * n, b, alpha=2.0, alpha_dev=alpha_dev, tc=False)
    ao = a.transpose(1, 2) if ta else a
    bo = bb.transpose(1, 2) if tb else bb
    ref = ao.double() @ bo.double()
    scale = (ao.double().abs() @ bo.double().abs())           # sum |a||b| per output
    err = ((c.double() - ref).abs() / scale).max()
    err_simt = ((c_simt.double() - ref).abs() / scale).max()
    assert torch.isfinite(c).all()
    # worst case of the split: both operands carry 2^-22 after rounding lo, plus the dropped lo*lo
    assert err < 1.5e-6, (float(err), float(err_simt))
    assert err < 8 * err_simt + 1e-7, (float(err), float(err_simt))


@pytest.mark.parametrize("a_dtype,c_dtype", [(torch.bfloat16, torch.bfloat16), (torch.bfloat16, torch.float32),
                                             (torch.float32, torch.bfloat16)])
def test_gemm_tc3_extended_bf16_operand_column_shift_bf16_output(a_dtype, c_dtype):
    """(A - 1 mu^T) B with bf16 tokens as the A operand, the centring applied in the epilogue
    and the result stored in the token dtype (selector backward, layer_selector.py:90-91)."""
    eng = _eng()
    torch.manual_seed(4)
    m, n, k = 1000, 384, 384
    a = (torch.randn(m, k, device=DEV) + 0.7).to(a_dtype)
    w = torch.randn(k, n, device=DEV) / k ** 0.5
    mu = a.float().mean(dim=0)
    shift = (mu.double() @ w.double()).float()
    c = torch.full((m, n), float("nan"), device=DEV, dtype=c_dtype)
    alpha_dev = torch.tensor([0.25], device=DEV)
    assert eng.gemm_tc_ex(0, 0, m, n, k, a, k, 0, w, n, 0, c, n, 0, 1, alpha=2.0, alpha_dev=alpha_dev,
                          col_sub=shift)
    ref = 0.5 * ((a.double() - mu.double()) @ w.double())
    tol = 2e-2 if c_dtype == torch.bfloat16 else 2e-5
    assert torch.isfinite(c.float()).all()
    assert (c.double() - ref).abs().max() < tol * ref.abs().max()


def test_sgemm_bf16_a_with_shift_and_shared_operand():
    eng = _eng()
    torch.manual_seed(1)
    m, n, k = 200, 96, 96
    a = torch.randn(m, k, device=DEV).bfloat16()
    shift = torch.randn(k, device=DEV)
    w = torch.randn(k, n, device=DEV)
    c = torch.zeros(m, n, device=DEV)
    eng.sgemm(0, 0, m, n, k, a, k, 0, w, n, 0, c, n, 0, 1, a_shift=shift)
    ref = (a.double() - shift.double()) @ w.double()
    assert (c.double() - ref).abs().max() < 1e-3


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("rows,d", [(1024, 192), (4 * 196, 384), (3001, 200), (6272, 768), (1000, 256),
                                    (50, 128)])
def test_token_gram(dtype, rows, d):
    eng = _eng()
    torch.manual_seed(2)
    x = (torch.randn(rows, d, device=DEV) + 0.3).to(dtype)
    gram = torch.empty(d, d, device=DEV)
    col = torch.empty(d, device=DEV)
    eng.token_gram(x.view(1, rows, d), gram, col)
    ref = x.double().T @ x.double()
    assert (gram.double() - ref).abs().max() / ref.abs().max() < 1e-5
    assert (col.double() - x.double().sum(0)).abs().max() < 1e-2 * max(1.0, float(x.double().sum(0).abs().max())) * 1e-3
    assert (gram - gram.T).abs().max() == 0


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("rows,d", [(4 * 196, 384), (3001, 200), (50176, 384), (6272, 768), (50, 128)])
def test_token_gram_in_a_mean_shifted_frame(dtype, rows, d):
    """Large token means (|mu|^2 = 100 x the variance): the Gram and column sums of x - mu0 must come out with
    the accuracy of the COVARIANCE scale, not of M |mu|^2 -- the shift is worked into the accumulation
    (one negated UMMA per 64-row stage on the tensor-core path), the tokens are not rounded."""
    eng = _eng()
    torch.manual_seed(5)
    mu = 10.0 * torch.randn(d, device=DEV)
    x = (torch.randn(rows, d, device=DEV) * torch.logspace(0, -1, d, device=DEV) + mu).to(dtype)
    mu0 = x[:min(rows, 4096)].float().mean(0).to(torch.bfloat16).float()
    gram = torch.empty(d, d, device=DEV)
    col = torch.empty(d, device=DEV)
    eng.token_gram(x.view(1, rows, d), gram, col, mu0)
    xs = x.double() - mu0.double()
    ref = xs.T @ xs
    scale = float(ref.diagonal().max())                  # covariance scale (the shifted frame's largest entry)
    err = float((gram.double() - ref).abs().max())
    raw_scale = float((x.double().T @ x.double()).abs().max())
    print(f"rows {rows} d {d} {dtype}: max err / cov scale {err / scale:.2e} (uncentred scale is {raw_scale / scale:.0f}x)")
    assert err / scale < 2e-5
    assert (col.double() - xs.sum(0)).abs().max() < 1e-3 * max(1.0, float(xs.sum(0).abs().max()))
    assert (gram - gram.T).abs().max() == 0


@pytest.mark.parametrize("n,rank", [(64, 64), (196, 195), (196, 48), (384, 384)])
def test_pivoted_cholesky(n, rank):
    eng = _eng()
    torch.manual_seed(3)
    batch = 5
    a = torch.randn(batch, n, rank, device=DEV) * torch.logspace(0, -1.5, rank, device=DEV)
    k = a @ a.transpose(1, 2)
    k = 0.5 * (k + k.transpose(1, 2))
    work = k.clone()
    lt = torch.empty_like(k)
    ranks = torch.zeros(batch, dtype=torch.int32, device=DEV)
    eng.pivoted_cholesky(work, lt, 1e-6, rank_out=ranks)
    rec = lt.transpose(1, 2) @ lt
    err = (rec - k).abs().amax(dim=(1, 2)) / k.abs().amax(dim=(1, 2))
    assert err.max() < 2e-5, err
    # a rank+1-th pivot at noise level (just above rel_tol) is legitimate for fp32 Grams
    assert ranks.max() <= min(rank + 1, n) and ranks.min() >= min(rank, n) - 3, ranks


@pytest.mark.parametrize("n", [64, 131, 196, 384])
def test_jacobi_rows_orthogonalises(n):
    eng = _eng()
    torch.manual_seed(4)
    batch = 4
    ld = (n + 3) // 4 * 4
    g = torch.zeros(batch, n, ld, device=DEV)
    g[:, :, :n] = torch.randn(batch, n, n, device=DEV) * torch.logspace(0, -2, n, device=DEV)
    ref_sv = torch.linalg.svdvals(g[:, :, :n].double())
    sweeps = torch.zeros(batch, dtype=torch.int32, device=DEV)
    dims = torch.full((batch,), n, dtype=torch.int32, device=DEV)
    from basd_b200._native import call, ptr, stream
    call("basd_jacobi_rows", ptr(g), n, ld, ld, n * ld, batch, ptr(dims), 1e-6, 18, ptr(sweeps), stream())
    rows = g[:, :, :n].double()
    gram = rows @ rows.transpose(1, 2)
    nrm = gram.diagonal(dim1=1, dim2=2).sqrt()
    off = gram / (nrm.unsqueeze(1) * nrm.unsqueeze(2)).clamp(min=1e-30)
    off = off - torch.diag_embed(off.diagonal(dim1=1, dim2=2))
    assert off.abs().max() < 2e-5, (float(off.abs().max()), sweeps)
    sv = nrm.sort(dim=1, descending=True).values
    assert ((sv - ref_sv).abs().max(dim=1).values / ref_sv[:, 0]).max() < 2e-4
    assert sweeps.max() < 18, sweeps


@pytest.mark.parametrize("d", [192, 384])
def test_sym_eig(d):
    eng = _eng()
    torch.manual_seed(5)
    x = torch.randn(3, 4 * d, d, device=DEV) * torch.logspace(0, -2, d, device=DEV)
    k = x.transpose(1, 2) @ x
    k = 0.5 * (k + k.transpose(1, 2))
    lam, vt = eng.sym_eig(k)
    ref = torch.linalg.eigvalsh(k.double()).flip(1)
    assert ((lam.double() - ref).abs().max(dim=1).values / ref[:, 0]).max() < 2e-6
    eye = torch.eye(d, device=DEV)
    assert (vt @ vt.transpose(1, 2) - eye).abs().max() < 5e-5
    resid = (vt @ k - lam.unsqueeze(2) * vt).abs().max() / ref[:, 0].max()
    assert resid < 5e-5


def test_mp_rank_kernel_matches_rule():
    nat = _nat()
    from basd_b200._native import call, ptr, stream
    from oracle import kernel_model as km
    torch.manual_seed(6)
    d, rows = 96, 1000
    lam = (torch.rand(5, d, device=DEV) * torch.logspace(1, -2, d, device=DEV))
    ranks = torch.zeros(5, dtype=torch.int32, device=DEV)
    call("basd_mp_rank", ptr(lam), d, rows, d - 1, ptr(ranks), None, 5, stream())
    for i in range(5):
        want = km.mp_rank_from_spectrum(lam[i].cpu().sort(descending=True).values, rows, d - 1)
        assert int(ranks[i]) == want


@pytest.mark.parametrize("has_cls,dtype", [(True, torch.float32), (False, torch.float32), (True, torch.bfloat16)])
def test_attn_rows(has_cls, dtype):
    from basd_b200._native import call, ptr, stream, dtype_code
    torch.manual_seed(7)
    b, h, side = 5, 3, 50
    attn = torch.softmax(torch.randn(b, h, side, side, device=DEV), -1).to(dtype)
    n_tok = side - 1 if has_cls else side
    rows = torch.empty(b, n_tok, device=DEV)
    call("basd_attn_rows", ptr(attn), dtype_code(attn), b, h, side, side, int(has_cls), ptr(rows), stream())
    a = attn.float()
    ref = a[:, :, 0, 1:].mean(1) if has_cls else a.mean((1, 2))
    assert (rows - ref).abs().max() < 1e-6


@pytest.mark.parametrize("dtype,n_src,n_dst", [(torch.bfloat16, 196, 196), (torch.bfloat16, 49, 196),
                                               (torch.float32, 64, 64), (torch.float32, 256, 196)])
def test_mix_interp_and_weight_grad(dtype, n_src, n_dst):
    from basd_b200._native import call, ptr, stream, dtype_code, load
    from oracle import kernel_model as km
    torch.manual_seed(8)
    l, e, b, d = 5, 4, 3, 64
    layers = [torch.randn(b, n_src, d, device=DEV).to(dtype) for _ in range(l)]
    w = torch.softmax(torch.randn(e, l, device=DEV), 1)
    out_dtype = torch.bfloat16 if (dtype == torch.bfloat16 and n_src == n_dst) else torch.float32
    out = torch.empty(e, b, n_dst, d, dtype=out_dtype, device=DEV)
    ptrs = (ctypes.c_void_p * l)(*[t.data_ptr() for t in layers])
    call("basd_mix_interp", ptrs, l, e, ptr(w), dtype_code(layers[0]), b, n_src, n_dst, d, ptr(out),
         dtype_code(out), stream())
    stack = torch.stack(layers).cpu()
    for i in range(e):
        ref = km.mix_and_align(w[i].cpu(), stack, n_dst)
        tol = 2e-2 if out_dtype == torch.bfloat16 else 1e-5
        assert (out[i].float().cpu() - ref).abs().max() < tol
    # weight gradient: <Z_i, resample(T_l)> + scale * <gw_i, resample(rows_l)>
    z = torch.randn(e, b, n_dst, d, device=DEV)
    gw = torch.randn(e, b, n_dst, device=DEV)
    rows = torch.rand(l, b, n_src, device=DEV)
    slices = load().basd_weight_grad_slices()
    partial = torch.empty(slices * l * e, device=DEV)
    dw = torch.empty(e, l, device=DEV)
    sc = torch.tensor([0.7], device=DEV)
    call("basd_weight_grad", ptrs, l, e, ptr(z), ptr(gw), ptr(rows), dtype_code(layers[0]), b, n_src,
         n_src, n_dst, d, 0.5, ptr(sc), ptr(partial), ptr(dw), stream())
    lo, hi, fr = km.interp_taps(n_src, n_dst)
    lo, hi, fr = lo.to(DEV), hi.to(DEV), fr.to(DEV)
    for j in range(l):
        t = layers[j].double()
        r = rows[j].double()
        if n_src != n_dst:
            t = t[:, lo] * (1 - fr.double()).view(1, -1, 1) + t[:, hi] * fr.double().view(1, -1, 1)
            r = r[:, lo] * (1 - fr.double()) + r[:, hi] * fr.double()
        for i in range(e):
            ref = (z[i].double() * t).sum() + 0.35 * (gw[i].double() * r).sum()
            assert abs(float(dw[i, j]) - float(ref)) < 1e-3 * (1 + abs(float(ref)))


def test_mix_rows():
    from basd_b200._native import call, ptr, stream
    from oracle import kernel_model as km
    torch.manual_seed(9)
    l, e, b, n_src, n_dst = 4, 3, 5, 49, 196
    rows = torch.rand(l, b, n_src, device=DEV)
    w = torch.softmax(torch.randn(e, l, device=DEV), 1)
    out = torch.empty(e, b, n_dst, device=DEV)
    tot = torch.empty(e, b, device=DEV)
    call("basd_mix_rows", ptr(rows), ptr(w), e, l, b, n_src, n_dst, ptr(out), ptr(tot), stream())
    for i in range(e):
        ref, t = km.mix_importance(w[i].cpu(), rows.cpu(), n_dst)
        assert (out[i].cpu() - ref).abs().max() < 1e-6
        assert (tot[i].cpu() - t[:, 0]).abs().max() < 1e-4


@pytest.mark.parametrize("n,d_s,d_t,rank_t", [(64, 192, 384, 64), (196, 384, 768, 196), (196, 384, 512, 40)])
def test_procrustes_stage_value_and_gradients(n, d_s, d_t, rank_t):
    """procrustes_forward/backward on one mixing layer against the model and autograd."""
    eng = _eng()
    from oracle import kernel_model as km, ref_port as rp
    torch.manual_seed(10)
    b = 3
    s = torch.randn(b, n, d_s) * torch.logspace(0, -1.5, d_s) + 0.2
    base = torch.randn(b, rank_t, d_t) * torch.logspace(0, -1.5, d_t)
    t = base if rank_t == n else torch.nn.functional.interpolate(
        base.transpose(1, 2), size=n, mode="linear", align_corners=False).transpose(1, 2)
    rows = torch.rand(1, b, n) + 0.1
    stats = eng.Stats(None, None, None, None, rows.to(DEV))
    w1 = torch.ones(1, 1, device=DEV)
    sd, td = s.to(DEV), t.contiguous().to(DEV)
    pro = eng.procrustes_forward([sd], [td], stats, w1, n, True)
    go = torch.tensor(1.0, device=DEV)
    grads, dw, z = eng.procrustes_backward([sd], [td], stats, pro, go, n, want_teacher_grad=True)
    torch.cuda.synchronize()
    w = rows[0] / rows[0].sum(1, keepdim=True)
    for i in range(b):
        f, gs, gt, gwn = km.procrustes_sample(s[i], t[i], w[i])
        assert abs(float(pro.f[i]) - float(f)) / abs(float(f)) < 2e-4, (float(pro.f[i]), float(f))
        from tests._cases import cosine
        assert cosine(grads[0][i].cpu(), gs / b) > 0.9995
        assert cosine(z[0, i].cpu(), gt / b) > 0.999
    # and against the reference formulation by autograd (value)
    sg = s.clone().requires_grad_(True)
    fake_attn = torch.zeros(b, 1, n + 1, n + 1)
    fake_attn[:, 0, 0, 1:] = rows[0]
    ref = rp.procrustes_loss(sg, t, fake_attn, True)
    ref.backward()
    assert abs(float(pro.geo) - float(ref)) / abs(float(ref)) < 1e-3
    assert cosine(grads[0].cpu(), sg.grad) > 0.999


def test_mp_rank_secular_matches_direct_spectrum():
    """Rank from (centred eigendecomposition + secular equation) == rank from the uncentred spectrum."""
    eng = _eng()
    from basd_b200._native import call, ptr, stream
    from oracle import kernel_model as km
    torch.manual_seed(11)
    d, rows, layers = 192, 2048, 6
    x = torch.randn(layers, rows, d, device=DEV) * torch.logspace(0, -2, d, device=DEV) + \
        torch.randn(layers, 1, d, device=DEV) * 0.7          # sizeable mean: centring matters
    gram = x.transpose(1, 2) @ x
    gram = 0.5 * (gram + gram.transpose(1, 2))
    col = x.sum(1)
    kc = gram - col.unsqueeze(2) * col.unsqueeze(1) / rows
    lam, vt = eng.sym_eig(kc.contiguous())
    y = torch.einsum("lij,lj->li", vt, col).contiguous()
    ranks = torch.zeros(layers, dtype=torch.int32, device=DEV)
    edges = torch.zeros(layers, 3, device=DEV)
    call("basd_mp_rank_secular", ptr(lam), ptr(y), d, rows, d - 1, ptr(ranks), ptr(edges), layers, stream())
    for i in range(layers):
        spec = torch.linalg.eigvalsh(gram[i].double().cpu()).flip(0).float()
        want = km.mp_rank_from_spectrum(spec, rows, d - 1)
        n = spec.numel()
        med = spec.flip(0)[(n - 1) // 2]
        assert abs(float(edges[i, 0]) - float(med)) / float(med) < 1e-4
        assert int(ranks[i]) == want or edges[i, 2] == 1


@pytest.mark.parametrize("d_out,d_in,batch", [(384, 768, 3), (192, 192, 2), (100, 230, 2)])
def test_rotate_stats_f64_rounds_once(d_out, d_in, batch):
    """sym(P G P^T) - (P c)(P c)^T / M in double precision: each fp32 output equals the fp64 result
    rounded once (the fp32 product it replaces is off by ~1e-6 lambda_max on a graded spectrum)."""
    nat = _nat()
    torch.manual_seed(1)
    rows = 5000
    x = torch.randn(batch, rows, d_in, dtype=torch.float64) * torch.logspace(0, -3, d_in, dtype=torch.float64)
    x = x + 0.5
    gram = (x.transpose(1, 2) @ x).float().to(DEV)
    col = x.sum(dim=1).float().to(DEV)
    proj = torch.linalg.qr(torch.randn(d_in, d_in, dtype=torch.float64))[0][:d_out].contiguous().float().to(DEV)
    k = torch.empty(batch, d_out, d_out, device=DEV)
    chat = torch.empty(batch, d_out, device=DEV)
    ws = torch.empty(nat.load().basd_rotate_stats_f64_workspace_bytes(d_out, d_in, batch) // 8,
                     dtype=torch.float64, device=DEV)
    nat.call("basd_rotate_stats_f64", nat.ptr(proj), d_out, d_in, nat.ptr(gram), nat.ptr(col), batch,
             1.0 / rows, nat.ptr(ws), nat.ptr(k), nat.ptr(chat), nat.stream())
    p64 = proj.double()
    c64 = col.double() @ p64.T
    g64 = p64 @ gram.double() @ p64.T
    ref = 0.5 * (g64 + g64.transpose(1, 2)) - c64.unsqueeze(2) * c64.unsqueeze(1) / rows
    scale = ref.abs().amax(dim=(1, 2), keepdim=True)
    # one rounding: relative error <= 2^-24 of each entry, plus fp64 summation-order noise
    assert ((k.double() - ref).abs() <= 6.1e-8 * ref.abs() + 1e-11 * scale).all()
    assert (k - k.transpose(1, 2)).abs().max() == 0
    assert ((chat.double() - c64).abs() <= 6.1e-8 * c64.abs() + 1e-11 * c64.abs().max()).all()


def test_merge_shifted_stats_matches_the_common_frame():
    """Statistics accumulated by W ranks in their own mean-shifted frames, merged by
    basd_merge_shifted_stats, equal the statistics of the concatenated rows in the common frame
    mean_r mu0_r (center.cu; the data-parallel contract of _engine.statistics)."""
    from basd_b200._native import call, ptr, stream
    torch.manual_seed(11)
    world, tensors, d, rows = 3, 2, 96, 500
    x = (torch.randn(world, tensors, rows, d, device=DEV, dtype=torch.float64) * torch.logspace(0, -1, d, device=DEV)
         + 5.0 * torch.randn(tensors, 1, d, device=DEV, dtype=torch.float64))
    mu_r = (x[:, :, :64].mean(2) + 0.05 * torch.randn(world, tensors, d, device=DEV)).float()      # rough, per rank
    xs = x - mu_r.double().unsqueeze(2)
    gram = torch.einsum("wtnd,wtne->tde", xs, xs).float().contiguous()        # what the all-reduce sums
    d_slots = xs.sum(2).float().contiguous()
    col = torch.empty(tensors, d, device=DEV)
    mu0 = torch.empty(tensors, d, device=DEV)
    call("basd_merge_shifted_stats", ptr(gram), ptr(col), ptr(mu0), ptr(d_slots), ptr(mu_r.contiguous()), world,
         tensors, d, rows, stream())
    mu = mu_r.double().mean(0)
    xc = x - mu.unsqueeze(0).unsqueeze(2)
    ref_g = torch.einsum("wtnd,wtne->tde", xc, xc)
    ref_c = xc.sum((0, 2))
    assert (mu0.double() - mu).abs().max() < 1e-5          # fp32 mean of values ~10
    assert (gram.double() - ref_g).abs().max() / ref_g.abs().max() < 1e-5
    assert (col.double() - ref_c).abs().max() < 1e-3 * max(1.0, float(ref_c.abs().max()))
    assert (gram - gram.transpose(1, 2)).abs().max() == 0
