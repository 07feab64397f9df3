"""C4 (DeiT-B <- ViT-L/16: D_s = 768, D_t = 1024, 24 teacher layers mixed) through the CUDA path
against the live oracle at a batch the oracle finishes in seconds.  This is the only parity case
that reaches the 768-wide selector kernels (16-CTA cluster Jacobi and Cholesky, the k x k SVD size
window) -- the other workloads stop at 384."""
import pytest
import torch

import basd_b200.synthetic as syn
from tests import _cases as cs
from tests.test_loss_parity_gpu import COS_TOL, LOSS_TOL, W_TOL, _check_selector_at_kernel_rank, _ranks_ok

pytestmark = pytest.mark.gpu


def test_c4_against_live_oracle():
    work = cs.workload("c4", 8)                    # 8 x 196 = 1,568 token rows >= D_s = 768
    inputs = syn.make_inputs(work, seed=11)
    ref = cs.run_oracle(work, inputs)
    got = cs.run_cuda(work, inputs)
    print("c4 loss", float(got["loss"]), float(ref["loss"]), "geo", float(got["geo"]), float(ref["geo"]),
          "ranks", got["ranks"], ref["ranks"])
    assert _ranks_ok(got["ranks"], ref["ranks"], got["module"])
    ranks_equal = got["ranks"] == ref["ranks"]
    if ranks_equal:
        assert (got["weights"] - ref["weights"]).abs().max() < W_TOL
    else:
        _check_selector_at_kernel_rank(work, inputs, None, got)
    assert abs(float(got["geo"]) - float(ref["geo"])) / abs(float(ref["geo"])) < LOSS_TOL
    assert abs(float(got["loss"]) - float(ref["loss"])) / abs(float(ref["loss"])) < LOSS_TOL
    for layer in ref["layers"]:
        c = cs.cosine(got["grad_students"][layer], ref["grad_students"][layer])
        print("  layer", layer, "grad cosine", c)
        assert c > COS_TOL
    if ranks_equal:
        assert cs.cosine(got["grad_log_temps"], ref["grad_log_temps"]) > COS_TOL
    assert torch.allclose(got["grad_logits"], ref["grad_logits"], atol=1e-6, rtol=1e-3)


def test_selector_forward_returns_the_reference_shaped_dicts():
    """Secondary entry of SURVEY 8(b): GrassmannianLayerSelector.forward -> (mixed tokens, mixed attention
    maps) keyed by student layer (layer_selector.py:116-152), checked against the oracle's mixing weights."""
    work = cs.workload("c1", 16)
    inputs = syn.make_inputs(work, seed=4)
    ref = cs.run_oracle(work, inputs)
    logits, targets, st, te, at = inputs
    mod = cs.build_cuda_module(work)
    sel = mod.layer_selector
    st_d = {k: v.cuda() for k, v in st.items()}
    te_d = {k: v.cuda() for k, v in te.items()}
    at_d = {k: v.cuda() for k, v in at.items()}
    with torch.no_grad():
        mixed_tok, mixed_att = sel(st_d, te_d, at_d, mod.token_layers)
    assert sorted(mixed_tok) == sorted(mixed_att) == list(mod.token_layers)
    keys = sorted(te)
    tok = torch.stack([te[k].float() for k in keys])
    att = torch.stack([at[k].float() for k in keys])
    for i, layer in enumerate(mod.token_layers):
        w = ref["weights"][i]
        want_tok = (w.view(-1, 1, 1, 1) * tok).sum(0)
        want_att = (w.view(-1, 1, 1, 1, 1) * att).sum(0)
        assert mixed_tok[layer].shape == te[keys[0]].shape and mixed_att[layer].shape == at[keys[0]].shape
        assert (mixed_tok[layer].float().cpu() - want_tok).abs().max() < 1e-3 * want_tok.abs().max()
        assert (mixed_att[layer].float().cpu() - want_att).abs().max() < 1e-3 * want_att.abs().max()
    # the importance-row form of the mixed attention: what relational.py:22-34 reduces the mixed map to
    weights, step = sel.mixing_weights(st_d, te_d, at_d, mod.token_layers, has_cls=True)
    rows = sel.mixed_importance(weights, step).cpu()
    for i, layer in enumerate(mod.token_layers):
        want = (ref["weights"][i].view(-1, 1, 1, 1, 1) * att).sum(0)[:, :, 0, 1:].mean(1)
        want = want / want.sum(-1, keepdim=True)
        assert (rows[i] - want).abs().max() < 1e-5
    assert [sel.subspace_ranks[k] for k in keys] == ref["ranks"] or _ranks_ok(
        [sel.subspace_ranks[k] for k in keys], ref["ranks"], mod)
