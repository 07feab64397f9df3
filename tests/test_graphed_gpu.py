"""The whole step (forward + hand-written backward + CE + UW-SO) captured in ONE CUDA graph must give the
eager module's loss and gradients bit for bit (same kernels, same order), for fresh data on every replay."""
import types

import pytest
import torch

import basd_b200.synthetic as syn
from basd_b200.graphed import GraphedBASDLoss
from tests import _cases as cs

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("key,batch", [("c1", 16), ("c2", 4)])
def test_graph_replay_equals_eager(key, batch):
    work = cs.workload(key, batch)
    temps = [0.3, 0.541, 0.8, 1.2]
    eager = cs.build_cuda_module(work, temps)
    graphed = GraphedBASDLoss(cs.build_cuda_module(work, temps))
    for seed in (1, 2, 3):                                    # capture on the first, replay on fresh data
        logits, targets, st, te, at = syn.make_inputs(work, seed=seed)
        dev = lambda d: {k: v.cuda() for k, v in d.items()}
        te_d, at_d, tg = dev(te), dev(at), targets.cuda()
        out = []
        for mod in (eager, graphed):
            st_d = {k: v.cuda().requires_grad_(True) for k, v in st.items()}
            lg = logits.cuda().requires_grad_(True)
            mod.layer_selector.log_temperatures.grad = None
            loss = mod(lg, tg, st_d, te_d, at_d)
            (2.0 * loss).backward()                           # a non-unit upstream gradient
            torch.cuda.synchronize()
            out.append((loss.detach().clone(), lg.grad.clone(), {k: v.grad.clone() for k, v in st_d.items()},
                        mod.layer_selector.log_temperatures.grad.clone()))
        (l0, g0, s0, t0), (l1, g1, s1, t1) = out
        assert torch.equal(l0, l1), (seed, float(l0), float(l1))
        assert torch.equal(g0, g1)
        assert torch.equal(t0, t1)
        for k in eager.token_layers:
            assert torch.equal(s0[k], s1[k]), (seed, k)
    assert len(graphed._captured) == 1
