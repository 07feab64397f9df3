"""the loss inside a DistributedDataParallel-wrapped student step, wired exactly as the
reference's trainer wires it (trainer.py:64-84 builds the loss next to the wrapped student and adds its
parameters to the optimizer; :140-159 is the step) -- student hooks through the DDP wrapper, teacher
capture, BASDLoss, backward through the student, optimizer step.  The parameter gradients of the whole
student are compared with the CPU oracle's on the same images and weights.  World size 1 here (the
driver's box has one GPU); the N-rank equivalence with the oracle is `dp_parity` in bench.py's N > 1 runs
and tests/test_dp_gpu.py."""
import copy
import types

import pytest
import torch
import torch.distributed as dist

from basd_b200 import backbone_features as bb
from basd_b200 import capture
from basd_b200.losses import BASDLoss
from oracle import ref_port as rp
from tests import _cases as cs

pytestmark = pytest.mark.gpu


def _token_hooks(model, layers):
    """CPU twin of capture.extract_student for the oracle side."""
    got, hooks = {}, []
    for i in layers:
        hooks.append(model.blocks[i].register_forward_hook(
            lambda m, inp, out, i=i: got.__setitem__(i, out[:, 1:, :])))
    return got, hooks


def test_loss_inside_a_ddp_wrapped_student_step():
    created = False
    if not dist.is_initialized():
        dist.init_process_group("nccl", init_method="tcp://127.0.0.1:29533", rank=0, world_size=1,
                                device_id=torch.device("cuda", 0))
        created = True
    try:
        batch, classes, img, patch = 16, 100, 32, 4
        student = bb.vit("deit_tiny", 7, classes, img, patch, student=True).train()
        teacher_net = bb.vit("deit_small", 8, classes, img, patch)
        cpu_student = copy.deepcopy(student)
        gen = torch.Generator().manual_seed(3)
        images = torch.randn(batch, 3, img, img, generator=gen)
        targets = torch.randint(0, classes, (batch,), generator=gen)

        ddp = torch.nn.parallel.DistributedDataParallel(student.cuda(), device_ids=[0])
        teacher = bb.vit_teacher(teacher_net.cuda())
        crit = torch.nn.CrossEntropyLoss(label_smoothing=1.0 / classes)
        torch.manual_seed(cs.SELECTOR_SEED)
        loss_mod = BASDLoss(crit, 192, 384, 12, (img // patch) ** 2,
                            config=types.SimpleNamespace(num_extraction_points=4),
                            teacher_has_cls_token=True, process_group=dist.group.WORLD).cuda()
        opt = torch.optim.SGD(ddp.parameters(), lr=1e-3)
        opt.add_param_group({"params": list(loss_mod.parameters())})          # trainer.py:74-76
        paths = [f"blocks.{i}" for i in range(12)]
        x = images.cuda()
        logits, st = capture.extract_student(ddp, x, loss_mod.token_layers, layer_paths=paths,
                                             has_cls_token=True)
        te, at = capture.extract_intermediates(teacher, x)                  # (B, N) importance rows
        loss = loss_mod(logits, targets.cuda(), st, te, at)
        loss.backward()
        before = loss_mod.layer_selector.log_temperatures.detach().clone()
        opt.step()
        torch.cuda.synchronize()
        assert not torch.equal(before, loss_mod.layer_selector.log_temperatures.detach())

        # ---- the same step through the CPU oracle (full attention maps, as the reference captures them)
        te_full, at_full = capture.extract_intermediates(teacher, x, full_maps=True)
        got, hooks = _token_hooks(cpu_student, loss_mod.token_layers)
        ref_logits = cpu_student(images)
        for h in hooks:
            h.remove()
        torch.manual_seed(cs.SELECTOR_SEED)
        proj_s, proj_t, logt = rp.make_selector_state(4, 192, 384)
        logt.requires_grad_(True)
        ref_loss, _ = rp.basd_forward(
            ref_logits, targets, got, {k: v.float().cpu() for k, v in te_full.items()},
            {k: v.float().cpu() for k, v in at_full.items()}, layers=loss_mod.token_layers, proj_s=proj_s,
            proj_t=proj_t, log_temps=logt, n_student_tokens=(img // patch) ** 2, has_cls=True, criterion=crit)
        ref_loss.backward()
        assert abs(float(loss) - float(ref_loss)) / abs(float(ref_loss)) < 1e-3
        mine, theirs = [], []
        for (name, p), (_, q) in zip(ddp.module.named_parameters(), cpu_student.named_parameters()):
            assert (p.grad is None) == (q.grad is None), name
            if p.grad is not None:
                assert torch.isfinite(p.grad).all(), name
                mine.append(p.grad.flatten().cpu())
                theirs.append(q.grad.flatten())
        c = cs.cosine(torch.cat(mine), torch.cat(theirs))
        print("student parameter gradient cosine", c, "loss", float(loss), float(ref_loss))
        assert c > 0.999
        assert cs.cosine(loss_mod.layer_selector.log_temperatures.grad.cpu(), logt.grad) > 0.999
    finally:
        if created:
            dist.destroy_process_group()
