"""Replicates the pytest flow outside pytest: oracle first, then CUDA, compared with the model."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import basd_b200.synthetic as syn
from oracle import kernel_model as km
from tests import _cases as cs

key, batch, seed = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
temps = None if len(sys.argv) < 5 else [float(x) for x in sys.argv[4].split(",")]
order = sys.argv[5] if len(sys.argv) > 5 else "oracle_first"
work = cs.workload(key, batch)
inputs = syn.make_inputs(work, seed=seed)
logits, targets, st, te, at = inputs
if order == "oracle_first":
    ref = cs.run_oracle(work, inputs, temps)
    got = cs.run_cuda(work, inputs, temps)
else:
    got = cs.run_cuda(work, inputs, temps)
    ref = cs.run_oracle(work, inputs, temps)
proj_s, proj_t, logt0 = cs.selector_state(work)
logt = torch.tensor(temps) if temps else logt0
model = km.full_step_model(logits, targets, st, te, at, layers=ref["layers"], proj_s=proj_s, proj_t=proj_t,
                           log_temps=logt, n_student=work.n_student, has_cls=work.has_cls,
                           criterion=cs.criterion(work))
print("loss cuda/oracle/model", float(got["loss"]), float(ref["loss"]), float(model["loss"]))
print("logt grads cuda ", got["grad_log_temps"].tolist())
print("logt grads orcl ", ref["grad_log_temps"].tolist())
print("logt grads model", model["grad_log_temps"].tolist())
for l in ref["layers"]:
    g, r, m = got["grad_students"][l], ref["grad_students"][l], model["grad_students"][l]
    print(f"layer {l}: cos cuda-oracle {cs.cosine(g, r):.6f} cuda-model {cs.cosine(g, m):.6f} model-oracle {cs.cosine(m, r):.6f}"
          f" norms {float(g.norm()):.5f} {float(r.norm()):.5f} {float(m.norm()):.5f}")
