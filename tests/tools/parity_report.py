"""GPU-box aid: margins of the CUDA path against the live CPU oracle on the parity workloads
(loss rel. error, max mixing-weight error, min gradient cosine, Procrustes sweeps)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import basd_b200.synthetic as syn
from tests import _cases as cs

CASES = [("c1", 16, 5, [0.2, 0.6, 1.0, 1.4]), ("c2", 8, 2, None), ("c3", 8, 3, None)]

if __name__ == "__main__":
    for key, batch, seed, temps in CASES:
        work = cs.workload(key, batch)
        inputs = syn.make_inputs(work, seed=seed)
        ref = cs.run_oracle(work, inputs, temps)
        got = cs.run_cuda(work, inputs, temps)
        rel = abs(float(got["loss"]) - float(ref["loss"])) / abs(float(ref["loss"]))
        werr = float((got["weights"] - ref["weights"]).abs().max())
        cos = min(cs.cosine(got["grad_students"][l], ref["grad_students"][l]) for l in ref["layers"])
        sw = got["module"].layer_selector.last_state.sweeps["procrustes"].float()
        print(f"{key} b{batch}: loss rel {rel:.2e}  weights {werr:.2e}  min grad cos {cos:.6f}  "
              f"ranks equal {got['ranks'] == ref['ranks']}"
              + (f"  sweeps mean {float(sw.mean()):.2f}" if sw is not None else ""))
