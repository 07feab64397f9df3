"""GPU-box debugging aid: C4 (batch 8) student-gradient cosine against the oracle with individual
round-2 changes switched off.   python tests/tools/debug_c4_toggle.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import basd_b200.synthetic as syn
from basd_b200 import _engine as eng
from basd_b200 import _native as nat
from tests import _cases as cs

work = cs.workload("c4", 8)
inputs = syn.make_inputs(work, seed=11)
ref = cs.run_oracle(work, inputs)


def run(tag):
    got = cs.run_cuda(work, inputs)
    cos = [round(cs.cosine(got["grad_students"][l], ref["grad_students"][l]), 6) for l in ref["layers"]]
    print(f"{tag:28s} loss rel {abs(float(got['loss']) - float(ref['loss'])) / abs(float(ref['loss'])):.2e} "
          f"weights {float((got['weights'] - ref['weights']).abs().max()):.2e} cosines {cos}", flush=True)


run("as built")
orig_complete = eng.complete_null_space
eng.complete_null_space = lambda vt: None
run("no completion")
eng.complete_null_space = orig_complete
orig_shift = eng.GRAM_SHIFT
eng.GRAM_SHIFT = 0.0
run("no diagonal shift")
eng.GRAM_SHIFT = orig_shift
orig_call = eng.call


def no_mean(name, *a):
    if name == "basd_rough_means":
        rc = orig_call(name, *a)
        mu0_ptr, count, d = a[6], a[1], a[4]
        import ctypes
        torch.cuda.current_stream().synchronize()
        ctypes.CDLL("libcudart.so").cudaMemset(ctypes.c_void_p(mu0_ptr), 0, ctypes.c_size_t(count * d * 4))
        return rc
    return orig_call(name, *a)


eng.call = no_mean
run("no mean shift (mu0 = 0)")
eng.call = orig_call
