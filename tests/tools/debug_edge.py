"""GPU-box debugging aid for the edge-case workloads of tests/test_edge_cases_gpu.py:
stage-by-stage distance of the CUDA path from the CPU kernel model."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from tests.test_edge_cases_gpu import _work
from tests.tools import debug_stages as ds

CASES = {
    "nocls_down": (dict(n_student=196, n_teacher=256, d_student=192, d_teacher=384, has_cls=False,
                        num_points=2, batch=4), (0.4, 0.9)),
    "n256": (dict(n_student=256, n_teacher=256, d_student=128, d_teacher=128, batch=4, teacher_layers=2), None),
}

if __name__ == "__main__":
    for name in sys.argv[1:] or CASES:
        kw, temps = CASES[name]
        print("=" * 20, name)
        ds.main(work=_work(**kw), seed=0, temps=temps, uniform_attn=False)
