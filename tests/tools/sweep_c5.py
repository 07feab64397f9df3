"""C5 microbench sweep (BASELINE.json configs[4]): loss fwd+bwd over token counts, widths, teacher
depths and batch sizes on one GPU.  For every point: CUDA-event time of the step (4 warm-ups, median of 5
timed steps), samples/s, and -- with --parity -- loss / gradient agreement with the CPU oracle on a
batch-4 slice of the same distribution.  Prints one JSON line per point.
Usage: python tests/tools/sweep_c5.py [--parity] [--quick]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import basd_b200.synthetic as syn
from tests import _cases as cs

POINTS = [  # (N, D_s, D_t, L_t, B)
    (64, 192, 192, 12, 1024), (64, 192, 384, 12, 256), (196, 192, 384, 12, 256),
    (196, 384, 768, 12, 32), (196, 384, 768, 12, 1024), (196, 384, 1024, 24, 256),
    (256, 384, 768, 12, 256), (576, 384, 768, 12, 64), (1024, 192, 384, 12, 32),
    (1024, 384, 768, 12, 32),
]


def work_for(n, ds, dt, lt, b, dtype=torch.bfloat16):
    return syn.Workload(f"c5_n{n}_ds{ds}_dt{dt}_l{lt}_b{b}", b, n, n, ds, dt, lt, ds // 64, True, dtype)


def time_point(work):
    dev = torch.device("cuda")
    mod = cs.build_cuda_module(work)
    logits, targets, st, te, at = syn.make_inputs_fast(work, seed=0, device=dev)
    st = {k: v.requires_grad_(True) for k, v in st.items()}

    def step():
        for v in st.values():
            v.grad = None
        loss = mod(logits, targets, st, te, at)
        loss.backward()
        return loss

    for _ in range(4):
        step()
    torch.cuda.synchronize()
    times = []
    for _ in range(5):                      # per-step events, median: one allocator / host hiccup
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()                           # does not move the reported number
        loss = step()
        b.record()
        torch.cuda.synchronize()
        times.append(a.elapsed_time(b))
    ms = sorted(times)[len(times) // 2]
    return ms, float(loss.detach()), torch.cuda.max_memory_allocated() / 2 ** 30


def parity_point(n, ds, dt, lt):
    work = work_for(n, ds, dt, lt, 4, torch.float32)
    if 4 * n < ds:
        return None
    inputs = syn.make_inputs(work, seed=1)
    ref = cs.run_oracle(work, inputs, None)
    got = cs.run_cuda(work, inputs, None)
    rel = abs(float(got["loss"]) - float(ref["loss"])) / abs(float(ref["loss"]))
    cos = min(cs.cosine(got["grad_students"][l], ref["grad_students"][l]) for l in ref["layers"])
    werr = float((got["weights"] - ref["weights"]).abs().max())
    return dict(loss_rel=rel, min_grad_cos=cos, weights_err=werr, ranks_equal=got["ranks"] == ref["ranks"])


if __name__ == "__main__":
    parity = "--parity" in sys.argv
    pts = POINTS[:4] if "--quick" in sys.argv else POINTS
    if "--only" in sys.argv:
        want = [tuple(int(x) for x in a.split(",")) for a in sys.argv[sys.argv.index("--only") + 1:]]
        pts = [q for q in POINTS if q in want] + [q for q in want if q not in POINTS]
    for n, ds, dt, lt, b in pts:
        rec = dict(n=n, d_s=ds, d_t=dt, layers=lt, batch=b)
        try:
            torch.cuda.reset_peak_memory_stats()
            ms, loss, gib = time_point(work_for(n, ds, dt, lt, b))
            rec.update(ms_per_step=round(ms, 2), samples_per_s=round(b / ms * 1e3, 1), loss=loss,
                       peak_gib=round(gib, 2))
            if parity:
                t0 = time.time()
                rec["parity_b4_fp32"] = parity_point(n, ds, dt, lt)
                rec["oracle_seconds"] = round(time.time() - t0, 1)
        except Exception as err:  # noqa: BLE001 - the sweep reports what does not run
            rec["error"] = f"{type(err).__name__}: {err}"[:200]
        print(json.dumps(rec), flush=True)
        torch.cuda.empty_cache()
