import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import basd_b200.synthetic as syn
import bench
work = syn.scaled(syn.WORKLOADS["c2"], 256)
dev = torch.device("cuda", 0)
mod = bench.build_module(work, dev)
logits, targets, st, te, at = syn.make_inputs_fast(work, seed=0, device=dev)
st = {k: v.requires_grad_(True) for k, v in st.items()}
logits.requires_grad_(True)
for _ in range(3):
    bench.one_step(mod, logits, targets, st, te, at)
torch.cuda.synchronize()
for mode in ("no sampler", "nvml sampler", "no sampler, gc off"):
    sampler = None
    if mode == "nvml sampler":
        sampler = bench.ClockSampler(0); sampler.start()
    if mode.endswith("gc off"):
        import gc; gc.disable()
    n = 40
    marks = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
    marks[0].record()
    for i in range(n):
        bench.one_step(mod, logits, targets, st, te, at)
        marks[i + 1].record()
    torch.cuda.synchronize()
    if sampler: sampler.stop_flag = True; sampler.join(timeout=2)
    ts = [marks[i].elapsed_time(marks[i + 1]) for i in range(n)]
    print(mode, "median %.2f max %.2f" % (sorted(ts)[n // 2], max(ts)), "slow steps:", [round(t, 1) for t in ts if t > 70])
