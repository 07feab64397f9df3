"""Host enqueue time vs GPU time per step (is the step launch-bound?).  python tools/host_time.py [workload] [batch]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import basd_b200.synthetic as syn
import bench
key = sys.argv[1] if len(sys.argv) > 1 else "c2"
work = syn.scaled(syn.WORKLOADS[key], int(sys.argv[2]) if len(sys.argv) > 2 else 256)
dev = torch.device("cuda", 0)
mod = bench.build_module(work, dev)
logits, targets, st, te, at = syn.make_inputs_fast(work, seed=0, device=dev)
st = {k: v.requires_grad_(True) for k, v in st.items()}
logits.requires_grad_(True)
for _ in range(3):
    bench.one_step(mod, logits, targets, st, te, at)
torch.cuda.synchronize()
for trial in range(12):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    bench.one_step(mod, logits, targets, st, te, at)
    e1.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"trial {trial}: host enqueue {1e3*(t1-t0):.1f} ms, gpu {e0.elapsed_time(e1):.1f} ms, wall {1e3*(t2-t0):.1f} ms")
print(torch.cuda.memory_stats()["num_alloc_retries"], "alloc retries;", torch.cuda.memory_stats()["num_device_alloc"], "device allocs;",
      torch.cuda.memory_stats()["num_device_free"], "device frees")
