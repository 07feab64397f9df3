"""One profiled fwd+bwd step of the C2 workload between cudaProfilerStart/Stop (for ncu
--profile-from-start off).  Usage: python tools/profile_step.py [workload] [batch] [backbone|spectral]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import basd_b200.synthetic as syn
import bench

key = sys.argv[1] if len(sys.argv) > 1 else "c2"
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 256
work = syn.scaled(syn.WORKLOADS[key], batch)
dev = torch.device("cuda", 0)
mod = bench.build_module(work, dev)
features = sys.argv[3] if len(sys.argv) > 3 else "backbone"
if features == "backbone":
    from basd_b200 import backbone_features as bf
    logits, targets, st, te, at = bf.workload_inputs(key, work, seed=0, device=dev)
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
else:
    logits, targets, st, te, at = syn.make_inputs_fast(work, seed=0, device=dev)
st = {k: v.requires_grad_(True) for k, v in st.items()}
logits.requires_grad_(True)
for _ in range(2):
    bench.one_step(mod, logits, targets, st, te, at)
torch.cuda.synchronize()
torch.cuda.profiler.start()
loss = bench.one_step(mod, logits, targets, st, te, at)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("loss", float(loss))
