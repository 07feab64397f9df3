"""torchrun script (>= 2 GPUs): the data-parallel loss equals the single-process loss on the
concatenated batch.  python -m torch.distributed.run --nproc-per-node 2 tools/dp_parity.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import torch.distributed as dist

import basd_b200.synthetic as syn
from tests import _cases as cs

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
key, local_batch = (sys.argv[1], int(sys.argv[2])) if len(sys.argv) > 2 else ("c1", 8)
temps = [0.3, 0.541, 0.8, 1.2]
work = cs.workload(key, local_batch)
inputs = syn.make_inputs(work, seed=6, batch_offset=rank * local_batch)
got = cs.run_cuda(work, inputs, temps, device=dev)
glob = got["module"].last["global_loss"].cpu()
ok = True
if rank == 0:
    full_work = cs.workload(key, local_batch * world)
    full_inputs = syn.make_inputs(full_work, seed=6)
    mod = cs.build_cuda_module(full_work, temps, dev, sync_stats=False)
    logits, targets, st, te, at = full_inputs
    st_d = {k: v.to(dev).requires_grad_(True) for k, v in st.items()}
    loss = mod(logits.to(dev).requires_grad_(True), targets.to(dev), st_d, {k: v.to(dev) for k, v in te.items()},
               {k: v.to(dev) for k, v in at.items()})
    loss.backward()
    torch.cuda.synchronize()
    ref_w = mod.last["weights"].cpu()
    ref_ranks = [mod.layer_selector.subspace_ranks[k] for k in sorted(mod.layer_selector.subspace_ranks.keys())]
    print("ranks equal", got["ranks"] == ref_ranks, "weights max diff", float((got["weights"] - ref_w).abs().max()))
    print("global loss", float(glob), "single-process loss", float(loss))
    ok &= got["ranks"] == ref_ranks and float((got["weights"] - ref_w).abs().max()) < 1e-5
    ok &= abs(float(glob) - float(loss)) / abs(float(loss)) < 1e-5
    for l in mod.token_layers:
        g_dp = got["grad_students"][l] / world
        g_ref = st_d[l].grad.float().cpu()[:local_batch]
        c = cs.cosine(g_dp, g_ref)
        rel = float((g_dp - g_ref).norm() / g_ref.norm())
        print(f"layer {l}: grad cosine {c:.7f} rel {rel:.2e}")
        ok &= c > 0.9999 and rel < 2e-2
    g1, g2 = got["grad_log_temps"], mod.layer_selector.log_temperatures.grad.cpu()
    print("log_temperatures grad", g1.tolist(), g2.tolist())
    ok &= cs.cosine(g1, g2) > 0.9999
    print("DP PARITY", "OK" if ok else "FAILED")
flag = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
dist.destroy_process_group()
sys.exit(0 if int(flag.item()) == 1 else 1)
