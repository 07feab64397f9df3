#!/usr/bin/env python
"""BASD loss fwd+bwd benchmark (BASELINE.json metric) -- see DESIGN.md §6.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2] [--batch 256]
  python bench.py --impl reference ...      # the reference algorithm on the host CPU cores

One "step" = one forward + backward of the loss module over one batch of synthetic,
ImageNet-shaped DeiT-S<-DeiT-B features (C2: B=256/GPU, N=196, D 384<-768, 12 teacher
layers, bf16 tokens, fp32 attention maps).  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
import types

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
os.environ.setdefault("NCCL_DEBUG", "WARN")      # keep stdout to the one JSON line
# load every kernel image up front: with lazy loading the first launch of each kernel variant
# stalls for milliseconds, wherever in the run it happens to fall
os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import basd_b200.synthetic as syn  # noqa: E402

METRIC = "basd_loss_fwd_bwd_samples_per_sec"
UNIT = "samples/s"


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm=p["hbm_gbs"], bf16=p["bf16_tflops"], bf16_sustained=p["bf16_tflops_sustained"],
                    sm_max_mhz=p["sm_max_mhz"], source="measured")
    return dict(hbm=6650.0, bf16=1590.0, bf16_sustained=1400.0, sm_max_mhz=1965.0, source="fallback")


ORIG_AFFINITY = os.sched_getaffinity(0)


def _cpulist(text):
    ids = set()
    for part in text.strip().split(","):
        if part:
            lo, _, hi = part.partition("-")
            ids.update(range(int(lo), int(hi or lo) + 1))
    return ids


def bind_to_gpu_numa_node(index, world=1):
    """Pins this process to the CPUs local to GPU `index` BEFORE any pinned host buffer is allocated,
    so that cudaHostAlloc places the staging pages on the GPU's own NUMA node (eight ranks streaming
    1.1 GB per step each through a remote socket is what broke the end-to-end scaling in round 1).
    sysfs first (numa_node / local_cpulist of the GPU's PCI device); where the platform hides that
    (numa_node = -1) but exposes several nodes, ranks are dealt to the nodes in order (GPUs 0..W/2-1 on
    the first socket is the usual board layout).  Returns a short description for the JSON line."""
    info = {"numa_node": None}
    try:
        nodes = sorted(int(d[4:]) for d in os.listdir("/sys/devices/system/node") if d.startswith("node") and d[4:].isdigit())
    except OSError:
        nodes = []
    info["nodes_visible"] = len(nodes)
    try:
        prop = torch.cuda.get_device_properties(index)
        bus = f"{prop.pci_domain_id:04x}:{prop.pci_bus_id:02x}:{prop.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
        ids = _cpulist(open(f"/sys/bus/pci/devices/{bus}/local_cpulist").read())
        how = "sysfs"
        if node < 0 and len(nodes) > 1:
            node = nodes[min(len(nodes) - 1, index * len(nodes) // max(1, world))]
            ids = _cpulist(open(f"/sys/devices/system/node/node{node}/cpulist").read())
            how = "dealt by rank (platform reports numa_node -1)"
        info.update(numa_node=node, how=how)
        allowed = os.sched_getaffinity(0)
        if node >= 0 and ids & allowed:
            os.sched_setaffinity(0, ids & allowed)
            info["cpus_bound"] = len(ids & allowed)
        else:
            info["cpus_bound"] = 0
    except Exception as exc:                                 # no sysfs entry, restricted cpuset, ...
        info["note"] = f"not bound ({type(exc).__name__})"
    return info


def measured_traffic():
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of the kernels that have a
    byte model, from the committed `ncu --set full` captures (profiles/r2_traffic.json)."""
    path = os.path.join(ROOT, "profiles", "r2_traffic.json")
    if os.path.exists(path):
        return json.load(open(path)).get("per_launch_dram_bytes", {})
    return {}


class ClockSampler(threading.Thread):
    """Samples SM clocks and throttle reasons DURING the timed region through in-process NVML
    (a polling `nvidia-smi` subprocess perturbs the launch thread); falls back to one
    `nvidia-smi` query if NVML is unavailable."""
    REASONS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
               "sw_power_cap": 0x4}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.mhz, self.mask, self.max_mhz, self.stop_flag = index, [], 0, None, False
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(visible.split(",")[index]) if visible and visible.split(",")[index].isdigit() else index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nvml = None

    def run(self):
        if self.nvml is None:
            return
        n = self.nvml
        while not self.stop_flag:
            try:
                self.mhz.append(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
                get = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
                    n.nvmlDeviceGetCurrentClocksThrottleReasons
                self.mask |= int(get(self.handle))
            except Exception:
                pass
            time.sleep(0.02)

    def summary(self):
        if self.nvml is None or not self.mhz:
            try:
                out = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm",
                                      "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=10).stdout.split(",")
                return {"sm_mhz": int(out[0]), "sm_max_mhz": int(out[1]), "reasons": ["unsampled"]}
            except Exception:
                return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        mhz = sorted(self.mhz)
        reasons = [k for k, bit in self.REASONS.items() if self.mask & bit]
        return {"sm_mhz": mhz[len(mhz) // 2], "sm_max_mhz": self.max_mhz, "reasons": reasons,
                "samples": len(mhz)}


def make_inputs(args, work, seed, device):
    """Synthetic step inputs.  "backbone": random images through random-init DeiT/ViT/ResNet-50
    backbones, tokens and attention captured as the reference's trainer does
    (basd_b200/backbone_features.py).  "spectral": hand-made token spectra (basd_b200/synthetic.py)."""
    if args.features == "backbone":
        from basd_b200 import backbone_features as bf
        out = bf.workload_inputs(args.workload, work, seed=seed, device=device)
        if torch.device(device).type == "cuda":
            torch.cuda.synchronize()
            torch.cuda.empty_cache()                      # the backbones' activations are not part of the step
        return out
    return syn.make_inputs_fast(work, seed=seed, device=device)


def data_label(args):
    return ("synthetic (random images through random-init backbones)" if args.features == "backbone"
            else "synthetic (hand-made token spectra)")


def build_module(work, device, impl="b200"):
    crit = torch.nn.CrossEntropyLoss(label_smoothing=1.0 / work.num_classes)
    cfg = types.SimpleNamespace(num_extraction_points=work.num_points)
    torch.manual_seed(0)
    if impl == "b200":
        from basd_b200.losses import BASDLoss
        mod = BASDLoss(crit, work.d_student, work.d_teacher, work.student_depth, work.n_student,
                       config=cfg, teacher_has_cls_token=work.has_cls)
    else:
        from oracle.ref_port import OracleBASD
        mod = OracleBASD(crit, work.d_student, work.d_teacher, work.student_depth, work.n_student,
                         config=cfg, teacher_has_cls_token=work.has_cls)
    return mod.to(device)


def one_step(mod, logits, targets, st, te, at):
    for v in st.values():
        v.grad = None
    logits.grad = None
    mod.layer_selector.log_temperatures.grad = None
    loss = mod(logits, targets, st, te, at)
    loss.backward()
    return loss


# ------------------------------------------------------------------ roofline leg
def flop_model(work, ranks_mean):
    """Algorithmic work of the formulation the kernels actually use."""
    b, n, ds, dt, l, e = work.batch, work.n_student, work.d_student, work.d_teacher, work.teacher_layers, work.num_points
    m_s, m_t = b * work.n_student, b * work.n_teacher
    gram = 2.0 * m_t * dt * dt * l / 2 + 2.0 * m_s * ds * ds * e / 2     # symmetric token-space Grams
    tok = 2.0 if work.token_dtype == torch.bfloat16 else 4.0
    out_b = tok if work.n_teacher == n else 4.0                          # resampled tokens are written in fp32
    mix_bytes = l * b * work.n_teacher * dt * tok + e * b * n * dt * out_b
    # dL/dweights: the teacher stack once + the upstream gradient (E,B,N,D_t) fp32 once
    wgrad_bytes = l * b * work.n_teacher * dt * tok + e * b * n * dt * 4.0
    # pivoted Cholesky: n^3 / 3 multiply-adds per factorisation (left-looking dot products), 2 flops each
    chol_flops = 2.0 * (2 * e * b * n ** 3 / 3.0 + (l + e) * ds ** 3 / 3.0)
    return dict(gram_flops=gram, mix_bytes=mix_bytes, wgrad_bytes=wgrad_bytes, chol_flops=chol_flops)


def roofline(work, mod, args5, pk, clocks):
    """Per-entry-point CUDA-event timing of one extra (untimed) step; returns the roofline
    record of the dominant kernel plus a per-kernel table."""
    from basd_b200 import _native as nat
    from basd_b200 import _engine as eng
    one_step(mod, *args5)
    torch.cuda.synchronize()
    eng.jacobi_log = []
    nat.gemm_flops = {}
    nat.start_timeline()
    one_step(mod, *args5)
    tl = nat.stop_timeline()
    jlog, eng.jacobi_log = eng.jacobi_log, None
    if os.environ.get("BASD_TIMELINE"):
        with open(os.environ["BASD_TIMELINE"], "w") as fh:
            for name, ms in tl:
                fh.write(f"{name},{ms:.4f}\n")
    agg = {}
    for name, ms in tl:
        agg.setdefault(name, [0.0, 0])
        agg[name][0] += ms
        agg[name][1] += 1
    total = sum(v[0] for v in agg.values())
    table = sorted(((k, v[0], v[1]) for k, v in agg.items()), key=lambda x: -x[1])
    fm = flop_model(work, None)
    gemm_flops = dict(nat.gemm_flops)
    sm_mhz = clocks.get("sm_mhz") or pk["sm_max_mhz"]
    fp32_peak = 148 * 128 * 2 * sm_mhz * 1e6 / 1e12
    # Jacobi: algorithmic flops of the launches of this step = pair visits x one dot product (2m)
    # + rotations actually applied x two rotated rows (6m); sweeps/rotation counts come from the device
    jac_ms = [ms for name, ms in tl if name.startswith("basd_jacobi_rows")]
    jac = []
    for (tag, n, m, dims, sweeps, rot, row_dims), ms in zip(jlog, jac_ms):
        sw = sweeps.double().cpu()
        # active rows: square sub-problem size (dims), factor rank (row_dims) or n; row length: dims or m
        kk = (dims.double().cpu() if dims is not None else
              row_dims.double().cpu().clamp(max=n) if row_dims is not None else torch.full_like(sw, float(n)))
        ln = dims.double().cpu() if dims is not None else torch.full_like(sw, float(m))
        visits = float((sw * kk * (kk - 1) / 2).sum())
        flops = float((sw * kk * (kk - 1) / 2 * 2 * ln).sum() + (rot.double().cpu() * 6 * ln).sum())
        jac.append({"kernel": f"basd_jacobi_rows[{tag}]", "ms": round(ms, 4), "problems": int(sw.numel()),
                    "n": int(kk.max()), "sweeps_mean": round(float(sw.mean()), 2), "pair_visits": visits,
                    "rotations": float(rot.sum()), "bound": "fp32", "achieved": flops / (ms * 1e-3) / 1e12,
                    "peak": fp32_peak, "unit": "TFLOP/s", "frac": flops / (ms * 1e-3) / 1e12 / fp32_peak,
                    "share": round(ms / total, 4)})
    records = []
    for name, ms, cnt in table:
        rec = {"kernel": name, "ms": round(ms, 4), "launches": cnt, "share": round(ms / total, 4)}
        if name in ("basd_token_gram_simt", "basd_token_gram_tc"):
            rec.update(bound="tensor" if name.endswith("tc") else "fp32", achieved=fm["gram_flops"] / (ms * 1e-3) / 1e12,
                       peak=pk["bf16_sustained"] if name.endswith("tc") else fp32_peak, unit="TFLOP/s")
        elif name == "basd_mix_interp":
            rec.update(bound="hbm", achieved=fm["mix_bytes"] / (ms * 1e-3) / 1e9, peak=pk["hbm"], unit="GB/s")
        elif name == "basd_weight_grad":
            rec.update(bound="hbm", achieved=fm["wgrad_bytes"] / (ms * 1e-3) / 1e9, peak=pk["hbm"], unit="GB/s")
        elif name == "basd_pivoted_cholesky" and work.d_student > work.n_student and work.d_teacher > work.n_student:
            rec.update(bound="fp32", achieved=fm["chol_flops"] / (ms * 1e-3) / 1e12, peak=fp32_peak, unit="TFLOP/s")
        elif name in gemm_flops:
            # 3xTF32: three tensor-core MMAs per fp32-equivalent product; the roofline counts the fp32-equivalent
            # flops against the measured dense bf16 rate / 2 (TF32 issues at half the bf16 rate) / 3
            rec.update(bound="tensor", achieved=gemm_flops[name] / (ms * 1e-3) / 1e12,
                       peak=pk["bf16_sustained"] / 6.0, unit="TFLOP/s(fp32-equivalent, 3xTF32)")
        if "achieved" in rec:
            rec["frac"] = rec["achieved"] / rec["peak"]
        records.append(rec)
    records = sorted(records + jac, key=lambda r: -r["ms"])
    return records, total, fp32_peak


def run_b200(args):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    numa = bind_to_gpu_numa_node(local, world)      # before the first pinned allocation
    if world > 1:
        import datetime
        dist.init_process_group("nccl", device_id=device, timeout=datetime.timedelta(seconds=300))
    parity = dp_parity(args, rank, world, device) if world > 1 and not args.no_dp_parity else None
    work = syn.scaled(syn.WORKLOADS[args.workload], args.batch)
    mod = build_module(work, device)
    eager_mod = mod
    if args.graph:                       # fwd+bwd replayed from one CUDA graph (single-GPU; see graphed.py)
        if world > 1:
            raise SystemExit("--graph is single-GPU only (the statistics all-reduces are not captured)")
        from basd_b200.graphed import GraphedBASDLoss
        mod = GraphedBASDLoss(mod)
    logits, targets, st, te, at = make_inputs(args, work, rank, device)
    st = {k: v.requires_grad_(True) for k, v in st.items()}
    logits.requires_grad_(True)
    args5 = (logits, targets, st, te, at)
    from basd_b200 import _native as nat

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    launches_eager = 0
    if args.graph:                         # the graph replays the launches of one eager step: count those
        n0 = nat.launch_count
        one_step(eager_mod, *args5)
        launches_eager = nat.launch_count - n0
    sampler = ClockSampler(local)          # NVML init and thread start-up stay outside the timed region
    sampler.start()
    loss = None
    for _ in range(args.warmup):
        loss = one_step(mod, *args5)       # keep `loss` alive exactly like the timed loop does, so
    barrier()                              # the caching allocator reaches its steady state here
    sampler.mhz.clear()
    sampler.mask = 0
    launches0 = nat.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    marks = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    e0.record()
    marks[0].record()
    for i in range(args.steps):
        loss = one_step(mod, *args5)
        marks[i + 1].record()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    per_step = [marks[i].elapsed_time(marks[i + 1]) for i in range(args.steps)]
    launches = nat.launch_count - launches0
    sampler.stop_flag = True
    sampler.join(timeout=2)
    clocks = sampler.summary()
    t = torch.tensor([ms], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    ms_per_step = ms / args.steps
    value = work.batch * world / (ms_per_step * 1e-3)

    # ---- end-to-end leg: host (pinned) inputs -> device copy -> fwd+bwd -> loss back on host
    host = {
        "st": {k: v.detach().cpu().pin_memory() for k, v in st.items()},
        "te": {k: v.cpu().pin_memory() for k, v in te.items()},
        # the loss reads only the CLS row (mean over heads is done on device): copy just that
        "at": {k: (v[:, :, 0, :].contiguous() if work.has_cls else v).cpu().pin_memory() for k, v in at.items()},
        "logits": logits.detach().cpu().pin_memory(), "targets": targets.cpu().pin_memory(),
    }
    h2d = sum(v.numel() * v.element_size() for d in (host["st"], host["te"], host["at"]) for v in d.values())
    h2d += host["logits"].numel() * 4 + host["targets"].numel() * 8

    copy_stream = torch.cuda.Stream(device=device)
    # two preallocated device input sets: step i uploads into set i % 2 while set (i-1) % 2 computes
    slots = []
    for _ in range(2):
        slots.append({
            "st": {k: torch.empty_like(v, device=device) for k, v in host["st"].items()},
            "te": {k: torch.empty_like(v, device=device) for k, v in host["te"].items()},
            "at": {k: torch.empty_like(v, device=device) for k, v in host["at"].items()},
            "lg": torch.empty_like(host["logits"], device=device),
            "tg": torch.empty_like(host["targets"], device=device),
            "free": None,
        })

    def upload(i):
        """Enqueues step i's host->device copies on the copy stream."""
        slot = slots[i % 2]
        with torch.cuda.stream(copy_stream):
            if slot["free"] is not None:
                copy_stream.wait_event(slot["free"])          # the step that used this set is done
            for grp in ("st", "te", "at"):
                for k, v in host[grp].items():
                    slot[grp][k].copy_(v, non_blocking=True)
            slot["lg"].copy_(host["logits"], non_blocking=True)
            slot["tg"].copy_(host["targets"], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return slot, ev

    loss_host = [torch.empty((), dtype=torch.float32).pin_memory() for _ in range(2)]

    def e2e_run(n_steps):
        """n_steps fwd+bwd steps from pinned host buffers; the upload of step i+1 overlaps the
        compute of step i (double buffering, as a training loop prefetches its next batch).
        Every step's loss is copied to pinned host memory and read by the host; the read of
        step i happens after step i+1 has been enqueued (a one-step-lagged logging read), so the
        host's launch work for the next step is not serialised behind the GPU's current one."""
        cur = torch.cuda.current_stream()
        nxt = upload(0)
        last, pending = None, None
        for i in range(n_steps):
            slot, ev = nxt
            cur.wait_event(ev)
            if i + 1 < n_steps:
                nxt = upload(i + 1)
            st_d = {k: v.detach().requires_grad_(True) for k, v in slot["st"].items()}
            loss = mod(slot["lg"].detach().requires_grad_(True), slot["tg"], st_d, slot["te"], slot["at"])
            loss.backward()
            loss_host[i % 2].copy_(loss.detach(), non_blocking=True)   # device->host read of the step's result
            done = torch.cuda.Event()
            done.record(cur)
            slot["free"] = done
            if pending is not None:
                pending[0].synchronize()
                last = float(pending[1])
            pending = (done, loss_host[i % 2])
        pending[0].synchronize()
        last = float(pending[1])
        return last

    # the first upload (1.1 GB over PCIe, ~22 ms) cannot overlap anything: run enough steps that the
    # pipeline fill is a small share of the wall-clock window, as it is in a training loop
    e2e_steps = max(24, args.steps)                      # enough steps that one host hiccup does not move the mean
    e2e_run(2)
    barrier()
    t0 = time.perf_counter()
    e2e_run(e2e_steps)
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps
    t = torch.tensor([e2e_ms], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = work.batch * world / (float(t.item()) * 1e-3)

    pk = peaks()
    line = {
        "metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(ms_per_step, 4), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": data_label(args),
        "config": {"workload": work.name, "batch_per_gpu": work.batch, "student_tokens": work.n_student,
                   "teacher_tokens": work.n_teacher, "student_dim": work.d_student,
                   "teacher_dim": work.d_teacher, "teacher_layers": work.teacher_layers,
                   "token_dtype": str(work.token_dtype).replace("torch.", ""),
                   "features": args.features, "cuda_graph": bool(args.graph),
                   "cache": "inputs (>=1 GB tokens + attention maps per step) exceed the 126 MB L2",
                   "parallelism": f"dp{world}"},
        "step_ms": [round(x, 2) for x in per_step],
        "clocks": clocks, "gpu_launches": (launches // max(1, args.steps)) or launches_eager,
        "e2e": {"value": round(e2e_value, 2), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": 4, "ms_per_step": round(float(t.item()), 3),
                "h2d_gb_per_s_per_gpu": round(h2d / (float(t.item()) * 1e-3) / 1e9, 2),
                "staging": numa,
                "inputs": ("student + teacher tokens in full; of every teacher attention map only the CLS query "
                           "row (B,H,N+1) is staged on the host and copied -- the loss reads nothing else "
                           "(SURVEY 8-f1 capture contract, not the reference's full-map input)"
                           if work.has_cls else "student + teacher tokens and attention maps in full")},
    }
    if parity is not None:
        line["dp_parity"] = parity
    # every rank runs the profiled step (the loss all-reduces inside); only rank 0 reports
    records, total, fp32_peak = roofline(work, eager_mod, args5, pk, clocks)
    if rank == 0:
        line["kernels"] = records[:14]
        line["kernel_ms_total"] = round(total, 3)
        # the dominant kernel that has a byte / flop model (records are sorted by time)
        top = next((r for r in records if "achieved" in r), None)
        traffic = measured_traffic() if args.workload == "c2" and args.batch == 256 else {}
        if top:
            line["roofline"] = {"kernel": top["kernel"], "bound": top["bound"], "achieved": round(top["achieved"], 2),
                                "peak": round(top["peak"], 2), "unit": top["unit"], "frac": round(top["frac"], 4),
                                "traffic": traffic.get(top["kernel"]), "peak_source": pk["source"]}
        # one entry per kernel family the north star asks evidence for: tensor pipe (Gram, 3xTF32 GEMMs), FP32
        # pipe (Jacobi, Cholesky), HBM (mix / resample, weight gradient); the fp32 peak uses the SM clock
        # observed during the run; `traffic` = DRAM bytes per launch from the committed ncu capture
        line["rooflines"] = [{"kernel": r["kernel"], "bound": r["bound"], "achieved": round(r["achieved"], 2),
                              "peak": round(r["peak"], 2), "unit": r["unit"], "frac": round(r["frac"], 4),
                              "ms": r["ms"], "launches": r.get("launches", 1),
                              "traffic": traffic.get(r["kernel"])} for r in records if "achieved" in r]
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(args, sample_batch=args.cpu_batch, steps=1)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------ data-parallel parity
def dp_parity(args, rank, world, device):
    """Before any timing at N > 1: one small-batch data-parallel step of the timed workload's shape,
    checked on rank 0 against the CPU oracle evaluated on the CONCATENATED batch (the
    parity target of the all-reduced statistics).  Returns the record printed as `dp_parity`."""
    from tests import _cases as cs            # test infrastructure: the oracle is the checker here
    local_batch = 4
    temps = [0.3, 0.541, 0.8, 1.2]
    work = cs.workload(args.workload, local_batch)
    temps = temps[:work.num_points] + [0.6] * max(0, work.num_points - len(temps))
    inputs = syn.make_inputs(work, seed=6, batch_offset=rank * local_batch)
    got = cs.run_cuda(work, inputs, temps, device=device)
    glob = float(got["module"].last["global_loss"])
    rec = None
    if rank == 0:
        full = cs.workload(args.workload, local_batch * world)
        ref = cs.run_oracle(full, syn.make_inputs(full, seed=6), temps)
        cos = min(cs.cosine(got["grad_students"][l] / world, ref["grad_students"][l][:local_batch])
                  for l in ref["layers"])
        rec = {"local_batch": local_batch, "world": world, "reference": "CPU oracle on the concatenated batch",
               "loss_rel": abs(glob - float(ref["loss"])) / abs(float(ref["loss"])),
               "weights_max_abs": float((got["weights"] - ref["weights"]).abs().max()),
               "ranks_equal": got["ranks"] == ref["ranks"],
               "grad_cos_min": cos,
               "log_temp_grad_cos": cs.cosine(got["grad_log_temps"], ref["grad_log_temps"])}
        rec["ok"] = bool(rec["loss_rel"] < 1e-3 and rec["weights_max_abs"] < 1e-4 and cos > 0.999
                         and rec["log_temp_grad_cos"] > 0.999)
    del got
    torch.cuda.empty_cache()
    dist.barrier()
    return rec


# ------------------------------------------------------------------ CPU arms
def cpu_setup(args, batch):
    os.sched_setaffinity(0, ORIG_AFFINITY)                # undo the NUMA pinning of the GPU arm: all host cores
    torch.set_num_threads(os.cpu_count())
    work = syn.scaled(syn.WORKLOADS[args.workload], batch)
    mod = build_module(work, "cpu", impl="reference")
    logits, targets, st, te, at = make_inputs(args, work, 0, "cpu")
    st = {k: v.float().requires_grad_(True) for k, v in st.items()}
    te = {k: v.float() for k, v in te.items()}
    logits.requires_grad_(True)

    def step():
        for v in st.values():
            v.grad = None
        loss = mod(logits, targets, st, te, at)
        loss.backward()

    return work, step


def cpu_baseline(args, sample_batch, steps):
    """The reference algorithm (oracle port: same torch.linalg calls as the reference) on the
    host cores, on a bounded sample of the workload."""
    work, step = cpu_setup(args, sample_batch)
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    return {"value": round(sample_batch / dt, 3), "unit": UNIT, "cores": torch.get_num_threads(),
            "kind": "port", "measured_batch": sample_batch,
            "sample": f"{steps} step(s) of {work.name}: batch {sample_batch} of the workload's {args.batch}, "
            "fp32, autocast off (bf16 tokens upcast exactly)", "seconds_per_step": round(dt, 3)}


def run_reference(args):
    """The reference's algorithm on the host cores (oracle port; the Python reference itself cannot travel
    to the GPU box).  `config.batch_per_gpu` is the batch actually run: a bounded sample of the workload,
    sized so that K + W steps end within a few minutes."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    work, step = cpu_setup(args, args.cpu_batch)
    budget_s = 240.0
    t0 = time.perf_counter()
    step()                                            # first warm-up step doubles as the time probe
    probe = time.perf_counter() - t0
    warm = max(1, min(args.warmup, int(0.2 * budget_s / probe)))
    steps = max(1, min(args.steps, int(0.8 * budget_s / probe)))
    for _ in range(warm - 1):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    value = args.cpu_batch / dt
    full = syn.scaled(syn.WORKLOADS[args.workload], args.batch)
    line = {
        "impl": "reference", "metric": METRIC, "value": round(value, 3), "unit": UNIT,
        "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": steps, "warmup": warm,
        "ms_per_step": round(dt * 1e3, 2), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": data_label(args),
        "config": {"workload": full.name, "batch_per_gpu": args.cpu_batch, "workload_batch": full.batch,
                   "student_tokens": full.n_student,
                   "teacher_tokens": full.n_teacher, "student_dim": full.d_student,
                   "teacher_dim": full.d_teacher, "teacher_layers": full.teacher_layers,
                   "token_dtype": "float32 (bf16 tokens upcast: the reference cannot take bf16)",
                   "features": args.features,
                   "parallelism": "cpu"},
        "measured_batch": args.cpu_batch,
        "cpu_baseline": {"value": round(value, 3), "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                         "measured_batch": args.cpu_batch,
                         "sample": f"{steps} step(s) (+{warm} warm-up) at batch {args.cpu_batch}: a bounded sample "
                                   f"of the workload's batch {args.batch}; requested steps/warm-up "
                                   f"{args.steps}/{args.warmup}, cut to a {int(budget_s)} s budget if needed"},
        "e2e": {"value": round(value, 3), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(syn.WORKLOADS))
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--cpu-batch", type=int, default=32)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-dp-parity", action="store_true")
    ap.add_argument("--graph", action="store_true", help="replay fwd+bwd from one CUDA graph (graphed.py)")
    ap.add_argument("--features", default="backbone", choices=["backbone", "spectral"])
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
