"""Summarises one round's ncu output into profiles/: the launch list of a C2 B=256 step (per-kernel time,
share, launches) and the `--set full` raw pages (duration, DRAM bytes, pipe utilisation, issue slots, stall
reasons, shared-memory conflicts), plus profiles/<tag>_traffic.json = DRAM bytes per launch of the kernels
bench.py attaches a byte model to.    python tools/ncu_summary.py <tag> [gpurun_out]"""
import csv
import json
import os
import re
import sys
from collections import OrderedDict

tag = sys.argv[1]
src = sys.argv[2] if len(sys.argv) > 2 else "gpurun_out"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
dst = os.path.join(ROOT, "profiles")


def short(name):
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*$", "", name)
    return name.replace("basd::", "")


def launch_list(path):
    rows = [r for r in csv.reader(l for l in open(path) if not l.startswith("=="))]
    hdr = rows[0]
    k_i, m_i, v_i = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
    u_i = hdr.index("Metric Unit")
    agg = OrderedDict()
    for r in rows[1:]:
        if len(r) <= v_i or r[m_i] != "gpu__time_duration.sum":
            continue
        ms = float(r[v_i].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[u_i], 1e-6)
        a = agg.setdefault(short(r[k_i]), [0.0, 0])
        a[0] += ms
        a[1] += 1
    return agg


KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__cluster_dim_x",
        "launch__registers_per_thread", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def full_pages(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        rec = {"kernel": short(d["Kernel Name"])}
        for k in KEYS:
            if k in d and d[k] != "":
                rec[k] = (d[k], u.get(k, ""))
        stalls = {}
        for k in hdr:
            m = re.match(r"smsp__average_warps_issue_stalled_(.*)_per_issue_active.ratio", k)
            if m and d[k]:
                try:
                    v = float(d[k].replace(",", ""))
                except ValueError:
                    continue
                if v >= 0.15:
                    stalls[m.group(1)] = round(v, 2)
        rec["stalls"] = stalls
        out.append(rec)
    return out


def fnum(t):
    return float(t[0].replace(",", ""))


lines = [f"# Round {tag} ncu summary -- C2 (DeiT-S <- DeiT-B, B=256, N=196, bf16 tokens), one fwd+bwd step", ""]
ll = os.path.join(src, f"{tag}_launches_c2_b256.csv")
if os.path.exists(ll):
    agg = launch_list(ll)
    total = sum(v[0] for v in agg.values())
    n = sum(v[1] for v in agg.values())
    lines += [f"## Launch list: {n} kernel launches, {total:.2f} ms of kernel time per step "
              "(`ncu --metrics gpu__time_duration.sum --clock-control none`: cold-cache, serialised launches -- "
              "compare SHARES with bench.py's `kernels` table, not absolutes)", "",
              "| kernel | ms | share | launches |", "|---|---|---|---|"]
    for k, (ms, c) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        if ms / total >= 0.001:
            lines.append(f"| `{k}` | {ms:.3f} | {100 * ms / total:.1f}% | {c} |")
    lines.append("")
    open(os.path.join(dst, f"{tag}_launches_c2_b256.csv"), "w").write(open(ll).read())
traffic = {}
recs_all = []
# later parts are later captures: a kernel captured again replaces its earlier records; kernels that no longer
# exist in the library are dropped
RETIRED = ("tc3::gemm_tc3_kernel", "colsum_partial_bf16x8_kernel", "colsum_reduce_kernel")
for part in ("", "_jacobi_proc", "_jacobi_kxk", "_chol_reg", "_stepk", "_final"):
    p = os.path.join(src, f"{tag}_full{part}_raw.csv")
    if os.path.exists(p):
        new = [r for r in full_pages(p) if not r["kernel"].startswith(RETIRED)]
        names = {r["kernel"] for r in new}
        recs_all = [r for r in recs_all if r["kernel"] not in names] + new
for part in ("step",):
    if not recs_all:
        continue
    recs = recs_all
    # the largest launch of every kernel
    best = OrderedDict()
    for r in recs:
        t = fnum(r["gpu__time_duration.sum"])
        if r["kernel"] not in best or t > fnum(best[r["kernel"]]["gpu__time_duration.sum"]):
            best[r["kernel"]] = r
    lines += ["## `--set full` capture of the whole step: the longest launch of each kernel (B = 256), kernels "
              "above 0.03 ms (statistics + selector from the whole-step capture, which ran into its time limit after 129 "
              "launches; the rest from targeted captures: `tools/gpu_profile2.sh`; kernels changed after those "
              "captures -- GEMM, token Gram, fp64 rotation, cluster Jacobi -- re-captured by `tools/gpu_profile3.sh`)", ""]
    for k, r in sorted(best.items(), key=lambda kv: -fnum(kv[1]["gpu__time_duration.sum"])):
        if fnum(r["gpu__time_duration.sum"]) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0}.get(r["gpu__time_duration.sum"][1], 1.0) < 0.03:
            continue
        lines.append(f"### `{k}`")
        lines.append("")
        for key in KEYS:
            if key in r:
                lines.append(f"- {key}: {r[key][0]} {r[key][1]}")
        if r["stalls"]:
            lines.append("- warps stalled per issue-active cycle (>= 0.15): " +
                         ", ".join(f"{a} {b}" for a, b in sorted(r["stalls"].items(), key=lambda kv: -kv[1])))
        lines.append("")
        if "dram__bytes_read.sum" in r and "dram__bytes_write.sum" in r:
            rd = fnum(r["dram__bytes_read.sum"]) * UNIT.get(r["dram__bytes_read.sum"][1], 1)
            wr = fnum(r["dram__bytes_write.sum"]) * UNIT.get(r["dram__bytes_write.sum"][1], 1)
            traffic[k] = {"dram_bytes": rd + wr, "ms": fnum(r["gpu__time_duration.sum"]) *
                          {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "msecond": 1.0, "usecond": 1e-3, "nsecond": 1e-6}.get(
                              r["gpu__time_duration.sum"][1], 1.0)}
open(os.path.join(dst, f"{tag}_ncu_summary.md"), "w").write("\n".join(lines) + "\n")
# entry point -> kernel that carries its byte model
entry = {"basd_mix_interp": "mix_interp_kernel", "basd_weight_grad": "weight_grad_onepass_kernel",
         "basd_token_gram_tc": "tc::token_gram_tc_kernel", "basd_jacobi_rows[procrustes]": "jacobi_rows_oe8_kernel<16, 13>",
         "basd_jacobi_rows[eig]": "oe8::jacobi_rows_oe8_cluster_kernel", "basd_jacobi_rows[kxk]": "jacobi_rows_oe8_split_kernel",
         "basd_pivoted_cholesky": "pivoted_cholesky_reg_kernel"}   # (the GEMM entries aggregate 25 + 5 launches of different shapes: no per-launch figure)
tj = {}
for e, k in entry.items():
    for name, rec in traffic.items():
        if name.startswith(k):
            tj[e] = int(rec["dram_bytes"])
json.dump({"per_launch_dram_bytes": tj, "kernels": traffic}, open(os.path.join(dst, f"{tag}_traffic.json"), "w"), indent=1)
print("\n".join(lines[:60]))
