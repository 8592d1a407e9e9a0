#!/usr/bin/env python
"""Turns the ncu captures under gpurun_out/ into the small text summaries committed under profiles/.

    python tools/summarize_profiles.py r01 4500     # reads gpurun_out/r01_*.{csv,ncu-rep,json}; 4500 = --pairs of the
                                                    # `ncu --set full` captures (one launch = one chunk of that many pairs)
"""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GO = os.path.join(ROOT, "gpurun_out")
PR = os.path.join(ROOT, "profiles")

KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed"]


def launches(tag):
    path = os.path.join(GO, f"{tag}_launches.csv")
    rows = list(csv.reader(open(path)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    agg = collections.OrderedDict()
    for r in rows[hdr + 1:]:
        if len(r) < 10:
            continue
        name = r[4].split("(")[0]
        try:
            v = float(r[-1])
        except ValueError:
            continue
        v *= {"us": 1e-3, "ns": 1e-6, "s": 1e3}.get(r[-2], 1.0)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    ours = {k: v for k, v in agg.items() if "rp::" in k or "tc::" in k or "<unnamed>" in k}
    tot = sum(v[1] for k, v in ours.items() if "pipe_kernel" not in k)
    out = [f"# ncu --metrics gpu__time_duration.sum --clock-control none, `python bench.py --steps 1 --warmup 1 --no-cpu-baseline`",
           "# per-launch times are cold-cache and serialised: compare SHARES (pipeline kernels only; the two",
           "# *_pipe_kernel micro-benchmarks of rp_measure_pipes are listed but excluded from the share)",
           f"{'kernel':58s} {'launches':>8s} {'total ms':>10s} {'share':>7s}"]
    for k, v in sorted(ours.items(), key=lambda kv: -kv[1][1]):
        share = "" if "pipe_kernel" in k else f"{v[1] / tot:7.3f}"
        out.append(f"{k[:58]:58s} {v[0]:8d} {v[1]:10.3f} {share}")
    out.append(f"{'total (pipeline kernels)':58s} {'':8s} {tot:10.3f}")
    return "\n".join(out)


def ncu_raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, u, v = rows[0], rows[1], rows[2]
    d = {k: (v[i], u[i]) for i, k in enumerate(h)}
    lines = [f"kernel: {d.get('Kernel Name', ('?',))[0]}"]
    for k in KEYS:
        if k in d:
            lines.append(f"  {k:72s} {d[k][0]:>18s} {d[k][1]}")
    stalls = sorted(((float(val[0]), k) for k, val in d.items()
                     if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("per_issue_active.ratio")
                     and val[0] not in ("", "n/a")), reverse=True)[:6]
    lines.append("  top stall reasons (warps stalled per issue): " +
                 ", ".join(f"{k.split('stalled_')[1].split('_per_')[0]}={x:.2f}" for x, k in stalls))
    return "\n".join(lines), d


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
    os.makedirs(PR, exist_ok=True)
    parts = []
    if os.path.exists(os.path.join(GO, f"{tag}_launches.csv")):
        txt = launches(tag)
        open(os.path.join(PR, f"{tag}_launches_summary.txt"), "w").write(txt + "\n")
        subprocess.run(["cp", os.path.join(GO, f"{tag}_launches.csv"), os.path.join(PR, f"{tag}_launches.csv")])
        parts.append(txt)
    traffic = {}
    for name in ("tc", "bound", "lm", "score_survivors", "solve"):
        rep = os.path.join(GO, f"{tag}_{name}.ncu-rep")
        if os.path.exists(rep):
            txt, d = ncu_raw(rep)
            parts.append(f"## ncu --set full --clock-control none: {name}\n{txt}")
            try:
                def mb(k):
                    val, unit = d[k]
                    return float(val) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[unit]
                traffic[name] = mb("dram__bytes_read.sum") + mb("dram__bytes_write.sum")
            except Exception:
                pass
    open(os.path.join(PR, f"{tag}_ncu_summary.txt"), "w").write("\n\n".join(parts) + "\n")
    # SASS evidence: the Blackwell-only instructions of the shipped library (B200_PROFILING.md "What proves a Blackwell-native kernel")
    so = os.path.join(ROOT, "mdrp_b200", "librepose_b200.so")
    if os.path.exists(so):
        sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
        import re
        ops = collections.Counter(m.group(1).split(".")[0] for m in re.finditer(r"^\s+/\*[0-9a-f]+\*/\s+([A-Z0-9_.]+)", sass, re.M))
        want = ["UTCHMMA", "UTCBAR", "UTCATOMSWS", "LDTM", "STTM", "UTMALDG", "SYNCS", "FFMA2", "FADD2", "FMUL2", "FFMA", "DFMA", "DMUL", "DADD", "HMMA"]
        lines = ["# cuobjdump -sass mdrp_b200/librepose_b200.so: instruction counts (static) of the opcodes that matter",
                 "# UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG = cp.async.bulk.tensor (TMA), UTCBAR = tcgen05.commit,",
                 "# SYNCS = mbarrier ops, FFMA2 / FADD2 / FMUL2 = packed f32x2 (sm_100 only); HMMA (legacy mma.sync) must be absent"]
        lines += [f"{k:12s} {ops.get(k, 0)}" for k in want]
        ex = [l for l in sass.splitlines() if re.search(r"UTCHMMA|LDTM|STTM|UTMALDG|FFMA\.SAT|FADD2", l)][:24]
        lines += ["", "# excerpt (tc_count_kernel):"] + [l.rstrip()[:150] for l in ex]
        open(os.path.join(PR, f"{tag}_sass_excerpt.txt"), "w").write("\n".join(lines) + "\n")
    for f in (f"{tag}_all_configs.txt", f"{tag}_head_sweep.txt"):
        if os.path.exists(os.path.join(GO, f)):
            subprocess.run(["cp", os.path.join(GO, f), os.path.join(PR, f)])
    for f in (f"{tag}_bench.json", f"{tag}_bench_reference.json"):
        if os.path.exists(os.path.join(GO, f)):
            subprocess.run(["cp", os.path.join(GO, f), os.path.join(PR, f)])
    traffic["pairs_in_launch"] = int(sys.argv[2]) if len(sys.argv) > 2 else 4500  # --pairs of the ncu captures
    json.dump(traffic, open(os.path.join(PR, f"{tag}_dram_traffic_bytes.json"), "w"), indent=1)
    print("\n\n".join(parts))
    print(traffic)


if __name__ == "__main__":
    main()
