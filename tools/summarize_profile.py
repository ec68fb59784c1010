#!/usr/bin/env python3
"""tools/summarize_profile.py <tag> — turn the raw ncu outputs in gpurun_out/ (launches.csv from the
gpu__time_duration pass, prof_<tag>.ncu-rep from the --set full pass) into the tracked summaries under profiles/."""
import csv, collections, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out"); P = os.path.join(ROOT, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
os.makedirs(P, exist_ok=True)
out = []

# ---- launch list: share of device time per kernel ----
lp = os.path.join(G, "launches.csv")
if os.path.exists(lp):
    rows = [r for r in csv.reader(open(lp)) if len(r) > 5]
    hdr = rows[0]; ik = hdr.index("Kernel Name"); iv = hdr.index("Metric Value"); iu = hdr.index("Metric Unit")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        v = float(r[iv].replace(",", "")); u = r[iu]
        us = v / 1000.0 if u in ("ns", "nsecond") else v * (1000.0 if u in ("ms", "msecond") else 1.0)
        name = r[ik].split("(")[0].replace("void ", "").replace("rt::", "")
        agg[name][0] += 1; agg[name][1] += us
    tot = sum(a[1] for a in agg.values())
    out.append("## Launch list (`ncu --metrics gpu__time_duration.sum --clock-control none`, `python tools/prof_cmd.py $SPPN` (tools/gpu_r02.sh ncu; 40 spp since r02v, 12 before):\n"
               "C4 800x800, Philox mode; per-launch times are serialised and cold-cache: compare SHARES)\n")
    out.append("| kernel | launches | total us | share | mean us |\n|---|---|---|---|---|")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append("| %s | %d | %.1f | %.1f%% | %.2f |" % (k, a[0], a[1], 100 * a[1] / tot, a[1] / a[0]))
    out.append("")
    with open(os.path.join(P, "%s_launches.csv" % tag), "w") as f:
        w = csv.writer(f); w.writerow(["kernel", "launches", "total_us", "share"])
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            w.writerow([k, a[0], "%.2f" % a[1], "%.4f" % (a[1] / tot)])

# ---- full capture: key metrics per kernel ----
rep = os.path.join(G, "prof_%s.ncu-rep" % tag)
traffic = {}
if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    open(os.path.join(G, "prof_%s_raw.csv" % tag), "w").write(raw)
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    keys = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "smsp__issue_active.avg.pct", "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__inst_executed.sum",
            "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
            "lts__t_bytes.sum", "l1tex__t_bytes.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
            "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum", "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum",
            "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum"]
    out.append("## `ncu --set full --clock-control none` (same command), first captured launch of each kernel\n")
    seen = {}
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "").replace("rt::", "")
        if name in seen:
            continue
        seen[name] = r
    names = list(seen)
    out.append("| metric | unit | " + " | ".join(names) + " |\n|---|---|" + "---|" * len(names))
    for k in keys:
        if k in hdr:
            i = hdr.index(k)
            out.append("| %s | %s | %s |" % (k, units[i], " | ".join(seen[n][i] for n in names)))
    out.append("")
    def num(x):
        return float(x.replace(",", ""))
    for n in names:
        r = seen[n]
        def gb(k):
            i = hdr.index(k); v = num(r[i]); u = units[i]
            return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        traffic[n] = {"dram_bytes": gb("dram__bytes_read.sum") + gb("dram__bytes_write.sum"),
                      "grid": int(num(r[hdr.index("launch__grid_size")])), "block": int(num(r[hdr.index("launch__block_size")])),
                      "duration_us": num(r[hdr.index("gpu__time_duration.sum")]) * {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "s": 1e6, "second": 1e6}.get(units[hdr.index("gpu__time_duration.sum")], 1.0),
                      "inst_issued_pct_of_peak": num(r[hdr.index("sm__inst_issued.avg.pct_of_peak_sustained_active")]) if "sm__inst_issued.avg.pct_of_peak_sustained_active" in hdr else None,
                      "threads_per_instruction": num(r[hdr.index("smsp__thread_inst_executed_per_inst_executed.ratio")]),
                      "l1_hit_pct": num(r[hdr.index("l1tex__t_sector_hit_rate.pct")]), "l2_hit_pct": num(r[hdr.index("lts__t_sector_hit_rate.pct")]),
                      "fma_pipe_pct": num(r[hdr.index("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active")]),
                      "warp_inst": num(r[hdr.index("smsp__inst_executed.sum")]),
                      "l2_bytes": 32.0 * num(r[hdr.index("lts__t_sectors.sum")]) if "lts__t_sectors.sum" in hdr else None,
                      "l1_bytes": 32.0 * num(r[hdr.index("SM_B.TriageCompute.l1tex__t_sectors.sum")]) if "SM_B.TriageCompute.l1tex__t_sectors.sum" in hdr else None,
                      # FP32 flop of the launch: (2 x FFMA + FADD + FMUL thread instructions per cycle) x elapsed cycles
                      "flop": (sum((2.0 if "ffma" in k else 1.0) * num(r[hdr.index(k)]) for k in
                                   ("smsp__sass_thread_inst_executed_op_ffma_pred_on.sum.per_cycle_elapsed",
                                    "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum.per_cycle_elapsed",
                                    "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum.per_cycle_elapsed") if k in hdr) *
                               num(r[hdr.index("smsp__cycles_elapsed.avg")])) if "smsp__cycles_elapsed.avg" in hdr else None}
        # rays of the launch: a k_trace warp owns RT_RANGE (env, default 64) rays of the dense layout, a k_shade thread one
        # (a k_shade block walks RT_SHADE_ITEMS chunks of its block size: env, default 8)
        per_thread = (int(os.environ.get("RT_RANGE", "64")) / 32.0) if n.startswith("k_trace") else float(os.environ.get("RT_SHADE_ITEMS", "8"))
        traffic[n]["rays"] = traffic[n]["grid"] * traffic[n]["block"] * per_thread
    # ---- per-source-line attribution ----
    sass = os.path.join(G, "elf", "all_%s.sass" % tag)  # nvdisasm -gi -c of the library's cubin (cuobjdump -xelf all)
    for kern, pref in (("k_trace", "_ZN2rt7k_trace"), ("k_shade", "_ZN2rt7k_shadeILi0E")):
        srccsv = os.path.join(G, "src_%s_%s.csv" % (kern, tag))
        with open(srccsv, "w") as f:
            subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern, "--launch-count", "1",
                            "--print-source", "sass"], stdout=f, text=True)
        # the sass_thread_inst_executed_op_f* counters do not see the packed f32x2 instructions (FADD2 / FMUL2 / FFMA2,
        # two IEEE operations per lane): count them from the SASS page and add them to the launch's flop
        rows = list(csv.reader(open(srccsv)))
        if len(rows) > 2 and "Source" in rows[1] and "Thread Instructions Executed" in rows[1]:
            isrc, ithr = rows[1].index("Source"), rows[1].index("Thread Instructions Executed")
            packed = 0.0
            first_addr = rows[2][0] if len(rows) > 2 else None
            for k, r in enumerate(rows[2:]):
                if k > 0 and r and r[0] == first_addr:
                    break  # the page lists every captured launch of the kernel one after the other: the first one only
                if len(r) <= max(isrc, ithr):
                    continue
                op = r[isrc].split()[1] if r[isrc].lstrip().startswith("@") and len(r[isrc].split()) > 1 else (r[isrc].split() or [""])[0]
                if op.startswith(("FADD2", "FMUL2")):
                    packed += 2.0 * float(r[ithr] or 0)
                elif op.startswith("FFMA2"):
                    packed += 4.0 * float(r[ithr] or 0)
            key = kern if kern in traffic else (kern + "<0>")
            # k_trace: the rays of the launch are the threads that executed the store of a traced ray's hit (the first
            # STG.E.64 of the kernel, rt_kernels.cuh `hit[gid] = ...`; the second one marks holes of the dense layout).
            # grid x block x 2 counts holes and the warps beyond order_len as rays (a late wave of a short job: 3x too many).
            if kern == "k_trace" and key in traffic:
                first = True
                for k, r in enumerate(rows[2:]):
                    if k > 0 and r and r[0] == first_addr:
                        break
                    if len(r) > max(isrc, ithr) and "STG.E.64" in r[isrc]:
                        if first:
                            traffic[key]["rays_upper_bound_grid"] = traffic[key]["rays"]
                            traffic[key]["rays"] = float(r[ithr] or 0)
                            first = False
                        else:
                            traffic[key]["holes"] = float(r[ithr] or 0)
                            break
            if key in traffic and traffic[key].get("flop") is not None:
                traffic[key]["flop_scalar_counters"] = traffic[key]["flop"]
                traffic[key]["flop_packed_f32x2"] = packed
                traffic[key]["flop"] += packed
    json.dump(traffic, open(os.path.join(P, "%s_traffic.json" % tag), "w"), indent=1)
    if os.path.exists(sass):
        for kern, pref in (("k_trace", "_ZN2rt7k_trace"), ("k_shade", "_ZN2rt7k_shadeILi0E")):
            srccsv = os.path.join(G, "src_%s_%s.csv" % (kern, tag))
            t = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_lines.py"), srccsv, sass, pref, "30"], stdout=subprocess.PIPE, text=True).stdout
            out.append("## %s: stall samples and instructions by source line (deepest inline frame outside rt_math.h / rng.h)\n\n```\n%s```\n" % (kern, t))
open(os.path.join(P, "%s_ncu_summary.md" % tag), "w").write("# ncu summary %s\n\n" % tag + "\n".join(out) + "\n")
print("wrote profiles/%s_ncu_summary.md" % tag)
