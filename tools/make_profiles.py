"""Turns the ncu outputs brought back in gpurun_out/ into the tracked summaries under profiles/.
    python tools/make_profiles.py raw <report.ncu-rep in gpurun_out> <name under profiles/>
    python tools/make_profiles.py launches <launch list .csv in gpurun_out> <prefix under profiles/>
    python tools/make_profiles.py r01        (the round-1 set)"""
import collections, csv, gzip, os, re, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out"); P = os.path.join(ROOT, "profiles")

def launch_list(src, dst_prefix):
    path = os.path.join(G, src)
    if not os.path.exists(path): return
    lines = [l for l in open(path) if not l.startswith("==")]
    with gzip.open(os.path.join(P, dst_prefix + "_launches_raw.csv.gz"), "wt") as f: f.writelines(lines)
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", "")); u = row["Metric Unit"]
        v = v / 1e3 if u in ("ns", "nsecond") else (v * 1e3 if u in ("ms", "msecond") else v)
        name = re.sub(r"\(.*", "", row["Kernel Name"])
        if "dgemm" in name:
            g = eval(row["Grid Size"].replace("(", "[").replace(")", "]"))
            name += " [trailing]" if g[0] * g[1] >= 600 else " [panel/trsm]"
        agg[name][0] += 1; agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    with open(os.path.join(P, dst_prefix + "_launches_by_kernel.csv"), "w") as f:
        f.write("kernel,launches,total_us,share_pct,avg_us\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"\"{k}\",{v[0]},{v[1]:.1f},{100*v[1]/tot:.2f},{v[1]/v[0]:.2f}\n")
        f.write(f"\"TOTAL (serialised, cold cache)\",{sum(v[0] for v in agg.values())},{tot:.1f},100,\n")

def raw_page(rep, dst):
    path = os.path.join(G, rep)
    if not os.path.exists(path): return
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    keep = re.compile(r"Kernel Name|gpu__time_duration|dram__bytes|dram__cycles_active|gpu__dram_throughput|sm__pipe_tensor|"
                      r"dmma|sm__warps_active|launch__|sm__throughput|lts__t_bytes|l1tex__data_bank_conflicts|"
                      r"smsp__inst_executed.sum|sm__cycles_(active|elapsed)|smsp__warp_issue_stalled.*_per_warp_active|sm__inst_executed_pipe_fp64")
    idx = [i for i, h in enumerate(hdr) if keep.search(h)]
    with open(os.path.join(P, dst), "w") as f:
        w = csv.writer(f); w.writerow(["metric", "unit"] + [f"launch{j}" for j in range(len(rows) - 2)])
        for i in idx: w.writerow([hdr[i], units[i]] + [r[i] for r in rows[2:]])

def round1(tag="r01"):
    launch_list("launches_r1c_cfg2a.csv", f"{tag}_cfg2a")
    raw_page("prof_gemm_r1_v1.ncu-rep", f"{tag}_ncu_full_dgemm_128x64_2cta.csv")
    raw_page("prof_gemm_r1.ncu-rep", f"{tag}_ncu_full_dgemm_128x128_1cta.csv")
    raw_page("prof_assemble_r1.ncu-rep", f"{tag}_ncu_full_assemble.csv")
    raw_page("prof_leaf_r3.ncu-rep", f"{tag}_ncu_full_panel_leaf.csv")
    raw_page("prof_trsm_r1.ncu-rep", f"{tag}_ncu_full_trsm_strip.csv")
    raw_page("prof_trisolve_r1.ncu-rep", f"{tag}_ncu_full_trisolve_step.csv")
    raw_page("prof_moments_r1.ncu-rep", f"{tag}_ncu_full_truncation_moments.csv")
    raw_page("prof_perm_r1.ncu-rep", f"{tag}_ncu_full_perm_gather.csv")
    for f in ("bench_1gpu.json", "leaf_timing.log"):
        if os.path.exists(os.path.join(G, f)): shutil.copy(os.path.join(G, f), os.path.join(P, f"{tag}_{f}"))


if __name__ == "__main__":
    if len(sys.argv) == 4 and sys.argv[1] == "raw":
        raw_page(sys.argv[2], sys.argv[3])
    elif len(sys.argv) == 4 and sys.argv[1] == "launches":
        launch_list(sys.argv[2], sys.argv[3])
    elif len(sys.argv) == 2 and sys.argv[1] == "r01":
        round1()
    else:
        sys.exit(__doc__)
    print(sorted(os.listdir(P)))
