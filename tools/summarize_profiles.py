#!/usr/bin/env python3
"""Turns the raw outputs of tools/gpu_profile.sh (gpurun_out/) into the summaries kept under profiles/:
launch shares of the bench command, the key metrics of the full ncu capture, roofline traffic for bench.py, bench lines
and parity reports. Run here (no GPU needed): python tools/summarize_profiles.py [round-tag]"""
import collections
import csv
import json
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT, PROF = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"


def last_json(path):
    lines = [l for l in open(path) if l.startswith("{")]
    return json.loads(lines[-1]) if lines else None


# ---- launch list -------------------------------------------------------------------------------------------------
src = os.path.join(OUT, "launches_bench_default.csv")
if os.path.isfile(src):
    shutil.copy(src, os.path.join(PROF, f"{tag}_launches_bench_default_cmd.csv"))
    lines = [l for l in open(src) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", row["Kernel Name"])
        v, unit = float(row["Metric Value"].replace(",", "")), row["Metric Unit"]
        us = v / 1000 if unit.startswith("n") else v * 1000 if unit.startswith("m") else v
        a = agg.setdefault(name, [0, 0.0, 0])
        a[0] += 1
        a[1] += us
        a[2] += 1 if us > 100 else 0
    tot = sum(a[1] for a in agg.values())
    it = {k: a for k, a in agg.items() if not re.search(r"generate_iid|k_stats|read_probe", k)}
    tot_it = sum(a[1] for a in it.values())
    with open(os.path.join(PROF, f"{tag}_launch_shares.txt"), "w") as f:
        f.write(f"# aggregated from profiles/{tag}_launches_bench_default_cmd.csv:\n"
                "#   ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-ab --no-parity\n"
                "# (the bench command minus the CPU-baseline and A/B legs; per-launch times are cold-cache and serialised: compare shares).\n"
                "# k_generate_iid / k_stats / k_read_probe run before the timed region. n counts every launch, including the look-ahead\n"
                "# launches of a finished solve that return at their first instruction (n_busy = launches longer than 100 us).\n"
                "# Shares of the iteration kernels alone: " +
                ", ".join(f"{k.replace('void ', '')} {100 * a[1] / tot_it:.1f} %" for k, a in sorted(it.items(), key=lambda kv: -kv[1][1])[:4]) + "\n")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{k[:58]:58s} n={a[0]:4d} n_busy={a[2]:4d} total={a[1] / 1000:10.3f} ms share={100 * a[1] / tot:5.1f}% avg={a[1] / a[0]:10.1f} us\n")
    print("launch shares written")

# ---- full capture ------------------------------------------------------------------------------------------------
rep = os.path.join(OUT, "prof_multi.ncu-rep")
if os.path.isfile(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    keep = [h for h in hdr if re.match(r"(gpu__time_duration.sum|dram__bytes_(read|write).sum($|\.per_second|\.pct)|gpu__dram_throughput|launch__(registers_per_thread$|grid_size|block_size|"
                                       r"occupancy_limit|shared_mem_per_block_dynamic|waves)|sm__warps_active.avg.pct|sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active|"
                                       r"l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum|l1tex__data_pipe_lsu_wavefronts_mem_shared.sum$|l1tex__throughput.avg.pct|lts__throughput.avg.pct|"
                                       r"lts__t_sector_hit_rate.pct|smsp__average_warps_issue_stalled_.*_per_issue_active|sm__throughput.avg.pct|sm__cycles_active.avg$|sm__cycles_elapsed.max$)", h)]
    traffic = {}
    with open(os.path.join(PROF, f"{tag}_ncu_full_matrix_kernels.txt"), "w") as f:
        f.write("# ncu --set full --clock-control none --import-source on -k regex:'k_gram|k_ax_multi' -s 6 -c 4,\n"
                "# `python bench.py --N 20000 --Mt 106250 --steps 1 --warmup 1 --no-cpu-baseline --no-ab --no-parity`\n"
                "# (one 8-GPU shard of the headline configuration: 17.000 GB of A per pass). Default (onepass) schedule: every pass of an\n"
                "# iteration is one of these two kernels (k_gram_wsx: A^T q and A A^T q of both systems; k_ax_multi: the first A p of the solves). Per-launch values.\n")
        for r in rows[2:]:
            name = r[idx["Kernel Name"]]
            f.write("\nKernel Name".ljust(77) + name[:160] + "\n")
            for h in keep:
                f.write(f"{h:76s}{r[idx[h]]} {units[idx[h]]}\n")
            short = "k_ax_multi" if "k_ax_multi" in name else "k_gram" if "k_gram" in name else "k_atx_smem"
            def gb(h):
                v, u = float(r[idx[h]].replace(",", "")), units[idx[h]]
                return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}[u]
            traffic.setdefault(short, []).append(gb("dram__bytes_read.sum") + gb("dram__bytes_write.sum"))
    tj_path = os.path.join(PROF, "ncu_traffic.json")
    tj = json.load(open(tj_path))
    for k, v in traffic.items():
        rec = {"N": 20000, "M_local": 106250, "algorithmic_bytes": 17000000000, "dram_bytes": sum(v) / len(v),
               "source": f"profiles/{tag}_ncu_full_matrix_kernels.txt ({k}, mean of {len(v)} launches)"}
        old = tj.get(k)
        others = [r for r in old if r.get("M_local") != 106250] if isinstance(old, list) else []    # captures of other launch shapes stay
        tj[k] = others + [rec] if others else rec
    json.dump(tj, open(tj_path, "w"), indent=1)
    print("ncu summary written", {k: sum(v) / len(v) for k, v in traffic.items()})

# ---- the fused pass at the 1-GPU bench's own launch shape (136 GB), one launch --------------------------------------
rep = os.path.join(OUT, "prof_gram136.ncu-rep")
if os.path.isfile(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    r = rows[2]

    def val(h):
        v, u = float(r[idx[h]].replace(",", "")), units[idx[h]]
        return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}.get(u, 1.0)
    dst = f"{tag}_ncu_full_k_gram_wsx_136GB.txt"
    with open(os.path.join(PROF, dst), "w") as f:
        f.write("# ncu --set full --clock-control none --import-source on -k regex:k_gram -s 8 -c 1,\n"
                "# `python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-ab --no-parity` (N = 20000, Mt = 850000: the 1-GPU bench's launch shape, 136.000 GB per pass)\n\n")
        f.write("Kernel Name".ljust(76) + r[idx["Kernel Name"]][:160] + "\n")
        for h in hdr:
            if re.match(r"(gpu__time_duration.sum|dram__bytes_(read|write).sum($|\.per_second)|launch__(registers_per_thread$|grid_size|block_size)|"
                        r"sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active|lts__t_sector_hit_rate.pct|sm__throughput.avg.pct)", h):
                f.write(f"{h:76s}{r[idx[h]]} {units[idx[h]]}\n")
    tj_path = os.path.join(PROF, "ncu_traffic.json")
    tj = json.load(open(tj_path))
    big = {"N": 20000, "M_local": 850000, "algorithmic_bytes": 136000000000, "dram_bytes": val("dram__bytes_read.sum") + val("dram__bytes_write.sum"),
           "source": f"profiles/{dst} (ncu --set full capture of one k_gram_wsx launch of the 1-GPU bench command)"}
    old = tj.get("k_gram")
    small = [x for x in (old if isinstance(old, list) else [old]) if x and x.get("M_local") != 850000]
    tj["k_gram"] = [big] + small
    json.dump(tj, open(tj_path, "w"), indent=1)
    print("136 GB capture written", big["dram_bytes"])

# ---- multi-GPU bench lines (tools/gpu_scale.sh) ------------------------------------------------------------------------
for f_ in sorted(os.listdir(OUT)) if os.path.isdir(OUT) else []:
    m = re.fullmatch(r"bench_g(\d+)_of(\d+)\.log", f_)
    if m and int(m.group(1)) > 1:
        d = last_json(os.path.join(OUT, f_))
        if d:
            json.dump(d, open(os.path.join(PROF, f"{tag}_bench_{m.group(1)}gpu.json"), "w"), indent=1)

# ---- bench lines, parity reports ---------------------------------------------------------------------------------
for src, dst in (("bench.log", f"{tag}_bench_1gpu.json"), ("bench_ref.log", f"{tag}_bench_reference_arm.json")):
    p = os.path.join(OUT, src)
    if os.path.isfile(p):
        d = last_json(p)
        if d:
            json.dump(d, open(os.path.join(PROF, dst), "w"), indent=1)
rep = {}
for s in ("onepass", "recycled", "fused", "plain"):
    p = os.path.join(OUT, f"parity_report_{s}.json")
    if os.path.isfile(p):
        try:
            rep[s] = json.load(open(p))
        except Exception:
            pass
if rep:
    json.dump(rep, open(os.path.join(PROF, f"{tag}_parity_report_vs_reference_fixtures.json"), "w"), indent=1)
p = os.path.join(OUT, "configs.json")
if os.path.isfile(p):
    shutil.copy(p, os.path.join(PROF, f"{tag}_configs_C4_C5.json"))
p = os.path.join(OUT, "parity_c2.json")
if os.path.isfile(p):
    shutil.copy(p, os.path.join(PROF, f"{tag}_parity_config2_vs_reference_binary.json"))
for f_ in ("pytest_gpu.log", "host.txt"):
    p = os.path.join(OUT, f_)
    if os.path.isfile(p):
        shutil.copy(p, os.path.join(PROF, f"{tag}_{f_}"))
print("done")
