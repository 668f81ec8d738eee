"""Summarise .ncu-rep captures and the extension's SASS into tracked files under profiles/.

    python tools/summarize_ncu.py gpurun_out/prof_gemm_v2.ncu-rep [...]   # -> profiles/<name>.md
    python tools/summarize_ncu.py --sass                                   # -> profiles/sass_summary.md
"""
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__cluster_size",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.avg", "sm__cycles_active.avg",
]


def summarize(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    if len(rows) < 3:
        return f"(could not read {rep})\n"
    hdr, units = rows[0], rows[1]
    name = os.path.splitext(os.path.basename(rep))[0]
    md = [f"# ncu summary: {name}", "", f"source: `{rep}` (captured with `ncu --set full --clock-control none --import-source on`)", ""]
    seen = set()
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        key = (d.get("Kernel Name", "?"), d.get("launch__grid_size", ""))
        if key in seen:        # a step capture launches the same kernel many times: keep the first of each (name, grid)
            continue
        seen.add(key)
        md.append(f"## {d.get('Kernel Name', '?')[:150]}")
        md.append("")
        md.append("| metric | value | unit |")
        md.append("|---|---|---|")
        for k in KEYS:
            for h, v, u in zip(hdr, r, units):
                if h == k or h.endswith("." + k):
                    md.append(f"| {k} | {v} | {u} |")
                    break
        stalls = sorted(((float(v), h) for h, v in zip(hdr, r)
                         if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")
                         and v not in ("", "n/a")), reverse=True)[:5]
        if stalls:
            md.append("")
            md.append("top warp-stall reasons (per issue): " + ", ".join(
                f"{h.split('stalled_')[1].split('_per_issue')[0]}={v:.2f}" for v, h in stalls))
        md.append("")
    return "\n".join(md) + "\n"


MNEMONICS = ["UTCHMMA", "UTCBAR", "UTCATOMSWS", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "HMMA", "MUFU.EX2",
             "LDGSTS", "MEMBAR", "RED", "ATOM"]


def sass_summary():
    so = os.path.join(ROOT, "distributed_training_guide_b200", "_C.so")
    out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
    md = ["# SASS evidence per kernel (`cuobjdump -sass distributed_training_guide_b200/_C.so`, sm_100a)", "",
          "Counts of the mnemonics that identify Blackwell-native paths: `UTCHMMA` = tcgen05.mma, `LDTM`/`STTM` = tcgen05.ld/st,",
          "`UTMALDG` = TMA load, `UTCBAR` = tcgen05.commit, `UTCATOMSWS` = TMEM alloc, `SYNCS` = mbarrier ops; `HMMA` would be the",
          "legacy mma.sync path (none expected).", "", "| kernel | " + " | ".join(MNEMONICS) + " |", "|---|" + "---|" * len(MNEMONICS)]
    cur, counts = None, {}
    order = []
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = {k: 0 for k in MNEMONICS}
            order.append(cur)
            continue
        if cur:
            for k in MNEMONICS:
                if re.search(r"\b" + re.escape(k), line):
                    counts[cur][k] += 1
    demangled = subprocess.run(["c++filt"], input="\n".join(order), capture_output=True, text=True).stdout.splitlines()
    for mangled, nice in zip(order, demangled):
        nice = re.sub(r"\(.*", "", nice).replace("dtg::", "")[:90]
        c = counts[mangled]
        md.append(f"| `{nice}` | " + " | ".join(str(c[k]) for k in MNEMONICS) + " |")
    return "\n".join(md) + "\n"


if __name__ == "__main__":
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    if "--sass" in sys.argv:
        open(os.path.join(ROOT, "profiles", "sass_summary.md"), "w").write(sass_summary())
        print("wrote profiles/sass_summary.md")
    for rep in [a for a in sys.argv[1:] if a.endswith(".ncu-rep")]:
        name = os.path.splitext(os.path.basename(rep))[0]
        open(os.path.join(ROOT, "profiles", f"{name}.md"), "w").write(summarize(rep))
        print(f"wrote profiles/{name}.md")
