"""Summaries of ncu output for profiles/:
   python tools/summarize_ncu.py launches gpurun_out/launches.csv > profiles/rNN_launches.md
   python tools/summarize_ncu.py report  gpurun_out/prof.ncu-rep  > profiles/rNN_kernel.md
"""
import collections
import csv
import re
import subprocess
import sys


def launches(path):
    rows = list(csv.reader(open(path)))
    start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr = rows[start]
    idx = {h: i for i, h in enumerate(hdr)}
    per = collections.defaultdict(list)
    for r in rows[start + 1:]:
        if len(r) < len(hdr) or r[idx["Metric Name"]] != "gpu__time_duration.sum":
            continue
        name = re.sub(r"^void ", "", r[idx["Kernel Name"]].split("(")[0])
        val = float(r[idx["Metric Value"]].replace(",", ""))
        unit = r[idx["Metric Unit"]]
        val = val / 1000 if unit in ("ns", "nsecond") else val * 1000 if unit in ("ms", "msecond") else val
        per[(name, r[idx["Grid Size"]], r[idx["Block Size"]])].append(val)
    tot = sum(sum(v) for v in per.values())
    print("| kernel | grid | block | launches | avg us | share of captured GPU time |")
    print("|---|---|---|---:|---:|---:|")
    for k, v in sorted(per.items(), key=lambda kv: -sum(kv[1])):
        print(f"| `{k[0]}` | {k[1]} | {k[2]} | {len(v)} | {sum(v) / len(v):.2f} | {100 * sum(v) / tot:.1f}% |")
    print(f"\ntotal captured: {tot:.1f} us over {sum(len(v) for v in per.values())} launches "
          "(ncu serialises launches and runs them cold: compare SHARES, not absolutes)")


KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg", "launch__grid_size", "launch__block_size",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32_sparsity_off.sum",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed",
    "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__sass_inst_executed_op_utcmma.sum",
    "smsp__sass_inst_executed_op_tmem_ldt.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum",
]


def report(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        print(f"### `{r[col['Kernel Name']]}`\n")
        print("| metric | value | unit |\n|---|---:|---|")
        for k in KEYS:
            if k in col:
                print(f"| {k} | {r[col[k]]} | {units[col[k]]} |")
        stalls = [(h, r[i]) for h, i in col.items() if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio")]
        stalls = sorted(((h, float(v)) for h, v in stalls if v not in ("", "n/a")), key=lambda kv: -kv[1])[:6]
        if stalls:
            print("\ntop warp-stall reasons (warps stalled per issue-active cycle):\n")
            for h, v in stalls:
                print(f"* {h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')}: {v:.2f}")
        print()


def traffic(path):
    """profiles/traffic.json: per kernel (name up to '<'), mean dram bytes per launch over the captured launches."""
    import json
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    acc = collections.defaultdict(list)
    for r in rows[2:]:
        name = re.sub(r"^void ", "", r[col["Kernel Name"]]).split("<")[0].split("(")[0].replace("tt::", "")
        tot = 0.0
        for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            tot += float(r[col[k]].replace(",", "")) * scale[units[col[k]]]
        acc[name].append((tot, float(r[col["gpu__time_duration.sum"]].replace(",", "")), units[col["gpu__time_duration.sum"]]))
    res = {k: {"dram_bytes_per_launch": sum(x[0] for x in v) / len(v), "launches": len(v),
               "ncu_duration": sum(x[1] for x in v) / len(v), "ncu_duration_unit": v[0][2], "source": path.split("/")[-1]}
           for k, v in acc.items()}
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    {"launches": launches, "report": report, "traffic": traffic}[sys.argv[1]](sys.argv[2])
