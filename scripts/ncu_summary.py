"""Condenses an ncu report (gpurun_out/*.ncu-rep) into the text summaries kept under profiles/.
usage: python scripts/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r1_name"""
import collections
import csv
import io
import json
import re
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__warps_eligible.avg.per_cycle_active", "smsp__warps_active.avg.per_cycle_active", "sm__cycles_elapsed.max",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct"]


def ncu(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def main():
    rep, out = sys.argv[1], sys.argv[2]
    rows = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units = rows[0], rows[1]
    lines = []
    traffic = {}
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        short = re.sub(r"\(.*", "", name).replace("void ", "").replace("smcb::", "")
        lines.append(f"== {name}")
        for k in KEYS:
            if k in hdr:
                lines.append(f"   {k:72s} {r[hdr.index(k)]:>16s} {units[hdr.index(k)]}")
        stalls = []
        for i, h in enumerate(hdr):
            if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued"):
                try:
                    stalls.append((float(r[i]), h.replace("smsp__pcsamp_warps_issue_stalled_", "")))
                except ValueError:
                    pass
        tot = sum(v for v, _ in stalls) or 1.0
        lines.append("   warp stall samples: " + ", ".join(f"{h} {100 * v / tot:.1f}%" for v, h in sorted(stalls, reverse=True)[:8]))
        try:
            rd = float(r[hdr.index("dram__bytes_read.sum")]); wr = float(r[hdr.index("dram__bytes_write.sum")])
            scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
            rd *= scale.get(units[hdr.index("dram__bytes_read.sum")], 1.0)
            wr *= scale.get(units[hdr.index("dram__bytes_write.sum")], 1.0)
            traffic.setdefault(short.split("<")[0], []).append(rd + wr)
        except Exception:
            pass
    # SASS opcode mix per kernel name
    for kname in sorted({re.sub(r"<.*", "", re.sub(r"\(.*", "", r[hdr.index("Kernel Name")]).replace("void ", "").replace("smcb::", "")) for r in rows[2:]}):
        src = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kname}"]))))
        if len(src) < 3:
            continue
        h2 = src[1]
        ci = {h: i for i, h in enumerate(h2)}
        ops = collections.Counter(); samp = collections.Counter(); launches = 0
        for r in src:
            if r and r[0] == "Kernel Name":
                launches += 1
            if len(r) < len(h2) or not r[ci["Instructions Executed"]].isdigit():
                continue
            m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[ci["Source"]])
            op = m.group(2).split(".")[0] if m else "?"
            ops[op] += int(r[ci["Instructions Executed"]]); samp[op] += int(r[ci["# Samples"]])
        tot = sum(ops.values()) or 1; ts = sum(samp.values()) or 1
        lines.append(f"== SASS mix {kname} ({launches} launches captured; warp instructions / launch = {tot // max(1, launches)})")
        for op, n in ops.most_common(16):
            lines.append(f"   {op:10s} {100 * n / tot:6.2f}% of instructions   {100 * samp[op] / ts:6.2f}% of stall samples")
    open(out + ".txt", "w").write("\n".join(lines) + "\n")
    json.dump({k: sum(v) / len(v) for k, v in traffic.items()}, open(out + "_traffic.json", "w"), indent=1)
    print("\n".join(lines))


if __name__ == "__main__":
    main()
