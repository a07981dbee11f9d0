#!/usr/bin/env python
"""Markdown summary of one ncu capture:  python scripts/ncu_summary.py X.ncu-rep [kernel regex]
Key counters (time, DRAM bytes, L2 reductions, instructions, issue utilisation, shared-memory wavefronts),
the warp-stall table, and the hottest source lines (needs -lineinfo + --import-source on)."""
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
kre = sys.argv[2] if len(sys.argv) > 2 else "scan_kernel"


def ncu(*args):
    return subprocess.run(["ncu", "-i", rep] + list(args), stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout


rows = list(csv.reader(io.StringIO(ncu("--page", "raw", "--csv"))))
hdr, units = rows[0], rows[1]
want = [
    ("gpu__time_duration.sum", "kernel time under ncu"),
    ("launch__grid_size", "grid"), ("launch__block_size", "threads per CTA"),
    ("launch__registers_per_thread", "registers per thread"),
    ("launch__shared_mem_per_block_dynamic", "dynamic shared memory per CTA"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active (% of 64 per SM)"),
    ("dram__bytes_read.sum", "DRAM bytes read"), ("dram__bytes_write.sum", "DRAM bytes written"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput (% of peak)"),
    ("lts__t_sectors_srcunit_tex_op_red.sum", "L2 sectors of reductions (RED)"),
    ("lts__t_sectors_srcunit_tex_op_red.sum.pct_of_peak_sustained_elapsed", "L2 RED sectors (% of peak)"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput (% of peak)"),
    ("smsp__inst_executed.sum", "warp instructions executed"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots used (%)"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "shared-memory wavefronts"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared-memory bank conflicts"),
    ("l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed", "LSU writeback active (%)"),
    ("smsp__inst_executed_op_shared_atom.sum", "shared atomics / reductions (warp instr.)"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe (% of peak)"),
]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    if not re.search(kre, d.get("Kernel Name", "")):
        continue
    u = dict(zip(hdr, units))
    print("kernel: `%s`\n" % d["Kernel Name"][:120])
    print("| metric | value |\n|---|---|")
    for k, label in want:
        if k in d and d[k] != "":
            print("| %s (`%s`) | %s %s |" % (label, k, d[k], u.get(k, "")))
    print("\n| warp stall reason | stalled warps per issued instruction |\n|---|---|")
    st = []
    for k, v in d.items():
        m = re.match(r"smsp__average_warps_issue_stalled_(\w+)_per_issue_active\.ratio", k)
        if m and v:
            st.append((float(v), m.group(1)))
    for v, k in sorted(st, reverse=True)[:10]:
        print("| %s | %.2f |" % (k, v))
    break

src = ncu("--page", "source", "--csv", "--print-source", "cuda,sass", "-k", "regex:" + kre)
rows = list(csv.reader(io.StringIO(src)))
hdr = None
lines = {}
for r in rows:
    if len(r) > 6 and r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) < len(hdr) or r[2] != "-":
        continue
    try:
        ln = int(r[0])
    except ValueError:
        continue
    e = lines.setdefault(ln, {"src": r[1], "samples": 0, "inst": 0, "stalls": {}})
    e["samples"] += int(r[hdr.index("# Samples")] or 0)
    e["inst"] += int(r[hdr.index("Instructions Executed")] or 0)
    for i, h in enumerate(hdr):
        if h.startswith("stall_") and "Not Issued" not in h:
            e["stalls"][h[6:]] = e["stalls"].get(h[6:], 0) + int(r[i] or 0)
tot = sum(e["samples"] for e in lines.values()) or 1
toti = sum(e["inst"] for e in lines.values()) or 1
if lines:
    print("\nhottest source lines (share of warp-stall samples, share of executed instructions, top stalls):\n")
    print("| samples | instr. | line | stalls | source |\n|---|---|---|---|---|")
    for ln, e in sorted(lines.items(), key=lambda kv: -kv[1]["samples"])[:18]:
        top = sorted(e["stalls"].items(), key=lambda kv: -kv[1])[:3]
        print("| %.1f%% | %.1f%% | %d | %s | `%s` |" % (100 * e["samples"] / tot, 100 * e["inst"] / toti, ln,
              ", ".join("%s %d%%" % (k, 100 * v / max(e["samples"], 1)) for k, v in top if v), e["src"].strip().replace("|", "\\|")[:80]))
