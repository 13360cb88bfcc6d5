#!/usr/bin/env python
"""Read an .ncu-rep here (no GPU): headline metrics + the hottest SASS instructions with their stall reasons.
usage: ncu_top.py file.ncu-rep [n_top] [kernel_substring]"""
import csv, io, subprocess, sys

rep = sys.argv[1]
ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 25
sub = sys.argv[3] if len(sys.argv) > 3 else ""
WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread", "sm__cycles_active.avg",
        "sm__cycles_elapsed.max", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "l1tex__t_bytes_pipe_lsu_mem_global_op_st.sum",
        "lts__t_sectors_op_write.sum", "lts__t_sectors_srcunit_tex_op_write.sum"]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
for r in rows[2:]:
    name = r[idx["Kernel Name"]]
    if sub and sub not in name:
        continue
    print("=====", name[:110])
    for w in WANT:
        if w in idx:
            print(f"  {w:75s} {r[idx[w]]} {units[idx[w]]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
blocks, cur = [], None
for r in csv.reader(io.StringIO(src)):
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}
        blocks.append(cur)
    elif cur is not None:
        cur["rows"].append(r)
seen = set()
for b in blocks:
    if (sub and sub not in b["name"]) or not b["rows"]:
        continue
    h = b["rows"][0]
    ix = {c: i for i, c in enumerate(h)}
    data = [r for r in b["rows"][1:] if len(r) == len(h)]
    key = (b["name"], len(data), sum(int(r[ix["# Samples"]]) for r in data))
    if key in seen:
        continue
    seen.add(key)
    stalls = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]
    tot = key[2]
    agg = {s: sum(int(r[ix[s]]) for r in data) for s in stalls}
    print("=====", b["name"][:110])
    print("  samples", tot, " ".join(f"{k[6:]}={100*v/max(tot,1):.0f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:7]))
    top = sorted(range(len(data)), key=lambda i: -int(data[i][ix["# Samples"]]))[:ntop]
    for i in sorted(top):
        r = data[i]
        st = sorted(((int(r[ix[s]]), s[6:]) for s in stalls), reverse=True)[:2]
        print(f"  {i:5d} {r[ix['Source']][:60]:60s} samp {int(r[ix['# Samples']]):6d} exec {r[ix['Instructions Executed']]:>9s} {st[0][1]}={st[0][0]} {st[1][1]}={st[1][0]}")
