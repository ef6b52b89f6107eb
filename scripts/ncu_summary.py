"""Summarises an `ncu --set full` capture: key raw metrics + per-phase (barrier-delimited) SASS statistics.
usage: python scripts/ncu_summary.py raw.csv src.csv n_elements"""
import csv, re, sys, collections
raw, src, nel = sys.argv[1], sys.argv[2], float(sys.argv[3])
rows = list(csv.reader(open(raw)))
d = dict(zip(rows[0], rows[2] if len(rows) > 2 else rows[1]))
keys = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__thread_inst_executed_per_inst_executed.ratio"]
for k in keys:
    print(k, d.get(k))
for k, v in d.items():
    if "issue_stalled" in k and k.endswith("per_issue_active.ratio"):
        try:
            if float(v) > 0.3: print("  stall", k.split("issue_stalled_")[1].split("_per_issue")[0], v)
        except ValueError: pass
rows = list(csv.reader(open(src)))
hdr = rows[1]
iI, iS, iSrc, iT = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Source"), hdr.index("Thread Instructions Executed")
iW, iWi = hdr.index("L1 Wavefronts Shared"), hdr.index("L1 Wavefronts Shared Ideal")
phase, acc, tot, ops = 0, {}, 0, {}
for r in rows[2:]:
    try: n, s, t = int(r[iI]), int(r[iS]), int(r[iT])
    except ValueError: continue
    w, wi = int(r[iW] or 0), int(r[iWi] or 0)
    a = acc.setdefault(phase, [0, 0, 0, 0, 0]); a[0] += n; a[1] += s; a[2] += t; a[3] += w; a[4] += wi
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[iSrc]); op = '.'.join((m.group(2) if m else "?").split('.')[:2])
    o = ops.setdefault(phase, collections.Counter()); o[op] += n
    tot += n
    if "BAR.SYNC" in r[iSrc]: phase += 1
for p, a in acc.items():
    print(f"phase {p}: warp-inst/elem {a[0]/nel:7.1f} ({a[0]/tot*100:4.1f}%) samples {a[1]:6d} thr/inst {a[2]/max(a[0],1):5.1f} shared wf/elem {a[3]/nel:6.1f} (ideal {a[4]/nel:6.1f})")
    print("     ", ", ".join(f"{k} {v/nel:.1f}" for k, v in ops[p].most_common(14)))
