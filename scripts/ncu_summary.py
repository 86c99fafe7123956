"""Summarise an .ncu-rep (read here, no GPU): headline metrics, stall mix, opcode mix, hottest source lines."""
import collections, csv, io, subprocess, sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, unit = rows[0], rows[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__grid_size", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_sectors_op_write.sum", "lts__t_sectors_op_read.sum"]
for r in rows[2:]:
    print("kernel:", r[hdr.index("Kernel Name")][:90])
    for h, u, v in zip(hdr, unit, r):
        if h in want:
            print(f"  {h:70s} {v} {u}")
    st = [(float(v), h) for h, v in zip(hdr, r) if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio")]
    for v, h in sorted(st, reverse=True)[:8]:
        print(f"  stall {h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''):28s} {v:.3f}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
cur, hdr, lines, ops = None, None, [], collections.Counter()
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]; continue
    if len(r) > 5 and r[0] == "Line No":
        hdr = r; iE = hdr.index("Instructions Executed"); iS = hdr.index("# Samples"); continue
    if not hdr or len(r) <= iE:
        continue
    try:
        n, sm = int(r[iE]), int(r[iS])
    except ValueError:
        continue
    if r[0] != "":
        lines.append((n, sm, cur, r[0], r[1].strip()[:100]))
    elif r[3] not in ("", "..."):
        t = r[3].split()
        op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
        ops[op] += n
tot = sum(l[0] for l in lines) or 1
print(f"instructions attributed to source lines: {tot}")
print("opcode mix:", ", ".join(f"{o} {100 * n / max(sum(ops.values()), 1):.1f}%" for o, n in ops.most_common(14)))
print("hottest lines by instructions executed:")
for n, sm, f, ln, txt in sorted(lines, reverse=True)[:top]:
    print(f"  {n:10d} {100 * n / tot:5.1f}%  smp {sm:5d}  {f}:{ln}  {txt}")
print("hottest lines by stall samples:")
for n, sm, f, ln, txt in sorted(lines, key=lambda l: -l[1])[:12]:
    print(f"  {n:10d} smp {sm:5d}  {f}:{ln}  {txt}")
