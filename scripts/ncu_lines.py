"""Per-source-line stall samples with their reasons from an .ncu-rep: python scripts/ncu_lines.py <rep> <kernel regex> [top]"""
import csv, io, subprocess, sys, collections
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", f"regex:{kern}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = None; cur = None; lines = []; tot = collections.Counter()
for r in rows:
    if len(r) >= 2 and r[0] in ("File Path", "File Name"):
        cur = r[1].split("/")[-1]; continue
    if len(r) > 5 and r[0] == "Line No":
        hdr = r
        iS = hdr.index("# Samples"); iE = hdr.index("Instructions Executed")
        st = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
        continue
    if not hdr or len(r) != len(hdr) or r[0] == "":
        continue
    try:
        n = int(r[iS])
    except ValueError:
        continue
    reasons = {h[6:]: int(r[i]) for i, h in st if r[i] not in ("", "0")}
    for k, v in reasons.items(): tot[k] += v
    lines.append((n, cur, r[0], r[1].strip()[:90], reasons, r[iE]))
total = sum(l[0] for l in lines)
print("total samples", total, "by reason:", ", ".join(f"{k} {v} ({100*v/max(total,1):.0f}%)" for k, v in tot.most_common(10)))
for n, f, ln, txt, reasons, ie in sorted(lines, key=lambda l: -l[0])[:top]:
    rs = " ".join(f"{k}:{v}" for k, v in sorted(reasons.items(), key=lambda kv: -kv[1])[:3])
    print(f"{n:6d} {100*n/max(total,1):4.1f}%  {f}:{ln:>5s} [{rs}] inst={ie} | {txt}")
