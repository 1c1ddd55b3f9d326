"""Per-source-line totals of an `ncu --page source --csv` export (SASS rows), using nvdisasm -g line markers of the same cubin.
usage: ncu_lines.py <sass_with_lineinfo> <mangled_kernel_name> <ncu_source_csv> [top]"""
import csv, re, sys
sass, kern, page = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
line_of = {}
cur, inside = None, False
for ln in open(sass):
    if ln.startswith(kern + ":"):
        inside = True; continue
    if inside:
        if ln.startswith("//---") or ln.lstrip().startswith(".section"):
            break
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+\S", ln)
        if m:
            line_of[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(page)))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
ci = {n: hdr.index(n) for n in ("# Samples", "Instructions Executed", "L1 Wavefronts Shared", "L1 Wavefronts Shared Ideal", "stall_barrier", "stall_short_sb", "stall_long_sb", "stall_wait", "stall_mio", "stall_lg")}
base = None
agg = {}
tot_s = tot_i = 0
for r in rows[hdr_i + 1:]:
    if len(r) < len(hdr) or not r[0].startswith("0x"):
        continue
    a = int(r[0], 16)
    if base is None:
        base = a
    key = line_of.get(a - base, ("?", 0))
    g = agg.setdefault(key, [0] * len(ci))
    for j, n in enumerate(ci):
        try:
            g[j] += int(float(r[ci[n]]))
        except ValueError:
            pass
    tot_s += int(float(r[ci["# Samples"]])); tot_i += int(float(r[ci["Instructions Executed"]]))
print("total samples %d, warp instructions %d" % (tot_s, tot_i))
print("%-28s %7s %6s %11s %6s | %s" % ("file:line", "samples", "%", "warp-inst", "%", " ".join(list(ci)[2:])))
for key, g in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%-28s %7d %6.2f %11d %6.2f | %s" % ("%s:%d" % key, g[0], 100.0 * g[0] / max(tot_s, 1), g[1], 100.0 * g[1] / max(tot_i, 1), " ".join(str(x) for x in g[2:])))
