"""Per-source-line instruction / sample shares of one kernel from an ncu report (source page) and the cubin's line table.
usage: ncu_hot.py <report.ncu-rep> <libgmqlcuda.so> <kernel substring> [top]"""
import csv, os, re, subprocess, sys, tempfile, glob
rep, so, kern = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, capture_output=True)
sass = None
for f in glob.glob(os.path.join(tmp, "*.cubin")):
    txt = subprocess.run(["nvdisasm", "-g", "-c", f], capture_output=True, text=True).stdout
    if kern in txt:
        sass = txt; break
page = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
line_of, cur, inside = {}, None, False
for ln in sass.splitlines():
    if ln.startswith(".text.") and ln.rstrip().endswith(":"):
        inside = kern in ln
        continue
    if inside:
        if ln.startswith("//---"):
            inside = False; continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+\S", ln)
        if m:
            line_of[int(m.group(1), 16)] = cur
rows = list(csv.reader(page.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
iS, iI = hdr.index("# Samples"), hdr.index("Instructions Executed")
iW = hdr.index("L1 Wavefronts Shared") if "L1 Wavefronts Shared" in hdr else None
base, agg, tot_i, tot_s = None, {}, 0, 0
for r in rows[hi + 1:]:
    if len(r) < len(hdr) or not r[0].startswith("0x"):
        continue
    a = int(r[0], 16)
    base = a if base is None else base
    key = line_of.get(a - base, ("?", 0))
    g = agg.setdefault(key, [0, 0, 0])
    g[0] += int(float(r[iS])); g[1] += int(float(r[iI])); g[2] += int(float(r[iW])) if iW is not None and r[iW] else 0
    tot_s += int(float(r[iS])); tot_i += int(float(r[iI]))
rr = list(csv.reader(raw.splitlines()))
vals = dict(zip(rr[0], rr[2] if len(rr) > 2 else rr[1]))
for k in ("gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
          "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
          "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread"):
    print(k, vals.get(k))
for k, v in vals.items():
    if "issue_stalled" in k and k.endswith("per_issue_active.ratio") and float(v or 0) > 0.3:
        print(" stall", k.split("issue_stalled_")[1].split("_per_")[0], v)
print("total samples %d, warp instructions %d" % (tot_s, tot_i))
src = {}
for key, g in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    text = ""
    for d in ("gmql_b200/csrc",):
        p = os.path.join(os.path.dirname(os.path.abspath(so)), "csrc", key[0])
        if os.path.exists(p):
            if p not in src:
                src[p] = open(p).read().splitlines()
            if 0 < key[1] <= len(src[p]):
                text = src[p][key[1] - 1].strip()[:100]
    print("%-30s %5.2f%% inst %5.2f%% smp %9d wf  %s" % ("%s:%d" % key, 100 * g[1] / max(tot_i, 1), 100 * g[0] / max(tot_s, 1), g[2], text))
