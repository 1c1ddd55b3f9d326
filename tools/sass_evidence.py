"""Per-kernel counts of the SASS instructions behind the Blackwell-specific pieces of the build:
    cuobjdump -sass gmql_b200/libgmqlcuda.so | python tools/sass_evidence.py > profiles/<round>_sass_evidence.txt"""
import collections
import re
import subprocess
import sys

PATS = [("UBLKCP", r"\bUBLKCP"), ("SYNCS", r"\bSYNCS"), ("UCGABAR", r"\bUCGABAR"), ("LDG.128", r"\bLDG\.E\.[A-Z.]*128"), ("LDS.128", r"\bLDS\.[A-Z.]*128"),
        ("ATOMS", r"\bATOMS"), ("REDG", r"\bRED\.|\bREDG"), ("ATOMG", r"\bATOMG"), ("CCTL", r"\bCCTL"), ("REDUX", r"\bREDUX")]
WANT = ("k_cover_tile_wc", "k_multisplit", "k_map_count_lean", "k_stitch_small", "k_map_index_small", "k_join_md_team", "k_join_scan", "k_gmap4_merge",
        "k_gmap4_finish_direct", "k_summit_bins", "k_group_rep", "k_os_pass", "k_dataset_stats", "k_format_text", "k_parse")


def main():
    counts = collections.OrderedDict()
    cur = None
    for ln in sys.stdin:
        m = re.match(r"\s*Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        for k, p in PATS:
            if re.search(p, ln):
                counts[cur][k] += 1
    names = list(counts)
    dem = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    print("# SASS evidence (cuobjdump -sass gmql_b200/libgmqlcuda.so, sm_100a): instruction counts per kernel")
    print("#   UBLKCP   TMA bulk copy (cp.async.bulk ... mbarrier::complete_tx::bytes);  SYNCS  mbarrier init / arrive.expect_tx / try_wait")
    print("#   UCGABAR  thread-block-cluster barrier;  LDG.128 / LDS.128  128-bit global / shared loads;  ATOMS  shared-memory atomic")
    print("#   REDG     global reduction without return value;  ATOMG  global atomic with return value;  CCTL  prefetch.global.L2;  REDUX  redux.sync")
    print()
    print("%-72s" % "kernel" + "".join("%9s" % k for k, _ in PATS))
    rows = []
    for n, d in zip(names, dem):
        if not any(w in d for w in WANT):
            continue
        head = d.split("(anonymous namespace)::")[-1] if "(anonymous namespace)::" in d.split("(")[0] + "(" else d
        head = re.sub(r"^void ", "", d)
        head = head.replace("(anonymous namespace)::", "")
        cut = head.find(">(") + 1 if ">(" in head else head.find("(")
        rows.append((head[:cut][:71], counts[n]))
    for short, c in sorted(rows):
        print("%-72s" % short + "".join("%9d" % c[k] for k, _ in PATS))


if __name__ == "__main__":
    main()
