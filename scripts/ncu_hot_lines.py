"""Summarise an `ncu --page source --csv --print-source cuda,sass` dump: hottest source lines of the first captured
launch (share of warp instructions, average active lanes, share of stall samples, top stall reasons).

    ncu -i rep.ncu-rep --page source --csv --print-source cuda,sass > src.csv ; python scripts/ncu_hot_lines.py src.csv [N]
"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
sections, cur, seen = [], None, set()
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur = {"file": r[1], "rows": []}
        if r[1] in seen:            # second launch starts: stop
            break
        seen.add(r[1]); sections.append(cur)
    elif len(r) > 5 and r[0] == "Line No":
        cur["hdr"] = r
    elif len(r) > 5 and cur is not None and "hdr" in cur:
        cur["rows"].append(r)


def num(v):
    try:
        return int(v)
    except ValueError:
        return 0


agg, allI, allS, allT = [], 0, 0, 0
for s in sections:
    h = s["hdr"]
    iS, iI, iT = h.index("# Samples"), h.index("Instructions Executed"), h.index("Thread Instructions Executed")
    stall_idx = {n: h.index(n) for n in h if n.startswith("stall_") and "Not Issued" not in n}
    for r in s["rows"]:
        if not r[0].isdigit():
            continue
        I, S, T = num(r[iI]), num(r[iS]), num(r[iT])
        if I == 0 and S == 0:
            continue
        st = {n: num(r[i]) for n, i in stall_idx.items()}
        agg.append((s["file"].split("/")[-1], int(r[0]), r[1].strip()[:100], I, T, S, st))
        allI += I; allS += S; allT += T
print("warp instructions %d  thread instructions %d  (%.1f lanes)  stall samples %d" % (allI, allT, allT / max(allI, 1), allS))
agg.sort(key=lambda a: -a[3])
for a in agg[:top_n]:
    top = sorted(a[6].items(), key=lambda kv: -kv[1])[:3]
    print("%-22s %4d  I %5.1f%%  lanes %4.1f  S %5.1f%% | %s | %s" % (
        a[0], a[1], 100 * a[3] / allI, a[4] / max(a[3], 1), 100 * a[5] / max(allS, 1), a[2],
        " ".join("%s=%d" % (k[6:], v) for k, v in top if v)))
tot = collections.Counter()
for a in agg:
    for k, v in a[6].items():
        tot[k] += v
print("stall share of all samples:", {k[6:]: round(100 * v / max(allS, 1), 1) for k, v in tot.most_common(12)})
by_file = collections.Counter()
for a in agg:
    by_file[a[0]] += a[3]
print("instructions by file:", {k: round(100 * v / allI, 1) for k, v in by_file.most_common(8)})
