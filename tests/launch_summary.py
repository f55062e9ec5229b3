"""Summarise an ncu gpu__time_duration launch list: per-kernel totals of the last N launches."""
import collections
import csv
import re
import sys

path, last_n = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0
lines = [l for l in open(path) if not l.startswith("==")]
rows = [(r["Kernel Name"], float(r["Metric Value"])) for r in csv.DictReader(lines)
        if r.get("Metric Name") == "gpu__time_duration.sum"]


def short(n):
    m = re.search(r"GemmCfgILi(\d)ELi(\d)ELi(\d+)ELb(\d)ELb(\d)ELb(\d)ELi(\d)ELi(\d)ELi(\d)ELi(\d)ELb(\d)", n)
    if m:
        cg, mt, bn, amn, bmn, dec, epi, st, ng, pst, xf = m.groups()
        return f"qlora_gemm<CG{cg} MT{mt} BN{bn} A_MN{amn} B_MN{bmn} DEC{dec} EPI{epi} XF{xf}>"
    m = re.search(r"GemmCfg<([^>]*)>", n)
    if m:
        return "qlora_gemm<" + m.group(1).replace(" ", "") + ">"
    return n.split("(")[0][:70]


sel = rows[-last_n:] if last_n else rows
agg = collections.OrderedDict()
for n, v in sel:
    k = short(n)
    agg.setdefault(k, [0, 0.0])
    agg[k][0] += 1
    agg[k][1] += v
tot = sum(v for _, v in sel)
print(f"{len(rows)} launches in file; summarising the last {len(sel)}; total {tot / 1e3:.1f} us")
for k, (c, v) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{k:75s} n={c:4d} total={v / 1e3:9.1f} us  avg={v / c / 1e3:8.1f} us  share={v / tot:6.1%}")
