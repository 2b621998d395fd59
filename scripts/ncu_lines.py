"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump per source line.
usage: ncu -i X.ncu-rep --page source --csv --print-source cuda,sass --kernel-name regex:K > f.csv; python scripts/ncu_lines.py f.csv [topN]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
agg = collections.OrderedDict()
fname = None; hdr = None; cur = None
for r in rows:
    if not r: continue
    if r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None: continue
    si = hdr.index("# Samples"); ii = hdr.index("Instructions Executed")
    wi = hdr.index("L1 Wavefronts Shared"); wid = hdr.index("L1 Wavefronts Shared Ideal")
    sb = {k: hdr.index(k) for k in ("stall_barrier", "stall_long_sb", "stall_short_sb", "stall_mio", "stall_wait", "stall_lg", "stall_branch_resolving")}
    if r[0] != "":
        cur = (fname, int(r[0])); agg.setdefault(cur, dict(src=r[1], samp=0, inst=0, wf=0, wfi=0, st=collections.Counter()))
    if len(r) > si and r[2] != "" and cur is not None:
        a = agg[cur]
        def f(x):
            try:
                return int(float(x))
            except ValueError:
                return 0
        a["samp"] += f(r[si]); a["inst"] += f(r[ii]); a["wf"] += f(r[wi]); a["wfi"] += f(r[wid])
        for k, i in sb.items(): a["st"][k] += f(r[i])
ts = sum(a["samp"] for a in agg.values()) or 1; ti = sum(a["inst"] for a in agg.values()) or 1; tw = sum(a["wf"] for a in agg.values()) or 1
print(f"total samples {ts} instructions {ti} smem wavefronts {tw}")
for (fn, ln), a in sorted(agg.items(), key=lambda kv: -kv[1]["samp"])[:top]:
    st = " ".join(f"{k[6:]}={v*100//max(a['samp'],1)}" for k, v in a["st"].most_common(3) if v)
    print(f"{fn[:22]:22s} L{ln:<4d} samp {a['samp']/ts*100:5.1f}% inst {a['inst']/ti*100:5.1f}% wf {a['wf']/tw*100:5.1f}% (x{a['wf']/max(a['wfi'],1):.1f}) [{st}] {a['src'].strip()[:90]}")
