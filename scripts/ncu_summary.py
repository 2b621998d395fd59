"""Summarise an ncu report: raw metrics per kernel + instruction/stall attribution per source line
(deduplicated by SASS address).  usage: python scripts/ncu_summary.py report.ncu-rep [kernel-regex] [topN]"""
import csv, sys, subprocess, collections, io
rep = sys.argv[1]; kre = sys.argv[2] if len(sys.argv) > 2 else ""; top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
def run(args):
    return subprocess.run(["ncu", "-i", rep] + args, capture_output=True, text=True).stdout
raw = list(csv.reader(io.StringIO(run(["--page", "raw", "--csv"]))))
hdr, units = raw[0], raw[1]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'smsp__inst_executed.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts.sum', 'smsp__inst_executed_op_shared_atom.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
        'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum']
for r in raw[2:]:
    name = r[hdr.index('Kernel Name')]
    if kre and kre not in name: continue
    print("kernel:", name[:90])
    for w in want:
        if w in hdr:
            i = hdr.index(w); print(f"  {w:72s} {r[i]:>22s} {units[i]}")
    st = [(float(r[i]), h) for i, h in enumerate(hdr) if 'issue_stalled' in h and h.endswith('per_issue_active.ratio') or False]
    st = [(float(r[i]), h) for i, h in enumerate(hdr) if h.startswith('smsp__average_warps_issue_stalled_') and h.endswith('_per_issue_active.ratio') and 'not_issued' not in h]
    print("  stalls per issue:", ", ".join(f"{h.split('stalled_')[1].replace('_per_issue_active.ratio','')}={v:.2f}" for v, h in sorted(st, reverse=True)[:7]))
    src = list(csv.reader(io.StringIO(run(["--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + name.split('<')[0].split('(')[0].split()[-1].split('::')[-1]]))))
    seen = {}; fname = None; h2 = None; cur = None; text = {}
    for x in src:
        if not x: continue
        if x[0] == "File Path": fname = x[1].split('/')[-1]; continue
        if x[0] == "Function Name": continue
        if x[0] == "Line No": h2 = x; continue
        if h2 is None: continue
        ai = h2.index("Address"); ii = h2.index("Instructions Executed"); si = h2.index("# Samples"); wi = h2.index("L1 Wavefronts Shared")
        if x[0] != "": cur = (fname, int(x[0])); text[cur] = x[1]
        if len(x) > wi and x[ai] != "":
            try: v = (int(float(x[ii])), int(float(x[si])), int(float(x[wi] or 0)))
            except ValueError: continue
            seen.setdefault(x[ai], []).append((cur, v))
    agg = collections.defaultdict(lambda: [0, 0, 0])
    for a, lst in seen.items():
        pick = None
        for pref in ("spgemm_device", "spgemm_numeric", "spgemm_symbolic", "amb_"):
            for c in lst:
                if c[0][0].startswith(pref): pick = c; break
            if pick: break
        pick = pick or lst[0]
        for k in range(3): agg[pick[0]][k] += pick[1][k]
    ti = sum(v[0] for v in agg.values()) or 1; ts = sum(v[1] for v in agg.values()) or 1; tw = sum(v[2] for v in agg.values()) or 1
    print(f"  source lines (inst {ti:.3e}, samples {ts}, smem wavefronts {tw:.3e}):")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"   {k[0][:20]:20s} L{k[1]:<4d} samp {v[1]/ts*100:5.1f}% inst {v[0]/ti*100:5.1f}% smemwf {v[2]/tw*100:5.1f}%  {text[k].strip()[:95]}")
