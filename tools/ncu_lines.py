"""Aggregate warp-stall samples of one profiled launch by CUDA source line (read here, no GPU):
   ncu -i REP --page source --csv --print-source cuda,sass --launch-skip K --launch-count 1 > x.csv; python tools/ncu_lines.py x.csv [N]"""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = None; cur = None; fname = ""
samples = collections.Counter(); insts = collections.Counter(); text = {}
per = collections.defaultdict(collections.Counter)
for r in rows:
    if not r: continue
    if r[0] == 'File Path': fname = r[1].split('/')[-1]; continue
    if r[0] == 'Function Name': continue
    if r[0] == 'Line No':
        hdr = r; iS = hdr.index('Warp Stall Sampling (All Samples)'); iN = hdr.index('Instructions Executed')
        stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
        continue
    if hdr is None: continue
    if r[0] != '':
        cur = (fname, int(r[0])); text[cur] = ','.join(r[1:len(r) - len(hdr) + 4])[:120]; continue
    # sass row: columns may be shifted by commas inside the source text; align from the right
    off = len(r) - len(hdr)
    try:
        s = int(r[iS + off]); n = int(r[iN + off])
    except Exception:
        continue
    samples[cur] += s; insts[cur] += n
    for i, h in stall_cols:
        try: per[cur][h] += int(r[i + off])
        except Exception: pass
tot = sum(samples.values())
print('total samples', tot, 'total warp insts', sum(insts.values()))
for ln, s in samples.most_common(top_n):
    top = ', '.join(f"{k[6:]}:{v}" for k, v in per[ln].most_common(3))
    print(f"{ln[0][:14]:14s}:{ln[1]:5d} {s:7d} {100*s/max(tot,1):5.1f}% inst {insts[ln]:9d}  [{top}]  {text.get(ln,'')[:80]}")
