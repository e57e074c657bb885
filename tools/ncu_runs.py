"""Address-ordered view of one launch's stall samples: contiguous SASS runs that map to the same CUDA line.
   ncu -i REP --page source --csv --print-source sass ... is not line-correlated, so use the cuda,sass export and re-sort by address.
   usage: python tools/ncu_runs.py x.csv [min_samples]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
thr = int(sys.argv[2]) if len(sys.argv) > 2 else 100
hdr = None; cur = None; fname = ""
sass = []
for r in rows:
    if not r: continue
    if r[0] == 'File Path': fname = r[1].split('/')[-1]; continue
    if r[0] == 'Function Name': continue
    if r[0] == 'Line No':
        hdr = r; iS = hdr.index('Warp Stall Sampling (All Samples)'); iN = hdr.index('Instructions Executed'); iA = hdr.index('Address')
        stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
        continue
    if hdr is None: continue
    if r[0] != '': cur = (fname, int(r[0])); continue
    off = len(r) - len(hdr)
    try:
        addr = int(r[iA], 16) if off == 0 else int([c for c in r if c.startswith('0x')][0], 16)
        s = int(r[iS + off]); n = int(r[iN + off])
    except Exception:
        continue
    st = {h[6:]: int(r[i + off]) for i, h in stall_cols if r[i + off].isdigit() and int(r[i + off]) > 0}
    sass.append((addr, cur, s, n, r[iA + 1 + (0 if off == 0 else 0)][:60], st))
sass.sort()
tot = sum(x[2] for x in sass)
run = None
out = []
for addr, cur, s, n, txt, st in sass:
    if run and run['line'] == cur:
        run['s'] += s; run['n'] = max(run['n'], n)
        for k, v in st.items(): run['st'][k] = run['st'].get(k, 0) + v
    else:
        if run: out.append(run)
        run = {'line': cur, 's': s, 'n': n, 'addr': addr, 'st': dict(st)}
out.append(run)
base = sass[0][0]
acc = 0
for r in out:
    acc += r['s']
    if r['s'] >= thr:
        top = ', '.join(f"{k}:{v}" for k, v in sorted(r['st'].items(), key=lambda kv: -kv[1])[:3])
        print(f"+{r['addr']-base:6x} {r['line'][0][:12]}:{r['line'][1]:4d} samples {r['s']:6d} ({100*r['s']/tot:4.1f}%) cum {100*acc/tot:5.1f}% maxexec {r['n']:9d} [{top}]")
