"""Per-source-line shares of executed instructions and stall samples from `ncu --page source --csv --print-source cuda,sass`."""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 50
sec = None; hdr = None; per = collections.defaultdict(lambda: [0, 0]); srcs = {}; cur = None
for r in rows:
    if not r: continue
    if r[0] == "File Path": sec = r[1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None: continue
    ln, csrc, addr = r[0], r[1], r[2]
    iex = hdr.index('Instructions Executed'); ismp = hdr.index('# Samples')
    if ln: cur = (sec.split('/')[-1], int(ln)); srcs[cur] = csrc
    if addr and cur:
        try:
            per[cur][0] += int(r[iex] or 0); per[cur][1] += int(r[ismp] or 0)
        except ValueError:
            pass
tot = sum(v[0] for v in per.values()); ts = sum(v[1] for v in per.values())
print("total warp-instructions", tot, "samples", ts)
for k, v in sorted(per.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{k[0]}:{k[1]:4d} {100*v[0]/tot:5.1f}% smp {100*v[1]/max(ts,1):5.1f}%  {srcs[k].strip()[:105]}")
