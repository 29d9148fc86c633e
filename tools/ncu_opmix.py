"""Aggregate an `ncu --page source --csv --print-source sass` dump by SASS opcode."""
import csv, sys, collections
path = sys.argv[1]
kernel_filter = sys.argv[2] if len(sys.argv) > 2 else None
rows = list(csv.reader(open(path)))
blocks = []
cur = None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "rows": []}
        blocks.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = r
    elif cur is not None and r:
        cur["rows"].append(r)
sel = blocks[:1] if kernel_filter is None else [x for x in blocks if kernel_filter in x["name"]][:1]
for b in sel:
    h = b["hdr"]
    i_src, i_ex, i_samp = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
    stall_cols = [i for i, n in enumerate(h) if n.startswith("stall_")]
    mix = collections.Counter(); samp = collections.Counter(); tot = 0; tots = 0
    stalls = collections.Counter()
    for r in b["rows"]:
        try:
            ex = int(r[i_ex]); sm = int(r[i_samp])
        except ValueError:
            continue
        toks = r[i_src].split()
        op = toks[0] if toks else "?"
        if op.startswith("@") and len(toks) > 1:
            op = toks[1]
        keep2 = op.startswith(("LDS", "STS", "LDL", "STL", "LDG", "STG", "MUFU"))
        op = ".".join(op.split(".")[:2]) if keep2 else op.split(".")[0]
        mix[op] += ex; samp[op] += sm; tot += ex; tots += sm
        for i in stall_cols:
            try:
                stalls[h[i]] += int(r[i])
            except ValueError:
                pass
    print(b["name"][:100]); print("total warp-inst", tot, "samples", tots)
    for op, c in mix.most_common(30):
        print(f"{op:14s} {c:12d} {100*c/tot:5.1f}%   samples {100*samp[op]/max(tots,1):5.1f}%")
    print("stalls:", [(k, round(100*v/max(sum(stalls.values()),1),1)) for k, v in stalls.most_common(8)])
