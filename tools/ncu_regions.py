"""Aggregate an ncu source page by SASS region: python tools/ncu_regions.py rep.ncu-rep [window]"""
import csv, io, re, subprocess, sys
from collections import Counter
rep = sys.argv[1]; W = int(sys.argv[2]) if len(sys.argv) > 2 else 200
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[1]; ci = {h: i for i, h in enumerate(hdr)}; data = rows[2:]
tot = sum(int(r[ci['# Samples']]) for r in data); toti = sum(int(r[ci['Instructions Executed']]) for r in data)
print(rows[0][1][:80], "samples", tot, "warp-instr", toti, "static", len(data))
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
for s in range(0, len(data), W):
    seg = data[s:s + W]
    samp = sum(int(r[ci['# Samples']]) for r in seg); inst = sum(int(r[ci['Instructions Executed']]) for r in seg)
    if samp < 0.005 * tot: continue
    st = {h: sum(int(r[ci[h]]) for r in seg) for h in stalls}
    top = sorted(st.items(), key=lambda x: -x[1])[:4]
    ops = Counter(re.sub(r'^@!?U?P\d+\s+', '', r[ci['Source']].strip()).split()[0] for r in seg).most_common(4)
    print(f"{s:5d}-{s+W:5d} samp {100*samp/tot:5.1f}% instr {100*inst/toti:5.1f}%  {[(k[6:], round(100*v/max(samp,1))) for k, v in top]} {ops}")
