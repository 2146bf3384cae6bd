"""Where one stall reason sits in an ncu capture (development aid): python tools/ncu_stall_where.py rep.ncu-rep stall_no_inst [top]"""
import csv, subprocess, sys
rep, col = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
src = subprocess.check_output(["ncu", "-i", rep, "--page", "source", "--csv"], text=True, stderr=subprocess.DEVNULL)
rows = list(csv.reader(src.split("\n")))
hi = 0 if 'Address' in rows[0] else 1
hdr = rows[hi]; ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[hi + 1:] if len(r) > 10]
S = lambda r: int(r[ix['# Samples']] or 0)
N = lambda r: int(r[ix['Instructions Executed']] or 0)
V = lambda r: int(r[ix[col]] or 0)
tot = sum(map(S, data)); nmax = max(map(N, data))
hot = [r for r in data if N(r) > 0.6 * nmax]
print("samples %d, %s %d (%.1f%%); inside the hot loop %d of its %d samples" % (tot, col, sum(map(V, data)), 100.0 * sum(map(V, data)) / tot, sum(map(V, hot)), sum(map(S, hot))))
for i, r in enumerate(data):
    r.append(i)
for r in sorted(data, key=V, reverse=True)[:top]:
    print("%6d of %6d samples  execs %.3g  #%d %s  %s" % (V(r), S(r), N(r), r[-1], r[ix['Address']] if 'Address' in ix else '', r[ix['Source']][:80]))
