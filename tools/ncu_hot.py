"""Per-instruction stall breakdown of the hottest contiguous SASS region of an ncu capture (development aid).
   python tools/ncu_hot.py rep.ncu-rep [min_count_frac]"""
import csv, subprocess, sys
rep = sys.argv[1]
src = subprocess.check_output(["ncu", "-i", rep, "--page", "source", "--csv"], text=True, stderr=subprocess.DEVNULL)
rows = list(csv.reader(src.split("\n")))
hi = 0 if 'Address' in rows[0] else 1
hdr = rows[hi]; ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith('stall_')]
data = [r for r in rows[hi + 1:] if len(r) > 10]
S = lambda r: int(r[ix['# Samples']] or 0)
N = lambda r: int(r[ix['Instructions Executed']] or 0)
tot_s = sum(map(S, data)); tot_i = sum(map(N, data))
nmax = max(map(N, data))
print("total samples", tot_s, "inst %.4g" % tot_i)
hot = [r for r in data if N(r) > 0.6 * nmax]
print("hot-loop: %d instrs/iter, %.1f%% of samples, %.1f%% of instrs, iterations %.4g" % (len(hot), 100 * sum(map(S, hot)) / tot_s, 100 * sum(map(N, hot)) / tot_i, nmax))
agg = {}
for r in data:
    for h in stalls:
        v = int(r[ix[h]] or 0)
        if v: agg[h] = agg.get(h, 0) + v
print("kernel stall mix:", sorted(((v * 100 // tot_s, k[6:]) for k, v in agg.items()), reverse=True)[:8])
for r in hot:
    st = {h[6:]: int(r[ix[h]] or 0) for h in stalls if r[ix[h]] and int(r[ix[h]]) > 0}
    top = sorted(st.items(), key=lambda kv: -kv[1])[:3]
    print("%-70s s=%6d (%.2f%%) %s" % (r[ix['Source']][:70], S(r), 100 * S(r) / tot_s, top))
