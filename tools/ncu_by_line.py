"""Attribute executed instructions / stall samples of an ncu capture to source lines and functions.
   python tools/ncu_by_line.py gpurun_out/prof.ncu-rep pathtrace_rs_b200/lib/libptgpu.so pt_megakernel_constILi2E"""
import csv, re, subprocess, sys, os, tempfile, collections
rep, so, kern = sys.argv[1], sys.argv[2], sys.argv[3]
tmp = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, stdout=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.check_output(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], text=True)
# address -> (file, line) for the kernel
addr2line = {}
cur = None; inside = False
for l in dis.split("\n"):
    if l.startswith("//---") and ".text." in l:
        inside = kern in l
    if not inside: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/", l)
    if m: addr2line[int(m.group(1), 16)] = cur
src = subprocess.check_output(["ncu", "-i", rep, "--page", "source", "--csv"], text=True, stderr=subprocess.DEVNULL)
rows = list(csv.reader(src.split("\n")))
hdr = rows[1]; iA, iN, iI = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed")
base = None
per_line = collections.Counter(); per_line_s = collections.Counter()
for r in rows[2:]:
    if len(r) <= iI or not r[iA]: continue
    a = int(r[iA], 16) if r[iA].startswith("0x") else int(r[iA])
    if base is None: base = a
    key = addr2line.get(a - base)
    per_line[key] += int(r[iI] or 0); per_line_s[key] += int(r[iN] or 0)
tot = sum(per_line.values()); tots = sum(per_line_s.values())
# function ranges by source line (pt files)
def bucket(key):
    if key is None: return "?"
    f, ln = key
    return f
per_file = collections.Counter(); per_file_s = collections.Counter()
for k, v in per_line.items(): per_file[bucket(k)] += v
for k, v in per_line_s.items(): per_file_s[bucket(k)] += v
print("total instrs %.4g, samples %d" % (tot, tots))
for f, v in per_file.most_common(): print("%-32s instrs %5.1f%%  samples %5.1f%%" % (f, 100 * v / tot, 100 * per_file_s[f] / tots))
print("--- top lines")
for k, v in per_line.most_common(45): print("%-28s:%-5s instrs %5.2f%%  samples %5.2f%%" % (k[0] if k else "?", k[1] if k else "", 100 * v / tot, 100 * per_line_s[k] / tots))
