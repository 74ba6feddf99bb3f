"""Attribute the executed instructions of one kernel (ncu source-page CSV, sass) to CUDA source lines (nvdisasm -g -c).
usage: python tools/src_lines.py <sass csv> <cubin> <kernel> [top]"""
import re, csv, collections, subprocess, sys
csv_path, cubin, kernel = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
txt = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
start = next(i for i, l in enumerate(txt) if l.strip() == ".text.%s:" % kernel)
cur = None; instr = []
for ln in txt[start + 1:]:
    if ln.startswith(".text.") or ln.strip().startswith(".section"):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
    if m:
        inl = re.search(r'inlined at "([^"]+)", line (\d+)', m.group(3))
        cur = (m.group(1).split("/")[-1], int(m.group(2)), (inl.group(1).split("/")[-1], int(inl.group(2))) if inl else None)
        continue
    m = re.search(r'/\*([0-9a-f]{4,})\*/\s+(.*?);', ln)
    if m: instr.append((m.group(2).strip(), cur))
rows = list(csv.reader(open(csv_path)))
kern = {}; c = None; hdr = None
for r in rows:
    if r and r[0] == "Kernel Name": c = r[1]; kern[c] = []; continue
    if r and r[0] == "Address": hdr = r; continue
    if c and len(r) > 8: kern[c].append(r)
rs = kern[kernel]; iS = hdr.index("Source"); iI = hdr.index("Instructions Executed")
print("sass rows", len(rs), "disassembled", len(instr))
n = min(len(rs), len(instr))
agg = collections.Counter(); fp = 0
for (t, li), r in zip(instr[:n], rs[:n]):
    k = int(r[iI])
    if any(x in r[iS] for x in ("DFMA", "DMUL", "DADD")): fp += k; continue
    agg[(li[0], li[1]) if li else ("?", 0)] += k
tot = sum(agg.values())
print("FP64 %.0fM, other %.0fM" % (fp / 1e6, tot / 1e6))
for k, v in agg.most_common(top): print("%-24s %5d  %7.1fM  %4.1f%%" % (k[0], k[1], v / 1e6, 100 * v / tot))
