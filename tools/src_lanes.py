"""Per-kernel lane occupancy from an ncu source-page CSV (ncu -i rep --page source --csv --print-source sass)."""
import csv, sys
rows=list(csv.reader(open(sys.argv[1])))
kern={}; cur=None; hdr=None
for r in rows:
    if r and r[0]=="Kernel Name": cur=r[1]; kern[cur]=[]; continue
    if r and r[0]=="Address": hdr=r; continue
    if cur and len(r)>8: kern[cur].append(r)
iS=hdr.index("Source"); iI=hdr.index("Instructions Executed"); iT=hdr.index("Thread Instructions Executed"); iSm=hdr.index("# Samples")
def isfp(s):
    return any(x in s for x in("DFMA","DMUL","DADD","FFMA","FMUL","FADD"))
for k,rs in kern.items():
    tot=sum(int(r[iI]) for r in rs); tt=sum(int(r[iT]) for r in rs); ts=sum(int(r[iSm]) for r in rs)
    fp=[r for r in rs if isfp(r[iS])]
    fi=sum(int(r[iI]) for r in fp); ft=sum(int(r[iT]) for r in fp); fs=sum(int(r[iSm]) for r in fp)
    print(k,"inst",tot,"avg lanes %.2f"%(tt/tot),"| FP inst",fi,"(%.1f%%)"%(100*fi/tot),"lanes %.2f"%(ft/fi), "samples fp %.1f%%"%(100*fs/ts), "n_sass",len(rs))
