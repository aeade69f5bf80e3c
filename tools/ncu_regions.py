"""Executed warp instructions and lanes per instruction of every SASS function region of lol_render in an .ncu-rep
(regions end at RET / EXIT): main kernel body, lol_sdf_slow, lol_near_collect, ...
    python tools/ncu_regions.py gpurun_out/x/prof.ncu-rep"""
import csv, io, subprocess, sys
rep=sys.argv[1]
out = subprocess.run(["ncu","-i",rep,"--page","source","--csv","--print-source","sass","--kernel-name","lol_render","--launch-count","1"],capture_output=True,text=True).stdout
rows=list(csv.reader(io.StringIO(out)))
hdr=next(r for r in rows if r and r[0]=="Address")
col={n:i for i,n in enumerate(hdr)}
I,T=col["Instructions Executed"],col["Thread Instructions Executed"]
ins=[]
for r in rows:
    if len(r)==len(hdr) and r[0].startswith("0x"):
        ins.append((int(r[0],16), r[1].strip(), int(r[I]), int(r[T])))
base=ins[0][0]
# regions split after RET / EXIT / BRA-to-self
regions=[]; cur=[]
for a,s,i,t in ins:
    cur.append((a,s,i,t))
    if s.startswith("RET") or s.startswith("EXIT"):
        regions.append(cur); cur=[]
if cur: regions.append(cur)
tot=sum(x[2] for x in ins)
print("total warp inst",tot, "lanes", sum(x[3] for x in ins)/tot)
for rg in regions:
    ri=sum(x[2] for x in rg); rt=sum(x[3] for x in rg)
    if ri==0: continue
    print(f"0x{rg[0][0]-base:06x}..0x{rg[-1][0]-base:06x} n={len(rg):5d} inst {100*ri/tot:6.2f}% lanes {rt/max(ri,1):5.1f}  first: {rg[0][1][:50]}")
