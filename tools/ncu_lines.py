"""Aggregate an `ncu --page source --print-source cuda,sass --csv` dump by CUDA source line."""
import csv, sys
path=sys.argv[1]; top=int(sys.argv[2]) if len(sys.argv)>2 else 30
rows=list(csv.reader(open(path)))
cur_file=None; hdr=None; out=[]; seen_kernels=0; kernel=None
for r in rows:
    if len(r)==2 and r[0]=="File Path": cur_file=r[1].split("/")[-1]; continue
    if len(r)==2 and r[0]=="Function Name":
        if kernel is None: kernel=r[1]
        continue
    if len(r)>5 and r[0]=="Line No": hdr=r; continue
    if hdr is None or len(r)<len(hdr): continue
    if r[0]=="" : continue   # sass line
    si=hdr.index("# Samples"); ii=hdr.index("Instructions Executed")
    try: out.append((int(r[si]), int(r[ii]), cur_file, r[0], r[1].strip()[:100]))
    except ValueError: pass
# the dump repeats per profiled kernel instance; keep the first occurrence of each (file,line)
first={}
for s,n,f,l,src in out:
    first.setdefault((f,l),(s,n,src))
tot=sum(v[0] for v in first.values()); toti=sum(v[1] for v in first.values())
print(kernel); print("total samples",tot,"warp-instr",toti)
for (f,l),(s,n,src) in sorted(first.items(), key=lambda kv:-kv[1][0])[:top]:
    print(f"{100*s/tot:5.1f}% smp {100*n/max(toti,1):5.1f}% ins {f}:{l:>4s} {src}")
