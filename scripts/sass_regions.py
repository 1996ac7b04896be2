"""Segment the ncu SASS source page of a kernel into regions of similar execution count.
    ncu -i x.ncu-rep --page source --csv --print-source sass > x.csv ; python scripts/sass_regions.py x.csv
"""
import csv, sys
rows=list(csv.reader(open(sys.argv[1])))
hdr=rows[1]; data=rows[2:]
ia=hdr.index('Instructions Executed'); it=hdr.index('Thread Instructions Executed'); isrc=hdr.index('Source'); ismp=hdr.index('# Samples')
tot=sum(int(r[ia]) for r in data); tott=sum(int(r[it]) for r in data); tots=sum(int(r[ismp]) for r in data)
print('total warp inst',tot,'thread inst',tott,'avg thr',tott/tot, 'samples',tots)
def close(a,b):
    return abs(a-b)<=0.15*max(a,b,1)
seg=[]; start=0; cur=int(data[0][ia])
for i,r in enumerate(data):
    n=int(r[ia])
    if not close(n,cur):
        seg.append((start,i)); start=i; cur=n
    else:
        cur=0.9*cur+0.1*n
seg.append((start,len(data)))
thr=float(sys.argv[2]) if len(sys.argv)>2 else 0.004
for a,b in seg:
    w=sum(int(r[ia]) for r in data[a:b]); t=sum(int(r[it]) for r in data[a:b]); s=sum(int(r[ismp]) for r in data[a:b])
    if w/tot>thr:
        print('%5d-%5d n=%4d  warp%% %5.2f  thr/inst %5.1f  execs/inst %9.0f samples%% %5.2f | %s'%(a,b,b-a,100*w/tot,t/max(w,1),w/(b-a),100*s/tots,data[a][isrc].strip()[:60]))
