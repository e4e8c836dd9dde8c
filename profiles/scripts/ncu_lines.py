import csv, sys
path=sys.argv[1]; top=int(sys.argv[2]) if len(sys.argv)>2 else 45
rows=list(csv.reader(open(path)))
cur='?'; hdr=None; agg=[]
for r in rows:
    if len(r)>=2 and r[0]=='File Path': cur=r[1].split('/')[-1]; continue
    if 'Instructions Executed' in r: hdr=r; ie=hdr.index('Instructions Executed'); ss=hdr.index('# Samples'); te=hdr.index('Thread Instructions Executed'); continue
    if hdr is None or len(r)<=ie or r[0]=='' or r[2]!='-': continue
    try: n=int(r[ie])
    except: continue
    agg.append((n, cur, int(r[0]), r[1][:110], int(r[ss] or 0), int(r[te] or 0)))
tot=sum(a[0] for a in agg); tots=sum(a[4] for a in agg)
print('total warp instr', tot, 'samples', tots)
for a in sorted(agg, reverse=True)[:top]: print(f"{a[0]:>11d} {100*a[0]/tot:5.1f}% smp={100*a[4]/max(tots,1):5.1f}% thr={a[5]/max(a[0],1):5.1f} {a[1][:14]:14s}:{a[2]:<4d} {a[3]}")
