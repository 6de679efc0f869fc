import csv,sys
def load(p):
    hdr=None; rows=[]
    for l in csv.reader(open(p)):
        if 'Kernel Name' in l: hdr=l; continue
        if hdr and len(l)==len(hdr) and l[0].isdigit(): rows.append(l)
    ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); gi=hdr.index('Grid Size'); ui=hdr.index('Metric Unit')
    ff=[(r[ki][:46], r[gi], float(r[vi].replace(',',''))/(1000 if r[ui]=='ns' else 1)) for r in rows if 'ffgp' in r[ki]]
    n=len(ff)//4
    return ff[3*n:]
a=load(sys.argv[1]); b=load(sys.argv[2])
for x,y in zip(a,b):
    print(f"{x[2]:8.1f} {y[2]:8.1f}  {x[1]:>16s} {y[1]:>14s}  {x[0]}")
print(sum(x[2] for x in a), sum(y[2] for y in b))
