import csv,sys
rows=list(csv.reader(open(sys.argv[1])))
pat=sys.argv[2]; chunk=int(sys.argv[3]) if len(sys.argv)>3 else 10
his=[i for i,r in enumerate(rows) if r and r[0]=='Address']
names=[rows[i-1][1] if i>0 else '' for i in his]
for n,hi in enumerate(his):
    if pat not in names[n]: continue
    hdr=rows[hi]; end=his[n+1]-1 if n+1<len(his) else len(rows)
    data=[r for r in rows[hi+1:end] if len(r)==len(hdr)]
    ia=hdr.index('Address'); isrc=hdr.index('Source'); iex=hdr.index('Instructions Executed'); ismp=hdr.index('# Samples'); ith=hdr.index('Avg. Threads Executed')
    tot=sum(int(r[iex]) for r in data); tots=sum(int(r[ismp]) for r in data)
    print(names[n][:60],'total warp-instr',tot,'samples',tots,'n',len(data))
    for k in range(0,len(data),chunk):
        seg=data[k:k+chunk]
        a=sum(int(r[iex]) for r in seg); sm=sum(int(r[ismp]) for r in seg)
        if a/tot<0.004: continue
        thr=sum(float(r[ith])*int(r[iex]) for r in seg)/max(1,a)
        print(seg[0][ia][-5:], f"{100*a/tot:5.1f}% instr {100*sm/tots:5.1f}% smp  thr {thr:4.1f}", seg[0][isrc][:60])
    break
