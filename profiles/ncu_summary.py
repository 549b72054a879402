import csv, sys
raw, src = sys.argv[1], sys.argv[2]
rows=list(csv.reader(open(raw)))
hdr=rows[0]; idx={h:i for i,h in enumerate(hdr)}
keys=['gpu__time_duration.sum','smsp__inst_executed.sum','smsp__thread_inst_executed_per_inst_executed.ratio','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','smsp__warps_eligible.avg.per_cycle_active','dram__bytes_read.sum','dram__bytes_write.sum','l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct','launch__registers_per_thread','launch__occupancy_limit_registers','sm__maximum_warps_per_active_cycle_pct']
for r in rows[2:]:
    print('##', r[idx['Kernel Name']][:50])
    for k in keys:
        if k in idx: print('   ',k,'=',r[idx[k]])
rows=list(csv.reader(open(src)))
kern=[]; cur=None
for r in rows:
    if r and r[0]=='Kernel Name': cur={'name':r[1],'rows':[]}; kern.append(cur); continue
    if r and r[0]=='Address': cur['hdr']=r; continue
    if cur is not None and r: cur['rows'].append(r)
for K in kern:
    h={n:i for i,n in enumerate(K['hdr'])}
    print('=====',K['name'][:50])
    data=[(r[h['Source']].strip(),int(r[h['Instructions Executed']]),int(r[h['Thread Instructions Executed']]),int(r[h['# Samples']])) for r in K['rows']]
    tot=sum(d[1] for d in data); smp=sum(d[3] for d in data); ttot=sum(d[2] for d in data)
    print('total warp-inst %.1fM thread-inst %.1fM avg %.2f'%(tot/1e6,ttot/1e6,ttot/max(tot,1)))
    segs=[]; start=0
    for i in range(1,len(data)+1):
        if i==len(data) or abs(data[i][1]-data[start][1])>0.03*max(data[start][1],1) :
            ie=sum(d[1] for d in data[start:i]); te=sum(d[2] for d in data[start:i]); sm=sum(d[3] for d in data[start:i])
            segs.append((start,i,ie,te,sm)); start=i
    for s,e,ie,te,sm in segs:
        if ie>0.02*tot or sm>0.03*smp:
            ops={}
            for d in data[s:e]:
                op=d[0].split()[0] if not d[0].startswith('@') else d[0].split()[1]
                ops[op]=ops.get(op,0)+1
            top=sorted(ops.items(),key=lambda x:-x[1])[:5]
            print(f'[{s:4d},{e:4d}) n={e-s:3d} exec/inst={data[s][1]/1e6:6.2f}M share={100*ie/tot:5.1f}% avgthr={te/max(ie,1):5.1f} samples={100*sm/smp:5.1f}%  {top}')
