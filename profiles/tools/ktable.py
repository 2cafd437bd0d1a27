import csv,collections,sys
rows=[r for r in csv.reader(open(sys.argv[1])) if len(r)>5]
hdr=rows[0]; ik=hdr.index('Kernel Name'); im=hdr.index('Metric Name'); iv=hdr.index('Metric Value'); iid=hdr.index('ID')
d=collections.OrderedDict()
for r in rows[1:]:
    key=(r[iid], r[ik].split('(')[0][:28]); d.setdefault(key,{})[r[im]]=float(r[iv].replace(',',''))
agg=collections.OrderedDict()
for (i,k),m in d.items():
    a=agg.setdefault(k,collections.Counter()); a['n']+=1
    for kk,v in m.items(): a[kk]+=v
print(f"{'kernel':28s} {'n':>3s} {'ms':>8s} {'rdMB':>8s} {'wrMB':>8s} {'Minst':>8s} {'issue%':>7s} {'fp64%':>6s} {'warps%':>7s}")
tot=0
for k,a in sorted(agg.items(), key=lambda x:-x[1]['gpu__time_duration.sum']):
    n=a['n']; tot+=a['gpu__time_duration.sum']/n/1e6
    if a['gpu__time_duration.sum']/n<5000: continue
    print(f"{k:28s} {n:3d} {a['gpu__time_duration.sum']/n/1e6:8.4f} {a['dram__bytes_read.sum']/n/1e6:8.1f} {a['dram__bytes_write.sum']/n/1e6:8.1f} {a['smsp__inst_executed.sum']/n/1e6:8.1f} {a['smsp__issue_active.avg.pct_of_peak_sustained_active']/n:7.1f} {a['sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active']/n:6.1f} {a['sm__warps_active.avg.pct_of_peak_sustained_active']/n:7.1f}")
print('sum ms', tot)
