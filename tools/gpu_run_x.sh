#!/bin/bash
# GPU call X: launch list of a NAIS-Net basket-100D step at M = 100 and M = 16384 (ncu gpu__time_duration)
mkdir -p gpurun_out
O=gpurun_out
for M in 100 16384; do
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/x_nais_m$M.csv \
  python bench.py --workload basket_nais --dim 100 --act Sine --paths $M --steps 2 --warmup 3 --skip-mc --skip-cpu --skip-small --skip-variants --skip-e2e --skip-workloads --skip-peak > $O/x_nais_m$M.log 2>&1
echo "rc=$?"
done
python - <<'PY'
import csv, collections
for M in (100, 16384):
    rows=[r for r in csv.reader(open(f'gpurun_out/x_nais_m{M}.csv')) if len(r)>10]
    hdr=rows[0]; ix={h:i for i,h in enumerate(hdr)}
    ks=[(r[ix['Kernel Name']].split('(')[0][:80], float(r[ix['Metric Value']].replace(',',''))) for r in rows[1:]]
    # last 63 launches = one step
    last=ks[-63:]
    by=collections.OrderedDict()
    for n,t in last:
        b=by.setdefault(n,[0,0.0]); b[0]+=1; b[1]+=t
    tot=sum(t for _,t in last)
    print(f"== M={M}: last 63 launches {tot/1e3:.1f} us")
    for n,b in sorted(by.items(), key=lambda x:-x[1][1]):
        print(f"  {b[0]:3d} {b[1]/1e3:9.1f} us  {n}")
PY
