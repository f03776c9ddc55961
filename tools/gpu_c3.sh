mkdir -p gpurun_out
export PATH=/usr/local/cuda/bin:$PATH
timeout 600 python tools/c3_probe.py 2>&1 | tail -n 5
C3_N=1000000 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_c3_launches.csv python tools/c3_probe.py > gpurun_out/r02_c3_ncu.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv,collections
rows=list(csv.reader(l for l in open('gpurun_out/r02_c3_launches.csv') if l.startswith('"')))
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
agg=collections.OrderedDict()
for r in rows[1:]:
    try: v=float(r[vi].replace(',',''))
    except: continue
    a=agg.setdefault(r[ki][:70],[0,0.0]); a[0]+=1; a[1]+=v
for k,(c,t) in sorted(agg.items(), key=lambda x:-x[1][1])[:12]: print('%-70s n=%3d total %.2f ms avg %.1f us'%(k,c,t/1e6,t/c/1e3))
PY
