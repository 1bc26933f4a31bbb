set -x
cd $GRAFT_REPO_ROOT
python -c "
import paos_b200; print(paos_b200._lib.lib.paos_build_info())"
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/r02k_pytest.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r02k_pytest.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/r02k_bench.json 2> gpurun_out/r02k_bench.err; echo "bench rc=$?"
python -c "
import json;d=json.load(open('gpurun_out/r02k_bench.json'));print('value',d['value'],'roof',d['roofline']['frac'],'launches',d['gpu_launches']); print({k:round(v['avg_us']) for k,v in d['passes'].items()})"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02k_launches.csv python bench.py --n-wl 16 --steps 1 --warmup 1 --no-cpu --slots 1 > gpurun_out/r02k_ncu1.log 2>&1; echo "ncu1 rc=$?"
python - <<'PY'
import csv
from collections import defaultdict
rows=[r for r in csv.reader(open('gpurun_out/r02k_launches.csv')) if r and not r[0].startswith('==')]
h=rows[0]; ki=h.index('Kernel Name'); vi=h.index('Metric Value')
agg=defaultdict(lambda:[0,0.0])
for r in rows[1:]:
    if len(r)>vi:
        a=agg[r[ki][:60]]; a[0]+=1; a[1]+=float(r[vi].replace(',',''))/1e3
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][1]): print(k, v[0], round(v[1],1), round(v[1]/v[0],1))
PY
