set -x
cd $GRAFT_REPO_ROOT
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/r02s_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02s_pytest.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/r02s_bench.json 2>/dev/null; python -c "
import json;d=json.load(open('gpurun_out/r02s_bench.json'));print('airs', d['value'], d['roofline']['frac'], d['e2e']['value'], d['e2e_ee']['value']); print({k:round(v['avg_us']) for k,v in d['passes'].items()})"
for c in hubble fgs1 ta_psd grid_sag; do
  timeout 900 python bench.py --config $c > gpurun_out/r02s_cfg_$c.json 2> gpurun_out/r02s_cfg_$c.err; echo "config $c rc=$?"
  python -c "
import json;d=json.load(open('gpurun_out/r02s_cfg_$c.json'));print('$c', round(d['value']), round(d['value_records_kept']['value']), round(d['e2e']['value']), d['parity'] and d['parity']['ok'])"
done
