set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_products.py tests/test_gpu_chains.py tests/test_gpu_batch.py -m gpu -q -x > gpurun_out/r02t_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02t_pytest.log
for c in hubble fgs1 ta_psd; do
  timeout 900 python bench.py --config $c > gpurun_out/r02t_cfg_$c.json 2> gpurun_out/r02t_cfg_$c.err; echo "config $c rc=$?"
  python -c "
import json;d=json.load(open('gpurun_out/r02t_cfg_$c.json'));print('$c', round(d['value']), round(d['value_records_kept']['value']), round(d['e2e']['value']), d['parity'] and d['parity']['ok'])"
done
timeout 900 python bench.py > gpurun_out/r02t_bench.json 2> gpurun_out/r02t_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r02t_bench.err
python -c "
import json;d=json.load(open('gpurun_out/r02t_bench.json'));print({k:d[k] for k in ('value','ms_per_step','gpu_launches','wall_ms_per_psf','host_plan_ms_per_psf')}, d['parity']['worst'], d['parity']['ok'], d['cpu_baseline']['value']); print(d['e2e']['value'], d['e2e_ee']['value'], d['e2e_reduced']['value']); print(d['roofline']['frac'], d['roofline']['wall']['frac'], d['roofline']['l1_smem_pipe']['frac'])"
