set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_ee.py tests/test_gpu_chains.py -m gpu -q -x > gpurun_out/r02d_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r02d_pytest.log
for v in tma tmarow2; do
  PAOS_LIB=$PWD/paos_b200/libpaos_b200_$v.so timeout 600 python -m pytest tests/test_gpu_batch.py tests/test_gpu_fullsize.py -m gpu -q -x > gpurun_out/r02d_pytest_$v.log 2>&1; echo "pytest $v rc=$?"; tail -2 gpurun_out/r02d_pytest_$v.log
  PAOS_LIB=$PWD/paos_b200/libpaos_b200_$v.so timeout 600 python bench.py --steps 2 --warmup 2 --no-cpu --batch 8 > gpurun_out/r02d_var_$v.json 2> gpurun_out/r02d_var_$v.err; echo "variant $v rc=$?"
  python -c "
import json;d=json.load(open('gpurun_out/r02d_var_$v.json'));print('$v','value',d['value'],'roof',d['roofline']['frac']); print({k:round(v['avg_us']) for k,v in d['passes'].items()})"
done
timeout 600 python bench.py --steps 2 --warmup 2 --no-cpu --batch 8 > gpurun_out/r02d_prod.json 2> gpurun_out/r02d_prod.err; echo "prod rc=$?"
python -c "
import json;d=json.load(open('gpurun_out/r02d_prod.json'));print('prod','value',d['value'],'roof',d['roofline']['frac']); print({k:round(v['avg_us']) for k,v in d['passes'].items()})"
