set -x
cd $GRAFT_REPO_ROOT
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/r02i_pytest.log 2>&1; echo "pytest rc=$?"
tail -6 gpurun_out/r02i_pytest.log
for v in on off; do
  if [ $v = off ]; then export PAOS_NO_EDGE_TABLES=1; else unset PAOS_NO_EDGE_TABLES; fi
  timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/r02i_bench_$v.json 2> gpurun_out/r02i_bench_$v.err; echo "bench $v rc=$?"
  python -c "
import json;d=json.load(open('gpurun_out/r02i_bench_$v.json'));print('$v','value',d['value'],'roof',d['roofline']['frac'],'launches',d['gpu_launches']); print({k:round(v['avg_us']) for k,v in d['passes'].items()})"
done
