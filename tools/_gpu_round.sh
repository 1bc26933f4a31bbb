set -x
cd $GRAFT_REPO_ROOT
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r02e_pytest.log 2>&1; echo "pytest rc=$?"
tail -12 gpurun_out/r02e_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02e_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r02e_smoke.log
timeout 900 python tools/config_bench.py --only hubble,airs512,airs1024,fgs1,ta_psd > gpurun_out/r02e_cfg_batched.log 2>&1; echo "cfg rc=$?"; grep -v "^{" gpurun_out/r02e_cfg_batched.log | cut -c1-260
timeout 900 python tools/config_bench.py --only hubble,airs512,airs1024,fgs1,ta_psd --batch 1 > gpurun_out/r02e_cfg_b1.log 2>&1; echo "cfg1 rc=$?"; grep -v "^{" gpurun_out/r02e_cfg_b1.log | cut -c1-260
timeout 900 python bench.py > gpurun_out/r02e_bench.json 2> gpurun_out/r02e_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r02e_bench.err
python -c "
import json;d=json.load(open('gpurun_out/r02e_bench.json'));print({k:d[k] for k in ('value','ms_per_step','gpu_launches','parity','cpu_baseline','e2e','e2e_ee','e2e_reduced')}); print(d['roofline'])"
