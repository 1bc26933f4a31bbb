set -x
cd $GRAFT_REPO_ROOT
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/r02o_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02o_pytest.log
timeout 600 python tools/config_bench.py --only hubble,airs1024,ta_psd > gpurun_out/r02o_cfg_prod.log 2>&1; grep -v "^{" gpurun_out/r02o_cfg_prod.log | cut -c1-130
export PAOS_LIB=$PWD/paos_b200/libpaos_b200_tma1024.so
timeout 600 python -m pytest tests/test_gpu_batch.py tests/test_gpu_chains.py -m gpu -q -x > gpurun_out/r02o_pytest_1024.log 2>&1; echo "pytest 1024 rc=$?"; tail -2 gpurun_out/r02o_pytest_1024.log
timeout 600 python tools/config_bench.py --only hubble,airs1024,ta_psd > gpurun_out/r02o_cfg_1024.log 2>&1; grep -v "^{" gpurun_out/r02o_cfg_1024.log | cut -c1-130
