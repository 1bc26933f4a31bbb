set -x
cd $GRAFT_REPO_ROOT
python -c "
import paos_b200; print(paos_b200._lib.lib.paos_build_info())"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02m_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02m_smoke.log
timeout 900 python bench.py > gpurun_out/r02m_bench.json 2> gpurun_out/r02m_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r02m_bench.err
python -c "
import json;d=json.load(open('gpurun_out/r02m_bench.json'));print({k:d[k] for k in ('value','ms_per_step','gpu_launches','parity','cpu_baseline')}); print(d['e2e']['value'], d['e2e_ee']['value'], d['e2e_reduced']['value']); print(d['roofline']['frac'], d['roofline']['l1_smem_pipe'])"
CMD="python bench.py --n-wl 16 --steps 1 --warmup 1 --no-cpu --slots 1"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches.csv $CMD > gpurun_out/r02_ncu1.log 2>&1; echo "ncu1 rc=$?"
timeout 1200 ncu --set full --import-source on --clock-control none -k regex:pass_kernel -s 14 -c 14 -f -o gpurun_out/r02_prof $CMD > gpurun_out/r02_ncu2.log 2>&1; echo "ncu2 rc=$?"
timeout 900 python tools/config_bench.py > gpurun_out/r02m_cfg.log 2>&1; echo "cfg rc=$?"; grep -v "^{" gpurun_out/r02m_cfg.log | cut -c1-200
timeout 900 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r02m_ref.json 2> gpurun_out/r02m_ref.err; echo "ref rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/r02m_ref.json'));print(d['value'], d['ms_per_step'], d['cpu_baseline'])"
