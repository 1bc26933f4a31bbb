set -x
cd $GRAFT_REPO_ROOT
for c in hubble fgs1 ta_psd grid_sag; do
  timeout 900 python bench.py --config $c > gpurun_out/r02r_cfg_$c.json 2> gpurun_out/r02r_cfg_$c.err; echo "config $c rc=$?"; tail -2 gpurun_out/r02r_cfg_$c.err
  python -c "
import json;d=json.load(open('gpurun_out/r02r_cfg_$c.json'));print('$c', round(d['value']), round(d['e2e']['value']), d['parity'], d['pass_launches_per_psf'])"
done
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/r02r_bench.json 2>/dev/null; python -c "
import json;d=json.load(open('gpurun_out/r02r_bench.json'));print('airs', d['value'], d['roofline']['frac'])"
