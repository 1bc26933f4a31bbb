set -x
cd $GRAFT_REPO_ROOT
CMD="python bench.py --n-wl 16 --steps 1 --warmup 1 --no-cpu --slots 1"
timeout 600 $CMD > gpurun_out/r02_plain.json 2> gpurun_out/r02_plain.err; echo "plain rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches.csv $CMD > gpurun_out/r02_ncu1.log 2>&1; echo "ncu1 rc=$?"
timeout 1200 ncu --set full --import-source on --clock-control none -k regex:pass_kernel -s 14 -c 14 -f -o gpurun_out/r02_prof $CMD > gpurun_out/r02_ncu2.log 2>&1; echo "ncu2 rc=$?"
ls -la gpurun_out | tail -8
