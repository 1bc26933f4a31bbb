set -x
cd $GRAFT_REPO_ROOT
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r02q_bench_8gpu.json 2> gpurun_out/r02q_bench_8gpu.err; echo "bench8 rc=$?"; tail -2 gpurun_out/r02q_bench_8gpu.err
python -c "
import json;d=json.load(open('gpurun_out/r02q_bench_8gpu.json'));print({k:(d[k]['value'] if isinstance(d[k],dict) else d[k]) for k in ('value','e2e','e2e_ee','e2e_reduced','e2e_gathered','strong_scaling','gather_ms')})"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29562 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02q_bench_2gpu.json 2> gpurun_out/r02q_bench_2gpu.err; echo "bench2 rc=$?"
python -c "
import json;d=json.load(open('gpurun_out/r02q_bench_2gpu.json'));print({k:(d[k]['value'] if isinstance(d[k],dict) else d[k]) for k in ('value','e2e','e2e_ee','e2e_reduced','e2e_gathered','strong_scaling','gather_ms')})"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29563 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/r02q_bench_4gpu.json 2> gpurun_out/r02q_bench_4gpu.err; echo "bench4 rc=$?"
python -c "
import json;d=json.load(open('gpurun_out/r02q_bench_4gpu.json'));print({k:(d[k]['value'] if isinstance(d[k],dict) else d[k]) for k in ('value','e2e','e2e_ee','e2e_reduced','e2e_gathered','strong_scaling','gather_ms')})"
