cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29511 tests/multi_gpu_check.py --out gpurun_out/r2s_multi_gpu_check.json > gpurun_out/r2s_check.log 2>&1; echo "check rc=$?"
tail -3 gpurun_out/r2s_check.log
timeout 300 $TR --master-port 29512 bench.py --gpus 2 --steps 40 --warmup 3 > gpurun_out/r2s_bench_2gpu.json 2> gpurun_out/r2s_bench_2gpu.err; echo "bench rc=$?"
SDT_COMM=serial timeout 300 $TR --master-port 29513 bench.py --gpus 2 --steps 40 --warmup 3 > gpurun_out/r2s_bench_2gpu_serial.json 2> gpurun_out/r2s_bench_2gpu_serial.err; echo "bench serial rc=$?"
python -c "
import json
for f in ('r2s_bench_2gpu','r2s_bench_2gpu_serial'):
    d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1]);print(f,d['value'],d['ms_per_step'],d['e2e']['value'],d.get('rank_param_spread'),d['impl_detail']['comm_mode'])"
