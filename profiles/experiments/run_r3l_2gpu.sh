cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
t0=$(date +%s)
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29811 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r3l_driver_like_2gpu.json 2> gpurun_out/r3l_driver_like_2gpu.err; echo "rc=$? in $(( $(date +%s) - t0 )) s"
wc -l gpurun_out/r3l_driver_like_2gpu.json; cut -c1-300 gpurun_out/r3l_driver_like_2gpu.json
python -c "
import json
d=json.loads(open('gpurun_out/r3l_driver_like_2gpu.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['e2e']['value'],d.get('rank_param_spread'),d['impl_detail'],d['roofline']['frac'],list(d.keys()))"
