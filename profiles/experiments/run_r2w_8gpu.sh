cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 240 $TR --master-port 29521 bench.py --gpus $N --steps 40 --warmup 3 > gpurun_out/r2w_sdt_bp_${N}gpu.json 2> gpurun_out/r2w_sdt_bp_${N}gpu.err; echo "sdt_bp rc=$?"
SDT_COMM=serial timeout 240 $TR --master-port 29522 bench.py --gpus $N --steps 40 --warmup 3 > gpurun_out/r2w_sdt_bp_${N}gpu_serial.json 2> gpurun_out/r2w_sdt_bp_${N}gpu_serial.err; echo "sdt_bp serial rc=$?"
timeout 240 $TR --master-port 29523 bench.py --gpus $N --config sdt_vae --steps 40 --warmup 3 > gpurun_out/r2w_sdt_vae_${N}gpu.json 2> gpurun_out/r2w_sdt_vae_${N}gpu.err; echo "sdt_vae rc=$?"
timeout 240 $TR --master-port 29524 bench.py --gpus $N --config pose2pose --steps 40 --warmup 3 > gpurun_out/r2w_pose2pose_${N}gpu.json 2> gpurun_out/r2w_pose2pose_${N}gpu.err; echo "pose2pose rc=$?"
python -c "
import json,sys
N=sys.argv[1]
for f in ('sdt_bp_%sgpu'%N,'sdt_bp_%sgpu_serial'%N,'sdt_vae_%sgpu'%N,'pose2pose_%sgpu'%N):
    try:
        d=json.loads(open('gpurun_out/r2w_%s.json'%f).read().strip().splitlines()[-1]);print(f,d['value'],d['ms_per_step'],d['e2e']['value'],d.get('rank_param_spread'),d['impl_detail']['comm_mode'],d['config']['global_batch'])
    except Exception as e: print(f,'FAILED',e)
" $N
