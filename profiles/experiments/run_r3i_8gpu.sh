cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
run() { tag=$1; shift; timeout 240 $TR --master-port $((29600 + RANDOM % 300)) bench.py --gpus $N --steps 40 --warmup 3 "$@" > gpurun_out/r3i_${N}gpu_$tag.json 2> gpurun_out/r3i_${N}gpu_$tag.err
python -c "
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]);print(sys.argv[2],d['value'],d['ms_per_step'],d['e2e']['value'],d.get('rank_param_spread'),d['impl_detail']['comm_mode'],d['impl_detail']['graphs_per_step'],d['config']['global_batch'])
except Exception as e: print(sys.argv[2],'FAILED',e)" gpurun_out/r3i_${N}gpu_$tag.json $tag; grep -i "warn" gpurun_out/r3i_${N}gpu_$tag.err | head -2; }
run sdt_bp
run sdt_vae --config sdt_vae
run pose2pose --config pose2pose
run demo --config demo --steps 3
