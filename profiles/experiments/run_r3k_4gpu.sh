cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=4
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
run() { tag=$1; shift; env "$@" timeout 240 $TR --master-port $((29600 + RANDOM % 300)) bench.py --gpus $N --steps 40 --warmup 3 --no-segments > gpurun_out/r3k_${N}gpu_$tag.json 2> gpurun_out/r3k_${N}gpu_$tag.err
python -c "
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]);print(sys.argv[2],d['value'],d['ms_per_step'],d['e2e']['value'],d.get('rank_param_spread'),d['impl_detail']['comm_mode'],d['impl_detail']['graphs_per_step'])
except Exception as e: print(sys.argv[2],'FAILED',e)" gpurun_out/r3k_${N}gpu_$tag.json $tag; grep -i "warn" gpurun_out/r3k_${N}gpu_$tag.err | head -2; }
run p2p_default
run p2p_nomc SDT_P2P_MULTICAST=0
timeout 200 $TR --master-port 29701 bench.py --impl reference --gpus $N --steps 3 --warmup 1 > gpurun_out/r3k_reference_4gpu.json 2> gpurun_out/r3k_reference_4gpu.err; echo "reference rc=$?"; cut -c1-200 gpurun_out/r3k_reference_4gpu.json
