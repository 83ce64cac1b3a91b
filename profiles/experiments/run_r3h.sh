cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
run() { tag=$1; shift; env "$@" timeout 240 $TR --master-port $((29600 + RANDOM % 300)) bench.py --gpus $N --steps 40 --warmup 3 --no-segments > gpurun_out/r3h_${N}gpu_$tag.json 2> gpurun_out/r3h_${N}gpu_$tag.err
python -c "
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]);print(sys.argv[2],d['value'],d['ms_per_step'],d['e2e']['value'],d.get('rank_param_spread'),d['impl_detail']['comm_mode'],d['impl_detail']['graphs_per_step'])
except Exception as e: print(sys.argv[2],'FAILED',e)" gpurun_out/r3h_${N}gpu_$tag.json $tag; grep -i "warn\|error" gpurun_out/r3h_${N}gpu_$tag.err | head -3; }
run p2p_1graph_nomc SDT_COMM=p2p SDT_P2P_MULTICAST=0
run p2p_2graphs_nomc SDT_COMM=p2p SDT_P2P_MULTICAST=0 SDT_P2P_2GRAPHS=1
run p2p_1graph_mc SDT_COMM=p2p
timeout 500 $TR --master-port 29511 tests/multi_gpu_check.py --out gpurun_out/r3h_multi_gpu_check.json > gpurun_out/r3h_check.log 2>&1; echo "check rc=$?"
grep -E '"ok"|p2p/graphs|p2p_vs' gpurun_out/r3h_check.log
