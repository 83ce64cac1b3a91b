cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 500 $TR --master-port 29511 tests/multi_gpu_check.py --out gpurun_out/r3f_multi_gpu_check.json > gpurun_out/r3f_check.log 2>&1; echo "check rc=$?"
grep -E "p2p|ok|Error|error|Traceback|warn" gpurun_out/r3f_check.log | head -40
run() { tag=$1; shift; env "$@" timeout 240 $TR --master-port $((29600 + RANDOM % 300)) bench.py --gpus $N --steps 40 --warmup 3 --no-segments > gpurun_out/r3f_${N}gpu_$tag.json 2> gpurun_out/r3f_${N}gpu_$tag.err
python -c "
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]);print(sys.argv[2],d['value'],d['ms_per_step'],d['e2e']['value'],d.get('rank_param_spread'),d['impl_detail']['comm_mode'])
except Exception as e: print(sys.argv[2],'FAILED',e)" gpurun_out/r3f_${N}gpu_$tag.json $tag; tail -2 gpurun_out/r3f_${N}gpu_$tag.err; }
run p2p SDT_COMM=p2p
run p2p_nomc SDT_COMM=p2p SDT_P2P_MULTICAST=0
run serial SDT_COMM=serial
