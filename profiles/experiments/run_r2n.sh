cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
B="python bench.py --no-cpu --no-gpu-torch --no-b128 --no-segments --steps 40"
show() { python -c "
import json,sys;d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]);print(sys.argv[1],d['value'],d['ms_per_step'],d['e2e']['value'])" $1; }
for P in -1 -3 0; do
SDT_WG_PRIORITY=$P timeout 300 $B > gpurun_out/r2n_wg$P.json 2> gpurun_out/r2n_wg$P.err; show gpurun_out/r2n_wg$P.json
done
