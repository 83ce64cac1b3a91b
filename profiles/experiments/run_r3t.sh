cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/r3t_bench.json 2> gpurun_out/r3t_bench.err; echo "bench rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/r3t_bench.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['e2e']['value'],d['clocks'],d['roofline']['achieved'],d['roofline']['frac'],d['b128']['clips_per_s'],d['b128']['ms_per_step'],d['north_star']['decoder']['ms'],d['gpu_torch_baseline']['value'])"
