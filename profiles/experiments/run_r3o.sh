cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r3o_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r3o_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/r3o_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r3o_smoke.log
timeout 300 python bench.py --steps 20 --warmup 3 --no-b128 --no-gpu-torch > gpurun_out/r3o_bench.json 2> gpurun_out/r3o_bench.err; echo "bench rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/r3o_bench.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['e2e']['value'],d['roofline']['frac'],d['cpu_baseline']['value'])"
