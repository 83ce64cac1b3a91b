cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r3j_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r3j_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r3j_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r3j_smoke.log
timeout 600 python bench.py > gpurun_out/r3j_bench.json 2> gpurun_out/r3j_bench.err; echo "bench rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/r3j_bench.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['e2e']['value'],d['clocks'],d['roofline']['achieved'],d['roofline']['frac'],d['b128']['clips_per_s'],d['north_star']['mel_enc'].get('frac_of_tf32_sustained'))"
