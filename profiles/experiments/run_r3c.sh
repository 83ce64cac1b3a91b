cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r3c_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r3c_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r3c_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r3c_smoke.log
timeout 600 python bench.py > gpurun_out/r3c_bench.json 2> gpurun_out/r3c_bench.err; echo "bench rc=$?"
F="--steps 2 --warmup 1 --no-graph --no-cpu --no-gpu-torch --no-b128 --no-segments"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r3c_launches.csv python bench.py $F > gpurun_out/r3c_ncu_bench.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_conv_ytap --launch-skip 70 -c 14 -f -o gpurun_out/r3c_ytap python bench.py $F > gpurun_out/r3c_ncu_full.log 2>&1; echo "ncu full rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/r3c_bench.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['e2e']['value'],d['clocks'],d['roofline']['achieved'],d['roofline']['frac'],d['b128']['clips_per_s'],d['gpu_torch_baseline']['value'],d['cpu_baseline']['value'])"
