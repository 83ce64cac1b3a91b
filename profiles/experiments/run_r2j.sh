cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2j_pytest.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r2j_pytest.log
timeout 300 python bench.py --no-cpu --no-gpu-torch --no-b128 > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err; echo "bench rc=$?"
python -c "
import json;d=json.loads(open('gpurun_out/r2j_bench.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['e2e']['value']);print(d['roofline']['by_kernel_ms_per_step'])"
timeout 600 python tests/diag_critical_path.py 32 > gpurun_out/r2j_critical_path.txt 2>&1; cat gpurun_out/r2j_critical_path.txt
SDT_DEFER_REDUCE=0 timeout 300 python bench.py --no-cpu --no-gpu-torch --no-b128 --no-segments > gpurun_out/r2j_bench_nodefer.json 2> gpurun_out/r2j_bench_nodefer.err; echo "bench rc=$?"
python -c "
import json;d=json.loads(open('gpurun_out/r2j_bench_nodefer.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['e2e']['value'])"
