cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2r_pytest.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r2r_pytest.log
timeout 300 python tests/diag_conv_timeline.py 32 3 > gpurun_out/r2r_conv_layers.txt 2>&1; cat gpurun_out/r2r_conv_layers.txt
B="python bench.py --no-cpu --no-gpu-torch --no-b128 --no-segments --steps 40"
show() { python -c "
import json,sys;d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]);print(sys.argv[1],d['value'],d['ms_per_step'],d['e2e']['value']);print(d['roofline']['achieved'], d['roofline']['by_kernel_ms_per_step'])" $1; }
timeout 300 $B > gpurun_out/r2r_a.json 2> gpurun_out/r2r_a.err; show gpurun_out/r2r_a.json
timeout 300 $B > gpurun_out/r2r_b.json 2> gpurun_out/r2r_b.err; show gpurun_out/r2r_b.json
timeout 300 $B --batch 128 > gpurun_out/r2r_b128.json 2> gpurun_out/r2r_b128.err; show gpurun_out/r2r_b128.json
