cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2k_pytest.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r2k_pytest.log
B="python bench.py --no-cpu --no-gpu-torch --no-b128 --no-segments --steps 40"
show() { python -c "
import json,sys;d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]);print(sys.argv[1],d['value'],d['ms_per_step'],d['e2e']['value']);print(d['roofline']['by_kernel_ms_per_step'])" $1; }
timeout 300 $B > gpurun_out/r2k_a.json 2> gpurun_out/r2k_a.err; show gpurun_out/r2k_a.json
SDT_YTAP_CLASS_MAJOR=1 timeout 300 $B > gpurun_out/r2k_b.json 2> gpurun_out/r2k_b.err; show gpurun_out/r2k_b.json
SDT_NORM_BWD_L2_MB=0 timeout 300 $B > gpurun_out/r2k_c.json 2> gpurun_out/r2k_c.err; show gpurun_out/r2k_c.json
SDT_NORM_BWD_L2_MB=40 timeout 300 $B > gpurun_out/r2k_d.json 2> gpurun_out/r2k_d.err; show gpurun_out/r2k_d.json
SDT_NORM_BWD_L2_MB=100 timeout 300 $B > gpurun_out/r2k_e.json 2> gpurun_out/r2k_e.err; show gpurun_out/r2k_e.json
