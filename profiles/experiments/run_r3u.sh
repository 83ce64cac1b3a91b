cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python tests/diag_conv1d_timeline.py 32 2>&1 | cut -c1-250 | grep "L 64 k 3 s 1:"
timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_rownorm_fused.py -q -x > gpurun_out/r3u_pytest.log 2>&1; echo "pytest rc=$?"
tail -2 gpurun_out/r3u_pytest.log
B="python bench.py --no-cpu --no-gpu-torch --no-b128 --steps 40"
show() { python -c "
import json,sys;d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]);print(sys.argv[1],d['value'],d['ms_per_step'],d['e2e']['value'],d['north_star']['decoder']['ms'])" $1; }
timeout 300 $B > gpurun_out/r3u_a.json 2> gpurun_out/r3u_a.err; show gpurun_out/r3u_a.json
timeout 300 $B > gpurun_out/r3u_b.json 2> gpurun_out/r3u_b.err; show gpurun_out/r3u_b.json
