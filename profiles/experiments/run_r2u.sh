cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_rownorm_fused.py -x -q > gpurun_out/r2u_fused_test.log 2>&1; echo "fused test rc=$?"
tail -15 gpurun_out/r2u_fused_test.log
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2u_pytest.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r2u_pytest.log
B="python bench.py --no-cpu --no-gpu-torch --no-b128 --steps 40"
show() { python -c "
import json,sys;d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]);print(sys.argv[1],d['value'],d['ms_per_step'],d['e2e']['value'],d['gpu_launches'], d.get('north_star',{}).get('decoder'))" $1; }
timeout 300 $B > gpurun_out/r2u_a.json 2> gpurun_out/r2u_a.err; show gpurun_out/r2u_a.json
SDT_FUSE_ROWNORM=0 timeout 300 $B > gpurun_out/r2u_b.json 2> gpurun_out/r2u_b.err; show gpurun_out/r2u_b.json
timeout 300 $B > gpurun_out/r2u_c.json 2> gpurun_out/r2u_c.err; show gpurun_out/r2u_c.json
