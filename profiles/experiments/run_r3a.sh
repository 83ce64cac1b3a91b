cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
SDT_YTAP_DYNAMIC=1 timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_step.py -q -x > gpurun_out/r3a_pytest_dyn.log 2>&1; echo "pytest dynamic rc=$?"
tail -3 gpurun_out/r3a_pytest_dyn.log
B="python bench.py --no-cpu --no-gpu-torch --no-b128 --no-segments --steps 40"
show() { python -c "
import json,sys;d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]);print(sys.argv[1],d['value'],d['ms_per_step'],d['e2e']['value'],d['roofline']['achieved'],d['last_losses']['G_loss'])" $1; }
SDT_YTAP_DYNAMIC=1 timeout 300 $B > gpurun_out/r3a_dyn.json 2> gpurun_out/r3a_dyn.err; show gpurun_out/r3a_dyn.json
timeout 300 $B > gpurun_out/r3a_static.json 2> gpurun_out/r3a_static.err; show gpurun_out/r3a_static.json
SDT_YTAP_DYNAMIC=1 timeout 300 $B > gpurun_out/r3a_dyn2.json 2> gpurun_out/r3a_dyn2.err; show gpurun_out/r3a_dyn2.json
SDT_YTAP_DYNAMIC=1 timeout 300 $B --batch 128 > gpurun_out/r3a_dyn_b128.json 2> gpurun_out/r3a_dyn_b128.err; show gpurun_out/r3a_dyn_b128.json
timeout 300 $B --batch 128 > gpurun_out/r3a_static_b128.json 2> gpurun_out/r3a_static_b128.err; show gpurun_out/r3a_static_b128.json
