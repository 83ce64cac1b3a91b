cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python tests/diag_conv1d_timeline.py 32 > gpurun_out/r2y_conv1d_timeline.txt 2>&1; cut -c1-230 gpurun_out/r2y_conv1d_timeline.txt
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2y_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r2y_pytest.log
B="python bench.py --no-cpu --no-gpu-torch --no-b128 --steps 40"
show() { python -c "
import json,sys;d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]);print(sys.argv[1],d['value'],d['ms_per_step'],d['e2e']['value'],d['gpu_launches'], d.get('north_star',{}).get('decoder',{}).get('ms'))" $1; }
timeout 300 $B > gpurun_out/r2y_a.json 2> gpurun_out/r2y_a.err; show gpurun_out/r2y_a.json
timeout 300 $B > gpurun_out/r2y_b.json 2> gpurun_out/r2y_b.err; show gpurun_out/r2y_b.json
