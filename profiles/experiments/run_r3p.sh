cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_inference.py -q -x > gpurun_out/r3p_pytest.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r3p_pytest.log
for SL in "3,5" "1,3,5"; do
tag=$(echo "sl_$SL" | tr ',' '_')
timeout 300 python bench.py --config demo --steps 5 --no-cpu --chunk-frames 1024 --store-layers "$SL" > gpurun_out/r3p_demo_$tag.json 2> gpurun_out/r3p_demo_$tag.err
python -c "
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]);print(sys.argv[2],d['value'],d['ms_per_step'],d['e2e']['ms_per_step'],d['impl_detail'])" gpurun_out/r3p_demo_$tag.json "$SL"
done
timeout 300 python bench.py --config demo --steps 5 --no-cpu --chunk-frames 1024 --no-graph > gpurun_out/r3p_demo_eager.json 2> gpurun_out/r3p_demo_eager.err
python -c "
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]);print('eager',d['value'],d['ms_per_step'],d['e2e']['ms_per_step'],d['impl_detail'])" gpurun_out/r3p_demo_eager.json
