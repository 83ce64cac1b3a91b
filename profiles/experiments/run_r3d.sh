cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for SL in "3,5" "1,3,5" "1,2,3,4,5,6" ""; do
tag=$(echo "sl_$SL" | tr ',' '_')
timeout 300 python bench.py --config demo --steps 3 --no-cpu --chunk-frames 1024 --store-layers "$SL" > gpurun_out/r3d_demo_$tag.json 2> gpurun_out/r3d_demo_$tag.err
python -c "
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]);print(sys.argv[2],d['value'],d['ms_per_step'],d['e2e']['ms_per_step'],d['impl_detail'])" gpurun_out/r3d_demo_$tag.json "$SL"
done
