cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
B="python bench.py --no-cpu --no-gpu-torch --no-b128 --no-segments --steps 40"
show() { python -c "
import json,sys;d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]);print(sys.argv[1],d['value'],d['ms_per_step'],d['e2e']['value'])" $1; }
for P in 0 -1 -5; do
SDT_MAIN_PRIORITY=$P timeout 300 $B > gpurun_out/r2m_p$P.json 2> gpurun_out/r2m_p$P.err; show gpurun_out/r2m_p$P.json
done
SDT_MAIN_PRIORITY=0 timeout 300 $B > gpurun_out/r2m_p0b.json 2> gpurun_out/r2m_p0b.err; show gpurun_out/r2m_p0b.json
SDT_MAIN_PRIORITY=-1 timeout 300 $B --batch 128 > gpurun_out/r2m_b128_p1.json 2> gpurun_out/r2m_b128_p1.err; show gpurun_out/r2m_b128_p1.json
SDT_MAIN_PRIORITY=0 timeout 300 $B --batch 128 > gpurun_out/r2m_b128_p0.json 2> gpurun_out/r2m_b128_p0.err; show gpurun_out/r2m_b128_p0.json
python -c "import torch; print(torch.cuda.Stream.priority_range())"
