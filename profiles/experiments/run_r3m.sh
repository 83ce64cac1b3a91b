cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
B="python bench.py --no-cpu --no-gpu-torch --no-b128 --no-segments --steps 40"
show() { python -c "
import json,sys;d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]);print(sys.argv[1],d['value'],d['ms_per_step'],d['e2e']['value'],d['last_losses']['G_loss'])" $1; }
SDT_WGRAD_AFTER_DGRAD=1 timeout 300 $B > gpurun_out/r3m_after.json 2> gpurun_out/r3m_after.err; show gpurun_out/r3m_after.json
timeout 300 $B > gpurun_out/r3m_before.json 2> gpurun_out/r3m_before.err; show gpurun_out/r3m_before.json
SDT_WGRAD_AFTER_DGRAD=1 timeout 300 $B > gpurun_out/r3m_after2.json 2> gpurun_out/r3m_after2.err; show gpurun_out/r3m_after2.json
SDT_WGRAD_AFTER_DGRAD=1 timeout 300 $B --batch 128 > gpurun_out/r3m_after_b128.json 2> gpurun_out/r3m_after_b128.err; show gpurun_out/r3m_after_b128.json
timeout 300 $B --batch 128 > gpurun_out/r3m_before_b128.json 2> gpurun_out/r3m_before_b128.err; show gpurun_out/r3m_before_b128.json
