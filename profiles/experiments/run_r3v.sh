cd $GRAFT_REPO_ROOT
timeout 200 python bench.py --no-cpu --no-gpu-torch --no-b128 --no-segments --steps 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['impl_detail']['comm_mode'], d['impl_detail']['p2p_nvls_multicast'])"
