cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
echo "== BN=64"; timeout 300 python tests/diag_conv1d_timeline.py 32 2>&1 | cut -c1-250 | grep "L 64 k 3"
echo "== BN=128"; SDT_TMA_BN128=1 timeout 300 python tests/diag_conv1d_timeline.py 32 2>&1 | cut -c1-250 | grep "L 64 k 3"
