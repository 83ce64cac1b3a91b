cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python tests/diag_conv1d_timeline.py 32 > gpurun_out/r3q_conv1d_timeline.txt 2>&1; cut -c1-260 gpurun_out/r3q_conv1d_timeline.txt
