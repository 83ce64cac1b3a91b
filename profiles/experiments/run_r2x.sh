cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python tests/diag_conv1d_timeline.py 32 > gpurun_out/r2x_conv1d_timeline.txt 2>&1; cat gpurun_out/r2x_conv1d_timeline.txt
