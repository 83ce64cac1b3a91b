cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
F="--steps 1 --warmup 1 --no-graph --no-cpu --no-gpu-torch --no-b128 --no-segments"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_conv_tma_kernel --launch-skip 208 -c 40 -f -o gpurun_out/r2p_tma python bench.py $F > gpurun_out/r2p_ncu.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/r2p_ncu.log
