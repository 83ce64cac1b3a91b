cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_inference.py -q -x > gpurun_out/r3n_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r3n_pytest.log
