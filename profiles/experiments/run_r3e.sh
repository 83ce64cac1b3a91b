cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_inference.py tests/test_gpu_step.py tests/test_gpu_pipelines.py tests/test_gpu_eval.py -q -x > gpurun_out/r3e_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r3e_pytest.log
