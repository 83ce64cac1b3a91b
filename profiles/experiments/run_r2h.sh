cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/r2h_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r2h_pytest.log
timeout 500 python bench.py > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err; echo "bench rc=$?"
tail -5 gpurun_out/r2h_bench.err
timeout 120 python profiles/measure_tf32_peak.py gpurun_out/r2_tf32_peak.json > gpurun_out/r2h_tf32.log 2>&1; echo "tf32 rc=$?"
timeout 300 python bench.py --config demo --steps 3 --no-cpu > gpurun_out/r2h_demo.json 2> gpurun_out/r2h_demo.err; echo "demo rc=$?"
timeout 300 python bench.py --config demo --steps 3 --no-cpu --chunk-frames 1024 > gpurun_out/r2h_demo_chunked.json 2> gpurun_out/r2h_demo_chunked.err; echo "demo chunked rc=$?"
cat gpurun_out/r2h_demo.json gpurun_out/r2h_demo_chunked.json | cut -c1-600
