cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2i_pytest.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/r2i_pytest.log
F="--steps 2 --warmup 1 --no-graph --no-cpu --no-gpu-torch --no-b128 --no-segments"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2i_launches.csv python bench.py $F > gpurun_out/r2i_ncu_bench.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_conv_ytap --launch-skip 70 -c 14 -f -o gpurun_out/r2i_ytap python bench.py $F > gpurun_out/r2i_ncu_full.log 2>&1; echo "ncu full rc=$?"
timeout 300 python bench.py --config sdt_vae --no-cpu > gpurun_out/r2i_sdt_vae.json 2> gpurun_out/r2i_sdt_vae.err; echo "sdt_vae rc=$?"
timeout 300 python bench.py --config pose2pose --no-cpu > gpurun_out/r2i_pose2pose.json 2> gpurun_out/r2i_pose2pose.err; echo "pose2pose rc=$?"
timeout 400 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2i_reference.json 2> gpurun_out/r2i_reference.err; echo "reference rc=$?"
cut -c1-400 gpurun_out/r2i_sdt_vae.json gpurun_out/r2i_pose2pose.json gpurun_out/r2i_reference.json
