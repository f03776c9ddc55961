# round-2 status run (one B200): GPU tests, default bench, launch list of one bench step
mkdir -p gpurun_out
export PATH=/usr/local/cuda/bin:$PATH
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/r2_pytest.txt 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|error" gpurun_out/r2_pytest.txt | tail -3; grep -E "^(FAILED|ERROR)" gpurun_out/r2_pytest.txt | head -20
timeout 500 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; echo "bench rc=$?"; cut -c1-2500 gpurun_out/r2_bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 2 --warmup 3 --skip-extras > gpurun_out/r02_ncu_launch.log 2>&1; echo "ncu launches rc=$?"
