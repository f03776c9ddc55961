mkdir -p gpurun_out
export PATH=/usr/local/cuda/bin:$PATH
timeout 600 python tests/ab_kernel.py product=kmcuda_b200/libKMCUDA.so mma1=variants/mma1/libKMCUDA.so hint500=variants/hint500/libKMCUDA.so b4=variants/b4/libKMCUDA.so > gpurun_out/r2_c3_ab.txt 2>&1
cut -c1-420 gpurun_out/r2_c3_ab.txt
timeout 1200 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/r2_c3_pytest.txt 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|error" gpurun_out/r2_c3_pytest.txt | tail -3; grep -E "^(FAILED|ERROR)" gpurun_out/r2_c3_pytest.txt | head -20
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_assign_kernel -s 3 -c 1 -f -o gpurun_out/r02_tc_assign_v6 \
    python bench.py --steps 2 --warmup 3 --skip-extras > gpurun_out/r02_ncu_full.log 2>&1; echo "ncu full rc=$?"; tail -2 gpurun_out/r02_ncu_full.log
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_c3_bench.json 2> gpurun_out/r2_c3_bench.err; echo "bench rc=$?"; cut -c1-1800 gpurun_out/r2_c3_bench.json
