mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > gpurun_out/r2_c1_smi.txt 2>&1
timeout 600 python tests/ab_kernel.py product=kmcuda_b200/libKMCUDA.so role0=variants/role0/libKMCUDA.so b3=variants/b3/libKMCUDA.so nohint=variants/nohint/libKMCUDA.so hint500=variants/hint500/libKMCUDA.so > gpurun_out/r2_c1_ab.txt 2>&1
timeout 300 python tests/ab_kernel.py product=kmcuda_b200/libKMCUDA.so --env "KMCUDA_B200_NO_CENTER=1;KMCUDA_B200_GRAPH=1" >> gpurun_out/r2_c1_ab.txt 2>&1
cat gpurun_out/r2_c1_ab.txt
timeout 1200 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/r2_c1_pytest.txt 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|error" gpurun_out/r2_c1_pytest.txt | tail -5; grep -E "^(FAILED|ERROR)" gpurun_out/r2_c1_pytest.txt | head -40
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_c1_bench.json 2> gpurun_out/r2_c1_bench.err; echo "bench rc=$?"; cut -c1-2500 gpurun_out/r2_c1_bench.json; tail -5 gpurun_out/r2_c1_bench.err
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_c1_smoke.txt 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r2_c1_smoke.txt
