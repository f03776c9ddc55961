mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > gpurun_out/r2_c1_smi.txt 2>&1
timeout 600 python tools/ab_kernel.py product=kmcuda_b200/libKMCUDA.so role0=variants/role0/libKMCUDA.so b3=variants/b3/libKMCUDA.so nohint=variants/nohint/libKMCUDA.so > gpurun_out/r2_c1_ab.txt 2>&1
timeout 200 python tools/ab_kernel.py product=kmcuda_b200/libKMCUDA.so --env "KMCUDA_B200_NO_CENTER=1" >> gpurun_out/r2_c1_ab.txt 2>&1
cat gpurun_out/r2_c1_ab.txt
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2_c1_pytest.txt 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r2_c1_pytest.txt
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_c1_bench.json 2> gpurun_out/r2_c1_bench.err; echo "bench rc=$?"; cat gpurun_out/r2_c1_bench.json | cut -c1-1500
