# fused centroid preparation / balanced member sums: A/B against the chain build at 8M and 1M rows, GPU tests, e2e phase tables
mkdir -p gpurun_out
export PATH=/usr/local/cuda/bin:$PATH
for n in 1000000 8000000; do
timeout 600 python tests/ab_kernel.py product=kmcuda_b200/libKMCUDA.so chain=variants/chain/libKMCUDA.so --n $n 2>&1 | cut -c1-600 | tee -a gpurun_out/r02_update_ab.txt
done
timeout 300 python tools/e2e_probe.py > gpurun_out/r02_update_e2e.txt 2>&1
timeout 300 python tools/e2e_probe.py 100000 pageable >> gpurun_out/r02_update_e2e.txt 2>&1; cat gpurun_out/r02_update_e2e.txt
timeout 1200 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/r02_update_pytest.txt 2>&1; echo "pytest rc=$?"; tail -n 8 gpurun_out/r02_update_pytest.txt
