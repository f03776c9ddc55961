# A/B of the update-path / preparation micro-changes against the previous commit's build + the tests they touch
mkdir -p gpurun_out
export PATH=/usr/local/cuda/bin:$PATH
for n in 1000000 8000000; do
for lib in kmcuda_b200/libKMCUDA.so variants/before/libKMCUDA.so; do
KMCUDA_B200_LIB=$PWD/$lib timeout 200 python tools/iteration_probe.py $n 2>&1 | grep ITER | tee -a gpurun_out/r3h_iter.txt
done; done
timeout 600 python tests/ab_kernel.py product=kmcuda_b200/libKMCUDA.so before=variants/before/libKMCUDA.so --n 1000000 2>&1 | cut -c1-420 | tee -a gpurun_out/r3h_ab.txt
timeout 1200 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/r3h_pytest.txt 2>&1; echo "pytest rc=$?"; tail -n 4 gpurun_out/r3h_pytest.txt
