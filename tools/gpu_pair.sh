# first runs of the CTA-pair kernel: short, under timeouts
mkdir -p gpurun_out
export PATH=/usr/local/cuda/bin:$PATH
echo "== pair off"; KMCUDA_B200_PAIR=0 timeout 200 python tests/ab_kernel.py product=kmcuda_b200/libKMCUDA.so --n 2000000 2>&1 | cut -c1-400
echo "== pair on"; timeout 200 python tests/ab_kernel.py product=kmcuda_b200/libKMCUDA.so --n 2000000 2>&1 | cut -c1-400
echo "== pair on, 8M"; timeout 200 python tests/ab_kernel.py product=kmcuda_b200/libKMCUDA.so 2>&1 | cut -c1-400
