# Yinyang changes: GPU tests + phase tables on the clustered 8M data + C2
mkdir -p gpurun_out
export PATH=/usr/local/cuda/bin:$PATH
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/r2_pytest.txt 2>&1; echo "pytest rc=$?"; tail -n 6 gpurun_out/r2_pytest.txt
KMCUDA_B200_TIMING=1 KMCUDA_B200_YY_ADAPTIVE=0 timeout 900 python tests/secondary_configs.py c2c --out gpurun_out/r02_secondary_c2c_timing.json > gpurun_out/r02_c2c_timing.log 2> gpurun_out/r02_c2c_timing.err; echo "c2c (yinyang forced) rc=$?"; grep "timing" gpurun_out/r02_c2c_timing.err | tail -n 12 | cut -c1-200; tail -n 1 gpurun_out/r02_c2c_timing.log | cut -c1-400
timeout 1500 python tests/secondary_configs.py c2 c2c --out gpurun_out/r02_secondary.json > gpurun_out/r02_secondary_yy.log 2>&1; echo "secondary rc=$?"; tail -n 3 gpurun_out/r02_secondary_yy.log | cut -c1-900
