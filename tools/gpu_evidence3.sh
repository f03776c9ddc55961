# round-2 evidence, part 3 (one B200): Yinyang phase tables on clustered data, C3 as specified
mkdir -p gpurun_out
export PATH=/usr/local/cuda/bin:$PATH
KMCUDA_B200_TIMING=1 timeout 900 python tests/secondary_configs.py c2c --out gpurun_out/r02_secondary_c2c_timing.json > gpurun_out/r02_c2c_timing.log 2> gpurun_out/r02_c2c_timing.err; echo "c2c rc=$?"; grep "timing" gpurun_out/r02_c2c_timing.err | tail -n 60 | cut -c1-200
timeout 3000 python tests/secondary_configs.py c3 --out gpurun_out/r02_secondary.json > gpurun_out/r02_secondary_c3.log 2>&1; echo "c3 rc=$?"; tail -n 3 gpurun_out/r02_secondary_c3.log | cut -c1-3000
