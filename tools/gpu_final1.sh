# final checks (one B200): GPU tests, secondary configurations (ingest change), C3 as specified, launch list, bench
mkdir -p gpurun_out
export PATH=/usr/local/cuda/bin:$PATH
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/r2_pytest.txt 2>&1; echo "pytest rc=$?"; tail -n 6 gpurun_out/r2_pytest.txt
timeout 1500 python tests/secondary_configs.py c1 c5 c2 c2c --out gpurun_out/r02_secondary.json > gpurun_out/r02_secondary_final.log 2>&1; echo "secondary rc=$?"; tail -n 4 gpurun_out/r02_secondary_final.log | cut -c1-700
timeout 2400 python tests/secondary_configs.py c3 --out gpurun_out/r02_secondary.json > gpurun_out/r02_secondary_c3.log 2>&1; echo "c3 rc=$?"; tail -n 1 gpurun_out/r02_secondary_c3.log | cut -c1-3500
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r02_bench_1gpu.json')); print(d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['roofline']['whole_step_frac'], d['roofline']['traffic'], d['e2e']['value'])"
