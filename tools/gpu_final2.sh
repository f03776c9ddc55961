# final validation (one B200): GPU tests, launch list, bench, staged-ingest thread count
mkdir -p gpurun_out
export PATH=/usr/local/cuda/bin:$PATH
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/r2_pytest.txt 2>&1; echo "pytest rc=$?"; tail -n 6 gpurun_out/r2_pytest.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_final.csv \
    python bench.py --steps 2 --warmup 3 --skip-extras > gpurun_out/r02_ncu_launch.log 2>&1; echo "ncu launches rc=$?"
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r02_bench_1gpu.json')); print(d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['roofline']['whole_step_frac'], d['roofline']['traffic'], d['e2e']['value'], d['clocks'])"
for t in 6 12; do KMCUDA_B200_INGEST_THREADS=$t timeout 600 python tests/secondary_configs.py c2c --out gpurun_out/r02_ingest_t$t.json 2>&1 | tail -n 1 | cut -c1-300; done
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 2
