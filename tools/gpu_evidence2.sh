# round-2 evidence, part 2 (one B200): GPU tests, launch list, bench, reference kernel ncu, secondary configurations
mkdir -p gpurun_out
export PATH=/usr/local/cuda/bin:$PATH
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/r2_pytest.txt 2>&1; echo "pytest rc=$?"; tail -n 4 gpurun_out/r2_pytest.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_final.csv \
    python bench.py --steps 2 --warmup 3 --skip-extras > gpurun_out/r02_ncu_launch.log 2>&1; echo "ncu launches rc=$?"
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/r02_bench_1gpu.json; python -c "
import json; d=json.load(open('gpurun_out/r02_bench_1gpu.json')); print(d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['roofline']['whole_step_frac'], d['clocks'])"
timeout 900 ncu --set full --clock-control none -k regex:kmeans_assign_lloyd -c 1 -f -o gpurun_out/r02_ref_assign_lloyd \
    python bench.py --impl reference --points 1000000 --steps 1 --warmup 1 > gpurun_out/r02_ncu_ref.log 2>&1; echo "ncu ref rc=$?"
timeout 2400 python tests/secondary_configs.py c1 c5 c2 c2c --out gpurun_out/r02_secondary.json > gpurun_out/r02_secondary.log 2>&1; echo "secondary rc=$?"; tail -n 6 gpurun_out/r02_secondary.log | cut -c1-1500
