# round-2 closing evidence (one B200): GPU tests, launch list + ncu of the update kernels, bench (both arms),
# secondary configurations, smoke
mkdir -p gpurun_out
export PATH=/usr/local/cuda/bin:$PATH
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/r2_pytest.txt 2>&1; echo "pytest rc=$?"; tail -n 4 gpurun_out/r2_pytest.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_iteration.csv \
    python tools/iteration_probe.py > gpurun_out/r02_ncu_iter.log 2>&1; echo "ncu launches rc=$?"; tail -3 gpurun_out/r02_ncu_iter.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cluster_sums_kernel -s 1 -c 1 -f -o gpurun_out/r02_cluster_sums \
    python tools/iteration_probe.py > gpurun_out/r02_ncu_sums.log 2>&1; echo "ncu sums rc=$?"
timeout 600 ncu --set full --clock-control none -k regex:tc_prep_fused_kernel -s 1 -c 1 -f -o gpurun_out/r02_prep_fused \
    python tools/iteration_probe.py 1000000 > gpurun_out/r02_ncu_prep.log 2>&1; echo "ncu prep rc=$?"
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r02_bench_1gpu.json')); print(d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['roofline']['whole_step_frac'], d['e2e']['value'], d['iteration']['ms'], d['iteration']['phase_ms'], d['clocks'])"
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_1gpu_reference_arm.json 2> gpurun_out/r02_bench_ref.err; echo "ref bench rc=$?"; cut -c1-300 gpurun_out/r02_bench_1gpu_reference_arm.json
timeout 1500 python tests/secondary_configs.py c1 c5 c2 c2c --out gpurun_out/r02_secondary.json > gpurun_out/r02_secondary_final.log 2>&1; echo "secondary rc=$?"; tail -n 5 gpurun_out/r02_secondary_final.log | cut -c1-900
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 2
