# round-2 evidence, part 1 (one B200): launch list of bench steps, full default bench run, reference-arm bench, SASS summary inputs
mkdir -p gpurun_out
export PATH=/usr/local/cuda/bin:$PATH
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_final.csv \
    python bench.py --steps 2 --warmup 3 --skip-extras > gpurun_out/r02_ncu_launch.log 2>&1; echo "ncu launches rc=$?"
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err; echo "bench rc=$?"; cut -c1-3000 gpurun_out/r02_bench_1gpu.json
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_1gpu_reference_arm.json 2> gpurun_out/r02_bench_ref.err; echo "ref bench rc=$?"; cut -c1-1500 gpurun_out/r02_bench_1gpu_reference_arm.json
