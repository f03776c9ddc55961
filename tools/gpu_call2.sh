# round-2 evidence run (one B200): ncu captures of the headline kernel and of the reference's assign kernel, launch
# list of one bench step, secondary configurations.  Outputs under gpurun_out/ (copied to profiles/ by hand).
mkdir -p gpurun_out
export PATH=/usr/local/cuda/bin:$PATH
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_assign_kernel -s 3 -c 1 -f -o gpurun_out/r02_tc_assign_v5 \
    python bench.py --steps 2 --warmup 3 --skip-extras > gpurun_out/r02_ncu_full.log 2>&1; echo "ncu full rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 2 --warmup 3 --skip-extras > gpurun_out/r02_ncu_launch.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none -k regex:kmeans_assign_lloyd -c 1 -f -o gpurun_out/r02_ref_assign_lloyd \
    python bench.py --impl reference --points 1000000 --steps 1 --warmup 1 > gpurun_out/r02_ncu_ref.log 2>&1; echo "ncu ref rc=$?"
timeout 1500 python tests/secondary_configs.py c1 c2 c5 --out gpurun_out/r02_secondary.json > gpurun_out/r02_secondary.log 2>&1; echo "secondary rc=$?"; tail -5 gpurun_out/r02_secondary.log | cut -c1-1200
ls -la gpurun_out | tail -12
