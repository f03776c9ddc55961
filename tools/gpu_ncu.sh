# one ncu --set full capture of the headline kernel (output name = $1)
mkdir -p gpurun_out
export PATH=/usr/local/cuda/bin:$PATH
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_assign_kernel -s 3 -c 1 -f -o gpurun_out/$1 \
    python bench.py --steps 2 --warmup 3 --skip-extras > gpurun_out/$1.log 2>&1; echo "ncu full rc=$?"; tail -2 gpurun_out/$1.log
