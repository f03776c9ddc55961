mkdir -p gpurun_out
export PATH=/usr/local/cuda/bin:$PATH
# knock-out timing experiments (results of the ko builds are garbage by construction; only kernel_ms matters)
timeout 900 python tests/ab_kernel.py product=kmcuda_b200/libKMCUDA.so ko1=variants/ko1/libKMCUDA.so ko2=variants/ko2/libKMCUDA.so ko3=variants/ko3/libKMCUDA.so ko5=variants/ko5/libKMCUDA.so ko6=variants/ko6/libKMCUDA.so --n 4000000 > gpurun_out/r2_c2_ko.txt 2>&1
cut -c1-420 gpurun_out/r2_c2_ko.txt
KMCUDA_B200_DEBUG=1 timeout 200 python tests/ab_kernel.py product=kmcuda_b200/libKMCUDA.so --env "KMCUDA_B200_GRAPH=1" --n 2000000 > gpurun_out/r2_c2_graph.txt 2>&1; tail -c 1500 gpurun_out/r2_c2_graph.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_assign_kernel -s 3 -c 1 -f -o gpurun_out/r02_tc_assign_v5 \
    python bench.py --steps 2 --warmup 3 --skip-extras > gpurun_out/r02_ncu_full.log 2>&1; echo "ncu full rc=$?"; tail -3 gpurun_out/r02_ncu_full.log
timeout 600 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -k "whole_run or offset_data or wide_shapes or golden or yinyang or headline_8m" > gpurun_out/r2_c2_pytest.txt 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|error" gpurun_out/r2_c2_pytest.txt | tail -3; grep -E "^(FAILED|ERROR)" gpurun_out/r2_c2_pytest.txt | head
