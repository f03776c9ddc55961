# multi-GPU validation (gpurun --gpus 2): the >= 2 GPU parity test, the strong-scaling bench under torchrun, the
# reference arm with device mask 0x3
mkdir -p gpurun_out
N=${1:-2}
nvidia-smi -L > gpurun_out/r2_mg_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -k "multi_gpu" > gpurun_out/r2_mg_pytest.txt 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r2_mg_pytest.txt
NCCL_DEBUG=INFO timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_mg_bench_$N.json 2> gpurun_out/r2_mg_bench_$N.err; echo "bench rc=$?"; cut -c1-2600 gpurun_out/r2_mg_bench_$N.json; grep -E "NCCL INFO.*(nranks|Connected|NVLS)" gpurun_out/r2_mg_bench_$N.err | head -5; tail -3 gpurun_out/r2_mg_bench_$N.err
timeout 600 python bench.py --impl reference --gpus $N --steps 3 --warmup 1 > gpurun_out/r2_mg_ref_$N.json 2> gpurun_out/r2_mg_ref_$N.err; echo "ref rc=$?"; cut -c1-1200 gpurun_out/r2_mg_ref_$N.json; tail -3 gpurun_out/r2_mg_ref_$N.err
