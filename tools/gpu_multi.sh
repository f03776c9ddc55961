# multi-GPU run (gpurun --gpus N): both bench arms under torchrun + the >= 2-GPU tests.  usage: gpu_multi.sh N
N=$1
mkdir -p gpurun_out
export PATH=/usr/local/cuda/bin:$PATH
export NCCL_DEBUG=INFO
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02_bench_${N}gpu.json 2> gpurun_out/r02_bench_${N}gpu.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/r02_bench_${N}gpu.json'))
print({k:d[k] for k in ('value','n_gpus','ms_per_step','scaling')}, d['iteration']['phase_ms'], d['iteration']['value'], d['e2e']['value'], d['roofline']['frac'], d['clocks'])
PY
grep -c "NCCL INFO" gpurun_out/r02_bench_${N}gpu.err; grep -m3 "NVLS\|comm 0x.*rank.*nranks" gpurun_out/r02_bench_${N}gpu.err | cut -c1-200
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/r02_bench_${N}gpu_reference_arm.json 2> gpurun_out/r02_bench_${N}gpu_ref.err; echo "ref bench rc=$?"; cut -c1-1200 gpurun_out/r02_bench_${N}gpu_reference_arm.json
unset NCCL_DEBUG
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -k "multi_gpu or pair" > gpurun_out/r2_pytest_${N}gpu.txt 2>&1; echo "pytest rc=$?"; tail -n 4 gpurun_out/r2_pytest_${N}gpu.txt
