# 2 GPUs: peer-memory exchange test (torchrun inside pytest) + the single-process multi-GPU test + bench under torchrun
N=${1:-2}
mkdir -p gpurun_out
export PATH=/usr/local/cuda/bin:$PATH
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -k "multi_gpu or member_sums" > gpurun_out/r02_exchange_pytest_${N}gpu.txt 2>&1; echo "pytest rc=$?"; tail -n 30 gpurun_out/r02_exchange_pytest_${N}gpu.txt
export NCCL_DEBUG=INFO
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02_exchange_bench_${N}gpu.json 2> gpurun_out/r02_exchange_bench_${N}gpu.err; echo "bench rc=$?"
tail -5 gpurun_out/r02_exchange_bench_${N}gpu.err | cut -c1-300
python - <<PY
import json
d=json.load(open('gpurun_out/r02_exchange_bench_${N}gpu.json'))
print({k:d[k] for k in ('value','n_gpus','ms_per_step','scaling')}, json.dumps(d['iteration'])[:1500], d['e2e']['value'], d['roofline']['frac'], d['clocks'])
PY
