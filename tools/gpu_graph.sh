mkdir -p gpurun_out
export PATH=/usr/local/cuda/bin:$PATH
timeout 600 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -k "graph" 2>&1 | tail -n 3
for g in 0 1; do for p in 8000000 1000000; do
KMCUDA_B200_GRAPH=$g timeout 300 python bench.py --steps 30 --warmup 5 --points $p --skip-extras 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('graph=$g points=$p', d['ms_per_step'], d['kernel_ms'])"
done; done
