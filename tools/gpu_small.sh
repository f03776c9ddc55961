export PATH=/usr/local/cuda/bin:$PATH
for n in 100000 300000 1000000; do for pm in 0 1; do
KMCUDA_B200_PAIR=$pm timeout 200 python tests/ab_kernel.py product=kmcuda_b200/libKMCUDA.so --n $n 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('AB '):
        d=json.loads(l[3:]); print('n=$n pair=$pm step %.4f kernel %.4f'%(d['step_ms'], d['kernel_ms']))"
done; done
