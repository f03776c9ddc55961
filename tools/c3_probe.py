"""C3-shaped Lloyd pass probe (run on the B200): 4M x 480 unit vectors (fp16-representable values) @ 40 000 centroids,
angular metric, shard-level API; prints device times of the pass and of the tensor-core kernel."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import kmcuda_b200
from kmcuda_b200.shard import Shard
n, D, K = int(os.environ.get("C3_N", "4000000")), 480, 40000
g = torch.Generator(device="cuda").manual_seed(777)
X = torch.empty((n, D), device="cuda", dtype=torch.float32)
for i in range(0, n, 500000):
    b = torch.randn((min(500000, n - i), D), generator=g, device="cuda")
    b /= b.norm(dim=1, keepdim=True)
    X[i:i + len(b)] = b.half().float()
C = X[torch.randperm(n, generator=g, device="cuda")[:K]].clone()
C += 0.02 * torch.randn(C.shape, generator=g, device="cuda")
C = (C / C.norm(dim=1, keepdim=True)).half().float().contiguous()
sh = Shard(n, D, K, "cos")
a = torch.full((n,), -1, dtype=torch.int32, device="cuda"); prev = a.clone(); ch = torch.zeros(1, dtype=torch.int32, device="cuda")
for it in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); sh.assign(X, C, a, prev, ch); e1.record(); torch.cuda.synchronize()
    kt = sh.kernel_times(1)
    print("pass %d: %.1f ms, tc kernel %.1f ms, info %s err %x" % (it, e0.elapsed_time(e1), kt[0], sh.last_pass_info(), sh.last_error()), flush=True)
