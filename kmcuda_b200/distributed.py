"""One-process-per-GPU Lloyd iterations over range-partitioned samples.

The reference drives all GPUs from one process and exchanges assignments / centroid slices with
peer copies after every kernel phase (reference src/kmeans.cu:980-990,1014-1024).  Here every rank
owns one shard of the samples; the assignment step needs no communication at all, and the centroid
update needs exactly one all-reduce of the [K][D] fp32 partial sums, the [K] integer counts and the
reassignment counter (counts are never summed as floats: a cluster may hold more than 2^24 samples).

`backend` is anything with the three shard steps of include/kmcuda_b200.h -- in production
`kmcuda_b200.shard.Shard` (CUDA, NCCL), in the CPU tests a numpy stand-in over gloo.
"""
import torch
import torch.distributed as dist


def _world():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def allreduce_update(sums, counts):
    """In-place sum over ranks of the per-cluster partial sums (fp32) and counts (integer)."""
    if _world() > 1:
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
        dist.all_reduce(counts, op=dist.ReduceOp.SUM)


def allreduce_scalar(value, device):
    t = torch.tensor([int(value)], dtype=torch.int64, device=device)
    if _world() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return int(t.item())


def sharded_lloyd(backend, X, C, total_samples, tolerance=0.01, max_iter=0, log=None, exchange=None):
    """Lloyd's algorithm on this rank's shard X ([n_local][D]) starting from centroids C ([K][D],
    identical on every rank; updated in place).  Stop rule of the reference (kmeans.cu:707):
    reassignments <= tolerance * total_samples.  Returns (assignments, iterations).
    exchange: a `kmcuda_b200.shard.PeerExchange` (ranks on one node with peer access): the partial sums are summed
    over peer memory in rank order instead of by the communicator's all-reduce."""
    n, K = X.shape[0], C.shape[0]
    dev = X.device
    assign = torch.full((n,), -1, dtype=torch.int32, device=dev)
    prev = torch.full((n,), -1, dtype=torch.int32, device=dev)
    changed = torch.zeros(1, dtype=torch.int32, device=dev)
    sums = torch.zeros((K, X.shape[1]), dtype=torch.float32, device=dev)
    counts = torch.zeros(K, dtype=torch.int32, device=dev)
    ccounts = torch.zeros(K, dtype=torch.int32, device=dev)
    if hasattr(backend, "reset"):
        backend.reset()          # a reused shard must not carry the previous run's member sums (angular update)
    it = 0
    while True:
        it += 1
        changed.zero_()
        backend.assign(X, C, assign, prev, changed)
        total_changed = allreduce_scalar(changed.item(), dev)     # .item() synchronises the stream
        if hasattr(backend, "last_error") and backend.last_error():
            raise RuntimeError("tensor-core pipeline error 0x%x: the assignment pass is invalid" % backend.last_error())
        if log:
            log("iteration %d: %d reassignments" % (it, total_changed))
        if float(total_changed) <= float(tolerance) * float(total_samples):
            break
        if max_iter and it >= max_iter:
            break
        if exchange is not None:
            exchange.update(backend, X, assign, sums, counts)
        else:
            backend.partial_sums(X, assign, sums, counts)
            allreduce_update(sums, counts)
        backend.finish_update(sums, counts, C, ccounts)
    return assign, it
