"""bench.py's control flow on CPU (tests/_bench_dryrun.py stubs everything that touches CUDA): one JSON line with the
contract's keys at 1 rank, the peer-memory iteration leg reported at 2 ranks, and the NCCL numbers kept -- with every rank
taking the same branch -- when the exchange fails on a rank."""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("world,fail,expect", [(1, 0, "none (1 GPU)"), (2, 0, "peer memory"), (2, 1, "not reported")])
def test_bench_control_flow(world, fail, expect):
    env = dict(os.environ, DRY_WORLD=str(world), DRY_PEER_FAIL=str(fail))
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"):
        env.pop(k, None)
    r = subprocess.run([sys.executable, os.path.join(HERE, "_bench_dryrun.py")], env=env, stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, text=True, timeout=300)
    assert "DRYRUN OK" in r.stdout and expect in r.stdout, r.stdout[-1500:]
