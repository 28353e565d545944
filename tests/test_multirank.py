"""Multi-rank host logic: world_size-2 gloo on CPU (always), nccl + NVLink p2p on >= 2 GPUs (gpu marker)."""
import os
import subprocess
import sys

import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))


def _torchrun(nproc, *args, port=29533):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(HERE, "mgpu_worker.py"), *args]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    return subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)


def test_fedavg_exchange_gloo_world2():
    r = _torchrun(2, "gloo", "nccl", port=29541)
    assert r.returncode == 0 and "MGPU_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


def test_client_assignment():
    from federated_multi_modal_b200.fed import clients_of_rank
    assert [clients_of_rank(32, r, 8) for r in (0, 7)] == [[0, 1, 2, 3], [28, 29, 30, 31]]
    assert sorted(sum((clients_of_rank(8, r, 2) for r in range(2)), [])) == list(range(8))
    with pytest.raises(ValueError):
        clients_of_rank(10, 0, 4)


def test_sharded_exchange_ranges_tile_the_arena():
    from federated_multi_modal_b200.fed import shard_range
    for n in (1, 3, 4, 5, 64, 12352, 13856768, 13856771):
        for world in (1, 2, 3, 4, 8):
            r = [shard_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(lo <= hi and (lo % 4 == 0 or lo == hi) for lo, hi in r)     # non-empty shards start on a float4
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))   # contiguous, no overlap, no gap


@pytest.mark.gpu
@pytest.mark.parametrize("transport", ["p2p", "p2p_sharded", "nccl"])
def test_fedavg_exchange_nccl(transport):
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = min(n, 8)
    r = _torchrun(world, "nccl", transport, port=29551 if transport == "p2p" else 29552)
    assert r.returncode == 0 and "MGPU_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
    if transport.startswith("p2p"):  # no silent fallback to the all-gather transport
        assert f"transport={transport}\n" in r.stdout + "\n", r.stdout[-2000:]
