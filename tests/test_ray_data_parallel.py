"""The N>1 path: batch i -> rank i % N partition, independent replicas, final gather. world_size-2 gloo on CPU."""
import os
import socket
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

from ipu_ray_lib_b200 import scene
from ipu_ray_lib_b200.parallel import batch_owner_mask, merge_shards, scatter_shards

ROOT = Path(__file__).resolve().parents[1]


def test_partition_follows_reference_round_robin():
    n, per = 100_000, 8640
    masks = [batch_owner_mask(n, per, 4, r) for r in range(4)]
    assert np.array_equal(sum(m.astype(int) for m in masks), np.ones(n, int))  # disjoint cover
    batch = np.arange(n) // per
    for r, m in enumerate(masks):
        assert np.all(np.unique(batch[m]) % 4 == r)  # src/IpuScene.cpp:682: replica = i % numReplicas
    with pytest.raises(ValueError):
        batch_owner_mask(10, 4, 2, 2)


def test_scatter_merge_round_trip():
    rays = scene.init_ray_stream(173, 131, 0.7)
    for world in (1, 2, 3, 8):
        shards = scatter_shards(rays, 1000, world)
        assert sum(s.size for s in shards) == rays.size
        assert merge_shards(shards, rays.size, 1000).tobytes() == rays.tobytes()


WORKER = r"""
import os, sys
import numpy as np
import torch.distributed as dist
sys.path.insert(0, os.environ["REPO_ROOT"])
from ipu_ray_lib_b200 import scene
from ipu_ray_lib_b200.parallel import batch_owner_mask, gather_stream
from oracle.oracle_py import Oracle

dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
w, h, per = 96, 64, 512
s = scene.HostScene.builtin("spheres").configure(w, h, path_trace=True, samples=3, seed=9)
full = scene.init_ray_stream(w, h, s.fov)
mine = np.ascontiguousarray(full[batch_owner_mask(full.size, per, world, rank)])
Oracle("port").path_trace(s, mine, threads=1)       # stand-in renderer for the CPU-only test
merged = gather_stream(mine, full.size, per, dist, dst=0)
if rank == 0:
    whole = full.copy()
    Oracle("port").path_trace(s, whole, threads=1)
    assert merged.tobytes() == whole.tobytes(), "partitioned render differs from the single-replica render"
    print("GATHER_OK")
else:
    assert merged is None
dist.barrier()
dist.destroy_process_group()
"""


def test_two_replicas_gloo(tmp_path):
    """Per-(pixel,sample) RNG streams make the image independent of the replica count: 2 ranks == 1 rank, bit for bit."""
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    env = dict(os.environ, REPO_ROOT=str(ROOT), OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                       capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "GATHER_OK" in r.stdout
