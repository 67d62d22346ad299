"""N ranks == 1 rank, byte for byte, on real GPUs through NCCL: one process per B200 (torch.distributed.run), each
rendering the batches i % N == rank with the CUDA path, stream gathered to rank 0 over NCCL and compared with the
1-rank render of the same process group and with the oracle. Skipped where fewer than 2 B200s are visible."""
import os
import socket
import subprocess
import sys
from pathlib import Path

import pytest

from ipu_ray_lib_b200 import _capi as capi

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]

WORKER = r"""
import os, sys
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, os.environ["REPO_ROOT"])
from ipu_ray_lib_b200 import scene
from ipu_ray_lib_b200.nif import NifWeights
from ipu_ray_lib_b200.parallel import batch_owner_mask, gather_stream
from ipu_ray_lib_b200.render import B200Scene
from oracle.oracle_py import Oracle

local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
w, h, per, spp = 480, 300, 8640, 6
s = scene.HostScene.builtin("box").configure(w, h, path_trace=True, samples=spp, seed=1442, device=local)
full = scene.init_ray_stream(w, h, s.fov)
nif = NifWeights.synthetic(seed=1442)
for use_nif in (False, True):
    mine = np.ascontiguousarray(full[batch_owner_mask(full.size, per, world, rank)])
    with B200Scene(s) as g:
        if use_nif:
            g.load_nif_model(nif)
        g.execute(mine)
    merged = gather_stream(mine, full.size, per, dist, dst=0)
    if rank == 0:
        whole = full.copy()
        with B200Scene(s) as g:
            if use_nif:
                g.load_nif_model(nif)
            g.execute(whole)
        assert merged.tobytes() == whole.tobytes(), f"{world}-rank render differs from the 1-rank render (nif={use_nif})"
        if not use_nif:
            want = full.copy()
            Oracle("port").path_trace(s, want)
            assert merged.tobytes() == want.tobytes(), "multi-rank render differs from the oracle"
    else:
        assert merged is None
if rank == 0:
    print("MULTI_GPU_OK", world)
dist.barrier()
dist.destroy_process_group()
"""


@pytest.mark.parametrize("world", [2, 4])
def test_n_ranks_equal_one_rank_over_nccl(tmp_path, world):
    if capi.lib().b200rt_device_count() < world:
        pytest.skip(f"needs {world} B200s")
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    env = dict(os.environ, REPO_ROOT=str(ROOT), NCCL_DEBUG="WARN")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                       capture_output=True, text=True, env=env, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert f"MULTI_GPU_OK {world}" in r.stdout
