"""A/B aid: per-kernel device times (the library's own CUDA-event spans) of the wavefront tracer on the bench scene,
without the NIF stage, repeated; prints min / median per launch so that builds can be compared below the run-to-run noise."""
import os
import sys
import numpy as np
import torch
sys.path.insert(0, '.')
from ipu_ray_lib_b200 import HostScene, init_ray_stream
from ipu_ray_lib_b200.render import B200Scene

spp = int(sys.argv[1]) if len(sys.argv) > 1 else 64
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 7
s = HostScene.builtin('box').configure(1440, 1440, path_trace=True, samples=spp, seed=1442)
rays = init_ray_stream(1440, 1440, s.fov)
dev = torch.from_numpy(rays.view(np.uint8).reshape(-1)).cuda()
pristine = dev.clone()
tr, sh, tot = [], [], []
extra = {'scene_residency': int(os.environ['WF_RESIDENCY'])} if 'WF_RESIDENCY' in os.environ else {}
with B200Scene(s) as g:
    for r in range(reps + 2):
        dev.copy_(pristine)
        torch.cuda.synchronize()
        g.execute_device(dev.data_ptr(), rays.size, stream=torch.cuda.current_stream().cuda_stream, **extra)
        torch.cuda.synchronize()
        st = g.stats()
        if r >= 2:
            tr.append(st["trace_kernel_ms"] / st["trace_kernel_launches"])
            sh.append(st["shade_kernel_ms"] / st["shade_kernel_launches"])
            tot.append(st["kernel_ms"])
print(f"trace ms/launch min {min(tr):.4f} median {np.median(tr):.4f} | shade min {min(sh):.4f} median {np.median(sh):.4f} | launches {st['trace_kernel_launches']} | step ms min {min(tot):.3f} median {np.median(tot):.3f}")
