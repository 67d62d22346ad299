"""A/B aid for the chunk overlap (b200rt_trace_params.chunk_overlap): step time and the library's per-kernel spans of the
headline workload (box scene + NIF, 1440^2) at a reduced sample count, device-resident rays; also checks that the
frame is byte-identical to the non-overlapped one."""
import os
import sys
import numpy as np
import torch
sys.path.insert(0, '.')
from ipu_ray_lib_b200 import HostScene, init_ray_stream
from ipu_ray_lib_b200.nif import NifWeights
from ipu_ray_lib_b200.render import B200Scene

spp = int(sys.argv[1]) if len(sys.argv) > 1 else 256
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
extra = {}
for k in ('scene_residency', 'chunk_overlap', 'samples_per_chunk'):
    if 'WF_' + k.upper() in os.environ:
        extra[k] = int(os.environ['WF_' + k.upper()])
s = HostScene.builtin('box').configure(1440, 1440, path_trace=True, samples=spp, seed=1442)
rays = init_ray_stream(1440, 1440, s.fov)
dev = torch.from_numpy(rays.view(np.uint8).reshape(-1)).cuda()
pristine = dev.clone()
tot, ev = [], []
with B200Scene(s) as g:
    g.load_nif_model(NifWeights.synthetic(seed=1442))
    for r in range(reps + 1):
        dev.copy_(pristine)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        g.execute_device(dev.data_ptr(), rays.size, stream=torch.cuda.current_stream().cuda_stream, **extra)
        b.record()
        torch.cuda.synchronize()
        st = g.stats()
        if r >= 1:
            tot.append(st["kernel_ms"])
            ev.append(a.elapsed_time(b))
    digest = int(dev.view(torch.int32).to(torch.int64).sum().item())
    if 'WF_CHECK' in os.environ:  # same frame with the overlap off
        dev2 = pristine.clone()
        g.execute_device(dev2.data_ptr(), rays.size, stream=torch.cuda.current_stream().cuda_stream, **dict(extra, chunk_overlap=1))
        torch.cuda.synchronize()
        print("identical to chunk_overlap=1:", bool(torch.equal(dev, dev2)))
print(f"spp {spp} {extra} step ms min {min(tot):.2f} median {np.median(tot):.2f} (events {min(ev):.2f}) | per launch: trace "
      f"{st['trace_kernel_ms'] / st['trace_kernel_launches']:.3f} shade {st['shade_kernel_ms'] / st['shade_kernel_launches']:.3f} "
      f"nif {st['nif_kernel_ms'] / max(st['nif_kernel_launches'], 1):.2f} acc {st['accumulate_kernel_ms']:.2f} | digest {digest}")
