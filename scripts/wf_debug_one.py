import sys
import numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from ipu_ray_lib_b200 import scene
from ipu_ray_lib_b200.render import B200Scene
w, h = 320, 200
s = scene.HostScene.builtin("box").configure(w, h, path_trace=True, samples=1, seed=1442, max_path_length=1)
base = scene.init_ray_stream(w, h, s.fov)
with B200Scene(s) as g:
    got = base.copy()
    g.execute(got, traversal=4, scene_residency=2)
    print(got[5518])
