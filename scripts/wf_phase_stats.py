import sys, numpy as np
sys.path.insert(0,'.')
from ipu_ray_lib_b200 import HostScene, init_ray_stream
from ipu_ray_lib_b200.render import B200Scene
s=HostScene.builtin('box').configure(1440,1440,path_trace=True,samples=8,seed=1442)
rays=init_ray_stream(1440,1440,s.fov)
with B200Scene(s) as g:
    g.execute(rays, count_visits=1, traversal=4)
    print(g.stats()['node_visits']/g.stats()['closest_hit_queries'])
