"""Debug aid: wavefront path tracer vs the oracle, one bounce at a time, with a breakdown of what differs."""
import sys
import numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from ipu_ray_lib_b200 import scene
from ipu_ray_lib_b200.render import B200Scene
from oracle.oracle_py import Oracle

port = Oracle("port")
w, h = 320, 200
for mpl, spp in ((1, 1), (3, 2)):
    s = scene.HostScene.builtin("box").configure(w, h, path_trace=True, samples=spp, seed=1442, max_path_length=mpl)
    base = scene.init_ray_stream(w, h, s.fov)
    want = base.copy()
    cw = port.path_trace(s, want)
    with B200Scene(s) as g:
        for kw in (dict(traversal=4),):
            got = base.copy()
            g.execute(got, **kw)
            st = g.stats()
            bad = np.nonzero(got.view(np.uint8).reshape(got.size, -1) != want.view(np.uint8).reshape(want.size, -1))[0]
            bad = np.unique(bad)
            print(f"maxPath {mpl} spp {spp} {kw}: {bad.size}/{got.size} rays differ; queries {st['closest_hit_queries']} vs {cw['closest_hit_queries']}, "
                  f"escaped {st['escaped_samples']} vs {cw['escaped_samples']}")
            if bad.size:
                wg = want['h']['geomID'][bad] if 'h' in want.dtype.names else None
                print("   first", bad[:8], " oracle geomIDs of differing rays:", np.unique(wg, return_counts=True) if wg is not None else '')
                print("   got ", got[bad[0]])
                print("   want", want[bad[0]])
                m = np.zeros(got.size, bool); m[bad] = True
                m = m.reshape(h, w)
                for r0 in range(0, h, 10):
                    print("   ", "".join(".:*#"[min(3, int(4 * m[r0:r0 + 10, c0:c0 + 8].mean()))] for c0 in range(0, w, 8)))
                wp = want["h"]["primID"][bad] if "h" in want.dtype.names else None
                print("   oracle (geomID, primID) of differing rays:", sorted(set(zip(want["h"]["geomID"][bad].tolist(), wp.tolist())))[:40])
