"""First GPU contact: parity vs the compiled reference kernels + rough timings per kernel variant."""
import sys, time, json
import numpy as np
sys.path.insert(0, '.')
from ipu_ray_lib_b200 import scene, _capi as capi
from ipu_ray_lib_b200.render import B200Scene
from oracle.oracle_py import Oracle

ref = Oracle('reference')

def diff_report(name, a, b):
    same = a.tobytes() == b.tobytes()
    print(f"[{name}] bit-identical: {same}")
    if not same:
        av = a.view(np.uint32).reshape(a.size, 21); bv = b.view(np.uint32).reshape(b.size, 21)
        bad = np.nonzero((av != bv).any(axis=1))[0]
        print(f"   differing rays: {bad.size}/{a.size}; word histogram: {(av != bv).sum(axis=0)}")
        for i in bad[:3]:
            print("   gpu", a[i]); print("   ref", b[i])
    return same

for scn in ['box', 'spheres']:
    s = scene.HostScene.builtin(scn)
    W = H = 512
    s.configure(W, H, path_trace=False)
    base = scene.init_ray_stream(W, H, s.fov)
    r_ref = base.copy(); cref = ref.shadow_trace(s, r_ref)
    with B200Scene(s) as g:
        for trav in (1, 2):
            for res in (1, 2):
                r = base.copy()
                g.execute(r, traversal=trav, scene_residency=res, count_visits=1)
                st = g.stats()
                ok = diff_report(f"{scn} shadow trav={trav} res={res}", r, r_ref)
                print("   stats", {k: st[k] for k in ('closest_hit_queries','occlusion_queries','node_visits','prim_tests','kernel_ms')}, "ref", cref)
    # path trace parity
    W = H = 128
    s.configure(W, H, path_trace=True, samples=8)
    base = scene.init_ray_stream(W, H, s.fov)
    r_ref = base.copy(); cref = ref.path_trace(s, r_ref)
    with B200Scene(s) as g:
        for trav in (1, 2):
            for res in (1, 2):
                r = base.copy()
                g.execute(r, traversal=trav, scene_residency=res, count_visits=1)
                st = g.stats()
                diff_report(f"{scn} path trav={trav} res={res}", r, r_ref)
                print("   stats", {k: st[k] for k in ('closest_hit_queries','samples','escaped_samples','node_visits','prim_tests','kernel_ms')}, "ref", cref)

# timings at 1440^2 on the box scene
s = scene.HostScene.builtin('box')
W = H = 1440
for mode, spp in (('shadow', 1), ('path', 16)):
    s.configure(W, H, path_trace=(mode == 'path'), samples=spp)
    base = scene.init_ray_stream(W, H, s.fov)
    with B200Scene(s) as g:
        for trav in (1, 2):
            for res in (1, 2):
                best = None
                for it in range(3):
                    r = base.copy()
                    g.execute(r, traversal=trav, scene_residency=res)
                    st = g.stats()
                    best = st if best is None or st['kernel_ms'] < best['kernel_ms'] else best
                q = best['closest_hit_queries'] + best['occlusion_queries']
                print(f"[time] {mode} 1440^2 spp={spp} trav={trav} res={res}: kernel {best['kernel_ms']:.2f} ms, "
                      f"{q/best['kernel_ms']/1e3:.1f} Mrays/s, h2d {best['h2d_ms']:.1f} ms d2h {best['d2h_ms']:.1f} ms")
