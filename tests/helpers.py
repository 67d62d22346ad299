"""Hand-built scenes for edge-case tests: flat arrays in the reference layouts + a BVH from the scene library."""
import ctypes as C

import numpy as np

from ipu_ray_lib_b200 import _capi as capi


class CustomScene:
    """Quacks like HostScene for the oracle / B200Scene (only `.desc` and `.fov` are used)."""

    def __init__(self, meshes=(), spheres=(), discs=(), materials=None, mat_ids=None, normals=False):
        self.keep = []
        geom, info, tris, verts, nrms, bounds, ids = [], [], [], [], [], [], []
        for m, (v, t) in enumerate(meshes):
            v = np.asarray(v, np.float32).reshape(-1, 3)
            t = np.asarray(t, np.uint16).reshape(-1, 3)
            info.append((len(tris), len(verts), len(t), len(v)))
            for k, tri in enumerate(t):
                p = v[tri]
                bounds.append(np.concatenate([p.min(0), p.max(0)]))
                ids.append((len(geom), k))
            tris.extend(t.tolist())
            verts.extend(v.tolist())
            if normals:
                n = np.cross(v[t[:, 1]] - v[t[:, 0]], v[t[:, 2]] - v[t[:, 0]])
                vn = np.zeros_like(v)
                for tri, nn in zip(t, n):
                    vn[tri] += nn
                vn /= np.maximum(np.linalg.norm(vn, axis=1, keepdims=True), 1e-20)
                nrms.extend(vn.tolist())
            geom.append((m, 0, 0))
        for i, s in enumerate(spheres):
            s = np.asarray(s, np.float32)
            bounds.append(np.concatenate([s[:3] - s[3], s[:3] + s[3]]))
            ids.append((len(geom), 0))
            geom.append((i, 1, 0))
        for i, d in enumerate(discs):
            d = np.asarray(d, np.float32)  # nx,ny,nz,r,cx,cy,cz
            bounds.append(np.concatenate([d[4:7] - d[3], d[4:7] + d[3]]))
            ids.append((len(geom), 0))
            geom.append((i, 2, 0))
        n_geom = len(geom)
        self.geometry = np.array(geom, dtype=capi.GEOM_REF) if geom else np.zeros(0, capi.GEOM_REF)
        self.mesh_info = np.array(info, dtype=capi.MESH_INFO) if info else np.zeros(0, capi.MESH_INFO)
        self.mesh_tris = np.asarray(tris, np.uint16).reshape(-1, 3)
        self.mesh_verts = np.asarray(verts, np.float32).reshape(-1, 3)
        self.mesh_normals = np.asarray(nrms, np.float32).reshape(-1, 3)
        self.spheres = np.asarray(spheres, np.float32).reshape(-1, 4)
        self.discs = np.asarray(discs, np.float32).reshape(-1, 7)
        if materials is None:
            materials = np.zeros(1, capi.MATERIAL)
            materials["albedo"] = 0.75
            materials["ior"] = 1.52
        self.materials = np.ascontiguousarray(materials)
        self.mat_ids = np.asarray(mat_ids if mat_ids is not None else [0] * n_geom, np.uint32)
        b = np.ascontiguousarray(np.asarray(bounds, np.float32))
        i = np.ascontiguousarray(np.asarray(ids, np.uint32))
        self.bvh_nodes = np.zeros(2 * len(ids) - 1, capi.BVH_NODE)
        depth = C.c_uint32()
        n = capi.scene_lib().b200rt_build_bvh(capi.ptr(b), capi.ptr(i), len(ids), capi.ptr(self.bvh_nodes), C.byref(depth))
        assert n == self.bvh_nodes.size, capi.scene_lib().b200rt_scene_last_error()
        d = self.desc = capi.SceneDesc()
        for name, arr, cnt in (("geometry", self.geometry, "num_geometry"), ("mesh_info", self.mesh_info, "num_meshes"),
                               ("mesh_tris", self.mesh_tris, "num_tris"), ("mesh_verts", self.mesh_verts, "num_verts"),
                               ("mesh_normals", self.mesh_normals, "num_normals"), ("mat_ids", self.mat_ids, "num_mat_ids"),
                               ("materials", self.materials, "num_materials"), ("bvh_nodes", self.bvh_nodes, "num_bvh_nodes"),
                               ("spheres", self.spheres, "num_spheres"), ("discs", self.discs, "num_discs")):
            setattr(d, name, arr.ctypes.data if arr.size else None)
            setattr(d, cnt, arr.shape[0])
        d.max_leaf_depth = depth.value
        d.fov_radians = 0.7853982
        d.anti_alias_scale = 0.25
        d.max_path_length = 10
        d.roulette_start_depth = 3
        d.samples_per_pixel = 4
        d.rng_seed = 1442
        d.path_trace = 1
        d.device = -1

    @property
    def fov(self):
        return self.desc.fov_radians

    def configure(self, width, height, *, path_trace=True, samples=4, seed=1442, anti_alias=0.25,
                  max_path_length=10, roulette_start_depth=3):
        d = self.desc
        d.image_width, d.image_height = float(width), float(height)
        d.path_trace = int(path_trace)
        d.samples_per_pixel = samples
        d.rng_seed = seed
        d.anti_alias_scale = anti_alias
        d.max_path_length = max_path_length
        d.roulette_start_depth = roulette_start_depth
        return self


def make_rays(origins, directions, t_min=0.0, t_max=np.inf):
    origins = np.asarray(origins, np.float32).reshape(-1, 3)
    rays = np.zeros(origins.shape[0], capi.RAY)
    rays["origin"] = origins
    rays["direction"] = np.asarray(directions, np.float32).reshape(-1, 3)
    rays["tMin"] = t_min
    rays["tMax"] = t_max
    return rays
