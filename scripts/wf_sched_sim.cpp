// DESIGN TOOL (not part of the product): CPU model of wf_trace_kernel's warp scheduler.
//
// wf_trace_kernel is issue-bound, and what it issues is decided by a per-warp scheduler: 32 lanes, one query each, one
// phase (inner-node step / leaf test / fetch) per warp iteration. This program runs the kernel's own step functions
// (csrc/rt_prims.h "Streaming traversal") for 32 simulated lanes on real rays of a built-in scene and charges every
// warp iteration the issue slots the phase costs in SASS, so that scheduling policies can be compared in seconds on
// the CPU before anything is built for the GPU. The absolute numbers are a model; the ranking of policies is what it
// is for (measured on the B200 afterwards).
//
//   g++ -O2 -std=c++17 -ffp-contract=off -I/usr/local/cuda/include scripts/wf_sched_sim.cpp \
//       -Lipu_ray_lib_b200 -lb200rt_scene -Wl,-rpath,$PWD/ipu_ray_lib_b200 -o /tmp/wf_sched_sim
//   /tmp/wf_sched_sim box assets/monkey_bust.glb 192 1
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../include/b200rt_scene.h"
#include "../ipu_ray_lib_b200/csrc/scene_tables.hpp"

using namespace rt;

struct Ray { V3 o, d; };

static uint64_t g_rng = 0x9E3779B97F4A7C15ull;
static float urand() {
  g_rng ^= g_rng << 13; g_rng ^= g_rng >> 7; g_rng ^= g_rng << 17;
  return (float)((g_rng >> 40) & 0xFFFFFF) / 16777216.f;
}

struct View {
  SceneTables tables;
  DevScene dev{};
};
static void make_view(const b200rt_scene_desc& d, View& v) {
  const std::string err = build_scene_tables(d, v.tables);
  if (!err.empty()) { std::fprintf(stderr, "%s\n", err.c_str()); std::exit(1); }
  DevScene& s = v.dev;
  s.nodes = (const uint2*)d.bvh_nodes;
  s.geoms = v.tables.geoms.data();
  s.triVerts = (const float4*)v.tables.triVerts.data();
  s.triNormals = nullptr;
  s.triFaceNormals = (const float4*)v.tables.triFaceNormals.data();
  s.spheres = (const float4*)d.spheres;
  s.discs = d.discs;
  s.matIDs = d.mat_ids;
  s.materials = (const float*)d.materials;
  s.numNodes = d.num_bvh_nodes;
  s.numMaterials = d.num_materials;
  s.pairs = (const uint4*)v.tables.pairs.words.data();
  s.leafOrig = v.tables.pairs.leafOrig.data();
  s.numPairs = v.tables.pairs.numPairs;
  s.rootRef = v.tables.pairs.rootRef;
  s.rootGeom = v.tables.pairs.rootGeom;
  s.boundsFinite = v.tables.pairs.boundsFinite ? 1u : 0u;
  s.leafInfo = (const uint4*)v.tables.pairs.leafInfo.data();
  s.numTris = d.num_tris;
  s.numSpheres = d.num_spheres;
  s.trisBounded = v.tables.trisBounded ? 1u : 0u;
}

// one whole query, serially (to generate the next bounce's rays)
static bool trace_one(const DevScene& sc, const Ray& r, StreamQuery& q, std::vector<uint2>& stack) {
  stream_begin(sc, q, r.o, r.d);
  while (q.ref != kRefNone) {
    bool again = ref_is_inner(q.ref) ? stream_trav<true>(q, fetch_pair<false>(sc.pairs, ref_pair(q.ref)), stack.data())
                                     : stream_leaf<true>(sc, q, stack.data());
    while (again) again = stream_pop(q, stack.data());
  }
  return q.hitRef != kRefNone;
}

// diffuse bounce off the hit of `r` (cosine-weighted about the geometric normal on the ray's side)
static Ray bounce(const DevScene& sc, const Ray& r, const StreamQuery& q) {
  uint32_t geomID, primID, tri;
  stream_hit_ids(sc, q.hitRef, geomID, primID, tri);
  const V3 p = r.o + r.d * q.hitT;
  V3 n = prim_normal(sc, geomID, tri, q.b0, q.b1, q.b2, p);
  if (dot(n, r.d) > 0.f) n = -n;
  // orthonormal basis + cosine sample
  const V3 a = std::fabs(n.x) > 0.5f ? mk(0.f, 1.f, 0.f) : mk(1.f, 0.f, 0.f);
  const V3 t = normalized(cross(n, a)), b = cross(n, t);
  const float u1 = urand(), u2 = urand();
  const float rr = std::sqrt(u1), ph = 6.2831853f * u2;
  const float x = rr * std::cos(ph), y = rr * std::sin(ph), z = std::sqrt(std::max(0.f, 1.f - u1));
  Ray out;
  out.d = normalized(t * x + b * y + n * z);
  const float m = (1.f + std::max(std::fabs(p.x), std::max(std::fabs(p.y), std::fabs(p.z)))) * kRayEpsilon;
  out.o = p + n * m;
  return out;
}

struct Cost { double trav, leaf, fetch, head, head2; };

struct Lane {
  StreamQuery q;
  std::vector<uint2> stack;
  bool done = false;
};

struct Result {
  double slots = 0;
  double its[3] = {0, 0, 0}, lanes[3] = {0, 0, 0};
  uint64_t queries = 0;
};

// policy 0: inner-node steps while >= thr lanes want one, else the phase most lanes wait for (the kernel's)
// policy 1: the phase with the most lane-work per issue slot (count / cost)
// policy 2: like 0 but leaf / fetch run as soon as leafThr / fetchThr lanes wait for them
static Result simulate(const DevScene& sc, const std::vector<Ray>& rays, const Cost& c, int policy, int thr, int leafThr, int fetchThr) {
  Result res;
  const size_t warps = 64;  // independent warps sharing the ray pool, like one SM's worth
  size_t cursor = 0;
  for (size_t w = 0; w < warps; ++w) {
    std::vector<Lane> L(32);
    for (auto& l : L) { l.stack.assign(kMaxStack + 1, make_uint2(0u, 0u)); l.q.ref = kRefNone; }
    // each warp works through its own slice of the pool
    const size_t begin = rays.size() * w / warps, end = rays.size() * (w + 1) / warps;
    cursor = begin;
    while (true) {
      int cT = 0, cL = 0, cF = 0;
      for (auto& l : L) {
        if (l.done) continue;
        if (l.q.ref == kRefNone) cF++;
        else if (ref_is_inner(l.q.ref)) cT++;
        else cL++;
      }
      if (cT + cL + cF == 0) break;
      int pick = 0;
      double head = c.head;
      if (policy == 0) {
        if (cT < thr) { head += c.head2; pick = (cT >= cL && cT >= cF) ? 0 : (cL >= cF ? 1 : 2); }
      } else if (policy == 1) {
        head += c.head2;
        const double eT = cT / c.trav, eL = cL / c.leaf, eF = cF / c.fetch;
        pick = (eT >= eL && eT >= eF) ? 0 : (eL >= eF ? 1 : 2);
      } else {
        head += c.head2;
        if (cL >= leafThr && cL >= cF) pick = 1;
        else if (cF >= fetchThr) pick = 2;
        else if (cT > 0) pick = 0;
        else pick = cL >= cF ? 1 : 2;
      }
      const int n = pick == 0 ? cT : (pick == 1 ? cL : cF);
      res.slots += head + (pick == 0 ? c.trav : (pick == 1 ? c.leaf : c.fetch));
      res.its[pick] += 1; res.lanes[pick] += n;
      for (auto& l : L) {
        if (l.done) continue;
        bool again = false;
        if (pick == 0 && ref_is_inner(l.q.ref)) again = stream_trav<true>(l.q, fetch_pair<false>(sc.pairs, ref_pair(l.q.ref)), l.stack.data());
        else if (pick == 1 && ref_is_leaf(l.q.ref)) again = stream_leaf<true>(sc, l.q, l.stack.data());
        else if (pick == 2 && l.q.ref == kRefNone) {
          if (cursor >= end) { l.done = true; continue; }
          stream_begin(sc, l.q, rays[cursor].o, rays[cursor].d);
          cursor++; res.queries++;
        }
        while (again) again = stream_pop(l.q, l.stack.data());
      }
    }
  }
  return res;
}

// (A variant with POSTPONED leaf tests -- a lane parks the first leaf it reaches and goes on with the next deferred
// node, so that inner-node runs get longer -- was modelled here too: 0.49 efficiency against 0.52, the steps into nodes
// that the parked leaf's hit would have culled cost more than the longer runs gain. Removed with the single-entry
// register stack top it was written against; the result is recorded in DESIGN.md.)

// ---- variant: K queries per lane, each warp iteration advances (at most) one of them per lane ----------------------
static Result simulate_multi(const DevScene& sc, const std::vector<Ray>& rays, const Cost& c, int thr, int K, double extra) {
  Result res;
  const size_t warps = 64;
  for (size_t w = 0; w < warps; ++w) {
    std::vector<Lane> L(32 * K);
    for (auto& l : L) { l.stack.assign(kMaxStack + 1, make_uint2(0u, 0u)); l.q.ref = kRefNone; }
    const size_t begin = rays.size() * w / warps, end = rays.size() * (w + 1) / warps;
    size_t cursor = begin;
    auto phase = [](const Lane& l) { return l.done ? 3 : (l.q.ref == kRefNone ? 2 : (ref_is_inner(l.q.ref) ? 0 : 1)); };
    while (true) {
      int cnt[4] = {0, 0, 0, 0};
      for (int lane = 0; lane < 32; ++lane) {
        bool has[4] = {false, false, false, false};
        for (int k = 0; k < K; ++k) has[phase(L[lane * K + k])] = true;
        for (int p = 0; p < 3; ++p) cnt[p] += has[p] ? 1 : 0;
      }
      if (cnt[0] + cnt[1] + cnt[2] == 0) break;
      int pick = 0;
      double head = c.head;
      if (cnt[0] < thr) { head += c.head2; pick = (cnt[0] >= cnt[1] && cnt[0] >= cnt[2]) ? 0 : (cnt[1] >= cnt[2] ? 1 : 2); }
      res.slots += head + extra + (pick == 0 ? c.trav : (pick == 1 ? c.leaf : c.fetch));
      res.its[pick] += 1; res.lanes[pick] += cnt[pick];
      for (int lane = 0; lane < 32; ++lane) {
        for (int k = 0; k < K; ++k) {
          Lane& l = L[lane * K + k];
          if (phase(l) != pick) continue;
          bool again = false;
          if (pick == 0) again = stream_trav<true>(l.q, fetch_pair<false>(sc.pairs, ref_pair(l.q.ref)), l.stack.data());
          else if (pick == 1) again = stream_leaf<true>(sc, l.q, l.stack.data());
          else {
            if (cursor >= end) { l.done = true; break; }
            stream_begin(sc, l.q, rays[cursor].o, rays[cursor].d);
            cursor++; res.queries++;
          }
          while (again) again = stream_pop(l.q, l.stack.data());
          break;  // one context per lane and iteration
        }
      }
    }
  }
  return res;
}

int main(int argc, char** argv) {
  const char* name = argc > 1 ? argv[1] : "box";
  const char* mesh = argc > 2 ? argv[2] : "assets/monkey_bust.glb";
  const int w = argc > 3 ? std::atoi(argv[3]) : 192;
  const int nb = argc > 4 ? std::atoi(argv[4]) : 1;  // which bounce's rays to schedule (0 = camera rays)
  b200rt_host_scene* hs = nullptr;
  if (b200rt_host_scene_builtin(name, mesh, &hs) != 0) { std::fprintf(stderr, "%s\n", b200rt_scene_last_error()); return 1; }
  b200rt_scene_desc d{};
  b200rt_host_scene_desc(hs, &d);
  View v;
  make_view(d, v);

  // camera rays (no jitter), then nb diffuse bounces
  std::vector<float> tr((size_t)w * w * 21);
  b200rt_init_ray_stream(tr.data(), w, w, w, w, 0, 0, d.fov_radians);
  std::vector<Ray> rays;
  for (size_t i = 0; i < (size_t)w * w; ++i) {
    const float* t = &tr[i * 21];
    Ray r;
    r.o = mk(t[5], t[6], t[7]); r.d = mk(t[9], t[10], t[11]);
    r.o = r.o + mk(0.f, 0.f, 1.f) * (kRayEpsilon * (dot(mk(0.f, 0.f, 1.f), r.d) < 0 ? -1.f : 1.f));
    rays.push_back(r);
  }
  std::vector<uint2> stack(kMaxStack + 1, make_uint2(0u, 0u));
  for (int b = 0; b < nb; ++b) {
    std::vector<Ray> next;
    for (const Ray& r : rays) {
      StreamQuery q;
      if (trace_one(v.dev, r, q, stack)) next.push_back(bounce(v.dev, r, q));
    }
    rays.swap(next);
  }
  std::printf("%zu rays of bounce %d\n", rays.size(), nb);

  // issue slots per warp iteration, from the SASS of wf_trace_kernel<true,false,false>
  Cost c{105, 190, 290, 12, 14};
  if (const char* e = std::getenv("SIM_COST")) std::sscanf(e, "%lf,%lf,%lf,%lf,%lf", &c.trav, &c.leaf, &c.fetch, &c.head, &c.head2);
  auto report = [&](const char* label, const Result& r) {
    const double ideal = (r.lanes[0] * c.trav + r.lanes[1] * c.leaf + r.lanes[2] * c.fetch) / 32.0;
    std::printf("%-34s slots/query %7.1f  (ideal %6.1f, eff %.3f)  T %.1f its x %4.1f lanes | L %.2f x %4.1f | F %.2f x %4.1f   per 32 queries\n", label,
                r.slots / r.queries, ideal / r.queries, ideal / r.slots, 32 * r.its[0] / r.queries, r.lanes[0] / r.its[0],
                32 * r.its[1] / r.queries, r.lanes[1] / r.its[1], 32 * r.its[2] / r.queries, r.lanes[2] / r.its[2]);
  };
  for (int thr : {1, 4, 8, 12, 16, 20, 24}) {
    char label[64];
    std::snprintf(label, sizeof label, "policy 0 (kernel), thr %d", thr);
    report(label, simulate(v.dev, rays, c, 0, thr, 0, 0));
  }
  const double extra = std::getenv("SIM_EXTRA") ? std::atof(std::getenv("SIM_EXTRA")) : 12.0;  // slots per iteration for routing contexts
  for (int K : {2, 3, 4})
    for (int thr : {12, 16, 24, 28}) {
      char label[64];
      std::snprintf(label, sizeof label, "%d queries per lane, thr %d (+%.0f)", K, thr, extra);
      report(label, simulate_multi(v.dev, rays, c, thr, K, extra));
    }
  report("policy 1 (lanes per slot)", simulate(v.dev, rays, c, 1, 0, 0, 0));
  for (int lt : {4, 8, 12, 16})
    for (int ft : {4, 8, 12, 16}) {
      char label[64];
      std::snprintf(label, sizeof label, "policy 2, leaf >= %d, fetch >= %d", lt, ft);
      report(label, simulate(v.dev, rays, c, 2, 0, lt, ft));
    }
  b200rt_host_scene_free(hs);
  return 0;
}
