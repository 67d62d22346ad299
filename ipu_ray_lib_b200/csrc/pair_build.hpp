// Host-side construction of the pair table (layout: rt_prims.h "Pair table") from the caller's
// CompactBVH2Node[] (include/CompactBVH2Node.hpp:52-85), with the structural validation a serialised
// scene from outside needs: child indices, leaf ids, primitive ranges, tree-ness and depth are all
// checked here once so that no kernel can be driven out of bounds by a malformed node array.
#pragma once
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "rt_prims.h"

namespace rt {

struct PairTable {
  std::vector<uint32_t> words;     // 12 per pair
  std::vector<uint32_t> leafOrig;  // 2 per pair
  std::vector<uint32_t> leafInfo;  // 4 per primitive slot (triangles, spheres, discs): DevScene::leafInfo
  uint32_t numPairs = 0;
  uint32_t rootRef = 0, rootGeom = kInvalidGeom;
  uint32_t maxDepth = 0;           // deepest leaf (root = depth 0) = bound on the traversal stack
  bool boundsFinite = true;
};

// `primCount[g]` = number of primitives geometry g may be addressed with (mesh: its triangles; sphere/disc: unbounded,
// pass 0xFFFFFFFF). Returns an empty string on success, else what is wrong with the node array.
// numTris / numSpheres / numDiscs size the per-primitive table (leafInfo).
inline std::string build_pair_table(const void* nodes24, uint32_t numNodes, const GeomEntry* geoms, uint32_t numGeoms,
                                    const uint32_t* primCount, uint32_t numTris, uint32_t numSpheres, uint32_t numDiscs,
                                    PairTable& out) {
  struct Node { float mn[3]; uint32_t primOrSecond; uint16_t d[3]; uint16_t geomID; };
  static_assert(sizeof(Node) == 24, "CompactBVH2Node");
  const Node* nodes = static_cast<const Node*>(nodes24);
  out = PairTable{};
  if (numNodes == 0) return "empty node array";
  const uint64_t numSlots = (uint64_t)numTris + numSpheres + numDiscs;
  if (numSlots > kLeafIndexMask) return "more than 2^30 primitives";
  out.leafInfo.assign((size_t)numSlots * 4, 0u);
  for (size_t i = 0; i < (size_t)numSlots; ++i) {
    out.leafInfo[4 * i] = kInvalidGeom; out.leafInfo[4 * i + 1] = kInvalidPrim; out.leafInfo[4 * i + 2] = 0xFFFFFFFFu;
  }
  // a leaf: record geomID / reported primID / node index under its primitive (the lowest node index wins when a
  // primitive sits in several leaves: that is the leaf the reference's pre-order walk meets first)
  auto note_leaf = [&](uint32_t node, uint32_t ref, std::string& err) {
    const uint32_t type = ref >> 30, index = ref & kLeafIndexMask;
    const uint64_t limit = type == 0u ? numTris : (type == 1u ? numSpheres : numDiscs);
    if (index >= limit) { err = "leaf primitive index out of range"; return; }
    const size_t slot = (size_t)index + (type == 0u ? 0u : (type == 1u ? numTris : numTris + numSpheres));
    uint32_t* info = &out.leafInfo[4 * slot];
    if (node < info[2]) {
      info[0] = nodes[node].geomID;
      info[1] = type == 0u ? nodes[node].primOrSecond : 0u;
      info[2] = node;
    }
  };

  auto child_ref = [&](uint32_t i, uint32_t& ref, std::string& err) {
    const Node& n = nodes[i];
    if (n.geomID == kInvalidGeom) return;  // inner: filled in when its pair index is known
    if (n.geomID >= numGeoms) { err = "leaf geomID out of range"; return; }
    const GeomEntry g = geoms[n.geomID];
    if (g.type == 0u) {
      if (n.primOrSecond >= primCount[n.geomID]) { err = "leaf primID outside its mesh"; return; }
      const uint64_t tri = (uint64_t)g.first + n.primOrSecond;
      if (tri > kLeafIndexMask) { err = "more than 2^30 triangles"; return; }
      ref = (uint32_t)tri;
    } else {
      if (g.first > kLeafIndexMask) { err = "primitive index too large"; return; }
      ref = (g.type << 30) | g.first;
    }
  };
  for (uint32_t i = 0; i < numNodes; ++i) {
    const Node& n = nodes[i];
    for (int k = 0; k < 3; ++k) {
      const float ext = half_bits_to_float(n.d[k]);
      if (!std::isfinite(n.mn[k]) || !std::isfinite(ext) || !std::isfinite(n.mn[k] + ext)) out.boundsFinite = false;
    }
  }

  // pre-order walk from the root: numbers the inner nodes, checks that every node is reached exactly once
  std::vector<uint32_t> pairOf(numNodes, 0xFFFFFFFFu);
  std::vector<uint8_t> seen(numNodes, 0);
  struct Item { uint32_t node, depth; };
  std::vector<Item> stack;
  stack.push_back({0u, 0u});
  std::vector<uint32_t> innerOrder;
  while (!stack.empty()) {
    const Item it = stack.back();
    stack.pop_back();
    if (it.node >= numNodes) return "child index out of range";
    if (seen[it.node]) return "node reachable twice (not a tree)";
    seen[it.node] = 1;
    const Node& n = nodes[it.node];
    if (n.geomID != kInvalidGeom) {
      if (it.depth > out.maxDepth) out.maxDepth = it.depth;
      continue;
    }
    if (it.node + 1 >= numNodes || n.primOrSecond >= numNodes) return "child index out of range";
    if (n.primOrSecond <= it.node + 1) return "second child does not follow the first child's subtree";
    pairOf[it.node] = (uint32_t)innerOrder.size();
    innerOrder.push_back(it.node);
    stack.push_back({n.primOrSecond, it.depth + 1});
    stack.push_back({it.node + 1, it.depth + 1});
  }
  if (out.maxDepth > (uint32_t)kMaxStack) return "BVH deeper than 64 levels";

  out.numPairs = (uint32_t)innerOrder.size();
  out.words.assign((size_t)out.numPairs * 12, 0u);
  out.leafOrig.assign((size_t)out.numPairs * 2, 0u);
  std::string err;
  for (uint32_t p = 0; p < out.numPairs; ++p) {
    const uint32_t parent = innerOrder[p];
    const uint32_t kids[2] = {parent + 1, nodes[parent].primOrSecond};
    uint32_t* w = &out.words[(size_t)p * 12];
    for (int side = 0; side < 2; ++side) {
      const Node& c = nodes[kids[side]];
      uint32_t ref = kRefInner | pairOf[kids[side]];
      child_ref(kids[side], ref, err);
      if (err.empty() && c.geomID != kInvalidGeom) note_leaf(kids[side], ref, err);
      if (!err.empty()) return err;
      uint32_t* d = w + 6 * side;
      std::memcpy(d, c.mn, 12);
      d[3] = ref;
      d[4] = (uint32_t)c.d[0] | ((uint32_t)c.d[1] << 16);
      d[5] = (uint32_t)c.d[2] | ((uint32_t)c.geomID << 16);
      out.leafOrig[(size_t)p * 2 + side] = kids[side];
    }
  }
  out.rootGeom = nodes[0].geomID;
  out.rootRef = kRefInner | 0u;
  if (out.rootGeom != kInvalidGeom) {
    child_ref(0u, out.rootRef, err);
    if (err.empty()) note_leaf(0u, out.rootRef, err);
    if (!err.empty()) return err;
  }
  return std::string();
}

}  // namespace rt
