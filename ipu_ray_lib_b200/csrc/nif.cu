// NIF (neural image field) environment light: (u,v) -> Fourier features -> dense ReLU MLP -> decode.
//
// What it computes (reference: the Poplar graph built by src/neural_networks/NifModel.cpp):
//   encode  (:186-219, host twin :417-433)  un = 2*(u-1), vn = 2*(v-1); for j < E: a = fp16(un*2^j) ...
//           features = [sin(un*2^j)]_E, [sin(vn*2^j)]_E, [cos(un*2^j)]_E, [cos(vn*2^j)]_E   (fp16)
//   MLP     (:296-327)  x = act(x.W + b) per Dense layer, fp16 weights and fp16 layer outputs; where a
//           layer's input width differs from the running width the encoded input is concatenated first
//           (auto-detected skip connection, :303-309)
//   decode  (:222-246)  fp32: exp(x*max + mean)   (mean already has -eps folded in, NifMetaData.cpp:48-53)
// Numerics chosen for B200: products of fp16 operands accumulated in fp32 (the IPU used fp16
// partials, src/IpuScene.cpp:256-262), one rounding to fp16 per layer output.
//
// This translation unit is compiled with FMA contraction enabled: nothing here is bit-compared with the
// reference (it has no CPU NIF); parity is checked against oracle/oracle_port.cpp within a stated tolerance.
#include "nif.cuh"

#include <cuda_fp16.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#ifdef B200RT_NIF_PAIR_KERNEL
#include "nif_tc_pair2.cuh"  // experiment build: the CTA-pair kernel, selected with B200RT_NIF_PAIR=2
#else
#include "nif_tc.cuh"
#endif

namespace rt {

namespace {
thread_local std::string g_nifError;

constexpr int kMaxLayers = 16;
constexpr int kTileRows = 16;     // rows per CTA pass in the CUDA-core kernel
constexpr int kThreads = 256;
constexpr int kMaxWidth = 512;    // widest activation (incl. concatenated input) the smem tile holds

struct LayerDev {
  const __half* w;     // [in][out] row-major
  const __half* b;     // [out] or nullptr
  int in, out;
  int relu;
  int concat;          // 1: input of this layer = concat(previous activations, encoded features)
};

struct NifParams {
  LayerDev layers[kMaxLayers];
  int numLayers;
  int embed;           // E; feature width = 4E
  float maxv, mean0, mean1, mean2;
  int logToneMap;
};
}  // namespace

struct NifModel {
  NifParams p{};
  std::vector<void*> allocs;
  int device = 0;
  // tensor-core path
  bool tcOk = false;
  tc::Params tc{};
  size_t tcSmem = 0;
  std::string tcWhyNot;
#ifdef B200RT_NIF_PAIR_KERNEL
  tc::Params tc2{};   // tiling of the CTA-pair kernel (nif_tc_pair2.cuh): 128 / 192 column split
  tc2::PairMaps maps{};
  bool tc2Ok = false;
  size_t pair2Smem = 0;
#endif
};

namespace {

// Builds the B-operand images for the tcgen05 kernel: per layer the four blocks B0..B3 of nif_tc.cuh back to back,
// each block stored as [k/8][n within its column half][k%8] fp16 (zero padded). K order of the lo blocks:
// [rows multiplying the previous layer's first 160 outputs | bias row + 15 zero rows (the "ones" slice) | rows
// multiplying the encoded input]; of the hi blocks: [rows multiplying the remaining outputs]. Returns false (with a
// reason) when the model does not fit the kernel's tiling rules.
bool prepare_tc(NifModel* m, const b200rt_nif_desc& d) {
  auto no = [&](const std::string& why) { m->tcWhyNot = why; return false; };
  const int E = (int)d.embedding_dimension, F = 4 * E;
  if (F % 16 != 0) return no("feature width 4E is not a multiple of 16");
  if (2 + F / 8 > tc::kStaticPlanesMax) return no("encoded input wider than the static region");
  if ((int)d.num_layers > tc::kMaxLayers) return no("too many layers");
  tc::Params& t = m->tc;
  t = tc::Params{};
  t.numLayers = (int)d.num_layers;
  t.embed = E;
  int width = F, widthPad = F, maxHidden = 16, prevN0 = 0;
  // actRows = K rows read from the (padded) activation planes; realActRows = how many of them carry real weights
  std::vector<int> actRows(d.num_layers), featRows(d.num_layers), realActRows(d.num_layers);
  for (uint32_t i = 0; i < d.num_layers; ++i) {
    const b200rt_nif_layer& L = d.layers[i];
    tc::Layer& o = t.layers[i];
    const int K = (int)L.in_features;
    o.N = (int)L.out_features; o.relu = L.relu;
    const bool last = i + 1 == d.num_layers;
    // The layer pipeline is built for column halves of exactly kHalfN accumulator columns. Any hidden width up to
    // 2 * kHalfN runs on it zero-padded to 160 or 320 columns: padded weight columns (and their bias entries) are zero, so
    // the padded activations are relu(0) = 0, and the K rows of the next layer that would multiply them are zero too --
    // every real output is the same sum of the same products (fp32 accumulation of exact zeros changes nothing).
    if (!last && o.N > 2 * tc::kHalfN) return no("hidden layer wider than 320");
    o.Npad = last ? (o.N + 15) / 16 * 16 : (o.N <= tc::kHalfN ? tc::kHalfN : 2 * tc::kHalfN);
    if (last && o.Npad > tc::kHalfN) return no("output layer wider than one accumulator slot");
    o.n0 = std::min(o.Npad, tc::kHalfN);
    o.n1 = o.Npad - o.n0;
    if (i == 0) {
      if (K != F) return no("first layer does not take the encoded input");
      actRows[i] = 0; featRows[i] = F;
    } else if (K == width + F) {  // skip-concat (NifModel.cpp:303-309)
      actRows[i] = widthPad; featRows[i] = F;
    } else if (K == width) {
      actRows[i] = widthPad; featRows[i] = 0;
    } else {
      return no("layer input width mismatch");
    }
    realActRows[i] = i == 0 ? 0 : width;
    o.actLoSlices = std::min(actRows[i], prevN0) / 16;
    o.actHiSlices = actRows[i] / 16 - o.actLoSlices;
    o.staticSlices = 1 + featRows[i] / 16;
    width = o.N;
    widthPad = o.Npad;
    prevN0 = o.n0;
    if (!last) maxHidden = std::max(maxHidden, o.Npad);
  }
  if (width != 3) return no("last layer must have 3 outputs");
  t.actPlanes = maxHidden / 8;
  t.maxv = d.max; t.mean0 = d.mean[0]; t.mean1 = d.mean[1]; t.mean2 = d.mean[2];
  t.logToneMap = d.log_tone_map;
  m->tcSmem = (size_t)(t.actPlanes + tc::kStaticPlanesMax) * tc::kPlaneBytes + (size_t)tc::kStages * tc::kStageBytes +
              (2 * tc::kStages + 4) * 8 + 16;
  int maxSmem = 0;
  cudaDeviceGetAttribute(&maxSmem, cudaDevAttrMaxSharedMemoryPerBlockOptin, m->device);
  if (m->tcSmem > (size_t)maxSmem) return no("activation tile + weight ring exceed shared memory");

  for (uint32_t i = 0; i < d.num_layers; ++i) {
    const b200rt_nif_layer& L = d.layers[i];
    tc::Layer& o = t.layers[i];
    const __half* src = reinterpret_cast<const __half*>(L.kernel_f16);
    const __half* bias = reinterpret_cast<const __half*>(L.bias_f16);
    const int loAct = 16 * o.actLoSlices, hiAct = 16 * o.actHiSlices;
    const int loK = loAct + 16 * o.staticSlices;
    std::vector<__half> img;
    // one block: kRows K-rows x nCols columns starting at column nBase; value(k, n) supplies the entries
    auto block = [&](int kRows, int nBase, int nCols, auto&& value) {
      const size_t at = img.size();
      img.resize(at + (size_t)kRows * nCols, __float2half(0.f));
      for (int k = 0; k < kRows; ++k)
        for (int n = 0; n < nCols; ++n)
          if (nBase + n < o.N) img[at + ((size_t)(k / 8) * nCols + n) * 8 + (k % 8)] = value(k, nBase + n);
    };
    const int realAct = realActRows[i];  // activation rows beyond this are padding (zero weights)
    auto loValue = [&](int k, int n) -> __half {
      if (k < loAct) return k < realAct ? src[(size_t)k * o.N + n] : __float2half(0.f);
      if (k == loAct) return bias ? bias[n] : __float2half(0.f);
      if (k < loAct + 16) return __float2half(0.f);
      return src[(size_t)(realAct + (k - loAct - 16)) * o.N + n];  // encoded-input rows follow the activation rows
    };
    auto hiValue = [&](int k, int n) -> __half {
      return loAct + k < realAct ? src[(size_t)(loAct + k) * o.N + n] : __float2half(0.f);
    };
    block(loK, 0, o.n0, loValue);
    if (o.n1) block(loK, o.n0, o.n1, loValue);
    if (hiAct) block(hiAct, 0, o.n0, hiValue);
    if (hiAct && o.n1) block(hiAct, o.n0, o.n1, hiValue);
    void* dw = nullptr;
    if (cudaMalloc(&dw, img.size() * 2) != cudaSuccess) return no("cudaMalloc failed");
    m->allocs.push_back(dw);
    cudaMemcpy(dw, img.data(), img.size() * 2, cudaMemcpyHostToDevice);
    o.wimg = (const __half*)dw;
  }
  if (cudaFuncSetAttribute(tc::nif_mlp_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)m->tcSmem) != cudaSuccess) {
    cudaGetLastError();
    return no("cudaFuncSetAttribute(max dynamic shared memory) failed");
  }
  return true;
}

#ifdef B200RT_NIF_PAIR_KERNEL
// Same for the CTA-pair kernel (nif_tc_pair2.cuh): N0 = 128 / N1 = 192 columns, K_lo = the previous layer's first 128
// outputs (+ ones slice + encoded input), and only the per-rank half images (columns [r n/2, (r+1) n/2) of every block).
// Requires hidden width 320. All layers' images go into one buffer that the kernel reads through TMA tensor maps over
// its 128-byte rows (one map per stage height).
bool prepare_tc2(NifModel* m, const b200rt_nif_desc& d) {
  const int E = (int)d.embedding_dimension, F = 4 * E;
  if (F % 16 != 0 || 2 + F / 8 > tc2::kStaticPlanesMax || (int)d.num_layers > tc::kMaxLayers) return false;
  tc::Params& t = m->tc2;
  t = tc::Params{};
  t.numLayers = (int)d.num_layers;
  t.embed = E;
  int width = F, prevN0 = 0;
  std::vector<int> actRows(d.num_layers), featRows(d.num_layers);
  for (uint32_t i = 0; i < d.num_layers; ++i) {
    const b200rt_nif_layer& L = d.layers[i];
    tc::Layer& o = t.layers[i];
    const int K = (int)L.in_features;
    o.N = (int)L.out_features; o.Npad = (o.N + 15) / 16 * 16; o.relu = L.relu;
    const bool last = i + 1 == d.num_layers;
    if (!last && o.N != tc2::kN0 + tc2::kN1) return false;
    if (last && o.Npad > 16) return false;
    o.n0 = std::min(o.Npad, tc2::kN0);
    o.n1 = o.Npad - o.n0;
    if (i == 0) { if (K != F) return false; actRows[i] = 0; featRows[i] = F; }
    else if (K == width + F) { actRows[i] = width; featRows[i] = F; }
    else if (K == width) { actRows[i] = width; featRows[i] = 0; }
    else return false;
    o.actLoSlices = std::min(actRows[i], prevN0) / 16;
    o.actHiSlices = actRows[i] / 16 - o.actLoSlices;
    o.staticSlices = 1 + featRows[i] / 16;
    width = o.N;
    prevN0 = o.n0;
  }
  if (width != 3) return false;
  t.maxv = d.max; t.mean0 = d.mean[0]; t.mean1 = d.mean[1]; t.mean2 = d.mean[2];
  t.logToneMap = d.log_tone_map;
  m->pair2Smem = (size_t)(tc2::kActPlanes + tc2::kStaticPlanesMax) * tc2::kPlaneBytes + (size_t)tc2::kStages * tc2::kStageBytes +
                 (2 * tc2::kStages + 6) * 8 + 16;
  int maxSmem = 0;
  cudaDeviceGetAttribute(&maxSmem, cudaDevAttrMaxSharedMemoryPerBlockOptin, m->device);
  if (m->pair2Smem > (size_t)maxSmem) return false;
  std::vector<__half> all;
  for (uint32_t i = 0; i < d.num_layers; ++i) {
    const b200rt_nif_layer& L = d.layers[i];
    tc::Layer& o = t.layers[i];
    const __half* src = reinterpret_cast<const __half*>(L.kernel_f16);
    const __half* bias = reinterpret_cast<const __half*>(L.bias_f16);
    const int loAct = 16 * o.actLoSlices, hiAct = 16 * o.actHiSlices;
    const int loK = loAct + 16 * o.staticSlices;
    std::vector<__half> img;
    auto block = [&](int kRows, int nBase, int nCols, auto&& value) {
      const size_t at = img.size();
      img.resize(at + (size_t)kRows * nCols, __float2half(0.f));
      for (int k = 0; k < kRows; ++k)
        for (int n = 0; n < nCols; ++n)
          if (nBase + n < o.N) img[at + ((size_t)(k / 8) * nCols + n) * 8 + (k % 8)] = value(k, nBase + n);
    };
    auto loValue = [&](int k, int n) -> __half {
      if (k < loAct) return src[(size_t)k * o.N + n];
      if (k == loAct) return bias ? bias[n] : __float2half(0.f);
      if (k < loAct + 16) return __float2half(0.f);
      return src[(size_t)(actRows[i] + (k - loAct - 16)) * o.N + n];
    };
    auto hiValue = [&](int k, int n) -> __half { return src[(size_t)(loAct + k) * o.N + n]; };
    o.pairRow0 = (uint32_t)(all.size() * 2 / 128);
    for (int r = 0; r < 2; ++r) {
      img.clear();
      block(loK, r * (o.n0 / 2), o.n0 / 2, loValue);
      if (o.n1) block(loK, o.n0 + r * (o.n1 / 2), o.n1 / 2, loValue);
      if (hiAct) block(hiAct, r * (o.n0 / 2), o.n0 / 2, hiValue);
      if (hiAct && o.n1) block(hiAct, o.n0 + r * (o.n1 / 2), o.n1 / 2, hiValue);
      if ((img.size() * 2) % 128 != 0) return false;
      if (r == 0) o.pairRankRows = (uint32_t)(img.size() * 2 / 128);
      all.insert(all.end(), img.begin(), img.end());
    }
    o.wimg = nullptr;
  }
  // slack rows behind the last image: a box never reaches past the tensor
  const size_t rows = all.size() * 2 / 128;
  void* dp = nullptr;
  if (cudaMalloc(&dp, (rows + 72) * 128) != cudaSuccess) return false;
  m->allocs.push_back(dp);
  cudaMemset(dp, 0, (rows + 72) * 128);
  cudaMemcpy(dp, all.data(), all.size() * 2, cudaMemcpyHostToDevice);
  // tensor maps (driver entry point fetched through the runtime: no libcuda link)
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn) { cudaGetLastError(); return false; }
  const uint32_t planeClass[3] = {128u, 1024u, 1536u};
  for (int c = 0; c < 3; ++c)
    for (int np = 2; np <= 6; np += 2) {
      const cuuint64_t gdim[2] = {32, (cuuint64_t)(rows + 72)};
      const cuuint64_t gstride[1] = {128};
      const cuuint32_t box[2] = {32, (cuuint32_t)(np * planeClass[c] / 128)};
      const cuuint32_t estr[2] = {1, 1};
      const CUresult rc = ((EncodeFn)fn)(&m->maps.m[tc2::pair_map_index(planeClass[c], (uint32_t)np)], CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, dp,
                                         gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (rc != CUDA_SUCCESS) { std::fprintf(stderr, "[nif pair] cuTensorMapEncodeTiled failed: %d (class %d planes %d)\n", (int)rc, c, np); return false; }
    }
  if (cudaFuncSetAttribute(tc2::nif_mlp_tc_pair2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)m->pair2Smem) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return true;
}
#endif

}  // namespace

const char* nif_last_error() { return g_nifError.c_str(); }

void nif_destroy(NifModel* m) {
  if (!m) return;
  for (void* a : m->allocs) cudaFree(a);
  delete m;
}

NifModel* nif_create(const b200rt_nif_desc& d, int device) {
  auto bad = [&](const std::string& s) -> NifModel* { g_nifError = s; return nullptr; };
  if (d.num_layers == 0 || d.num_layers > (uint32_t)kMaxLayers || !d.layers) return bad("bad layer count");
  if (d.embedding_dimension == 0 || d.embedding_dimension > 16) return bad("bad embedding dimension");
  const int feat = 4 * (int)d.embedding_dimension;
  NifModel* m = new NifModel();
  m->device = device;
  m->p.numLayers = (int)d.num_layers;
  m->p.embed = (int)d.embedding_dimension;
  m->p.maxv = d.max;
  m->p.mean0 = d.mean[0]; m->p.mean1 = d.mean[1]; m->p.mean2 = d.mean[2];
  m->p.logToneMap = d.log_tone_map;
  int width = feat;
  for (uint32_t i = 0; i < d.num_layers; ++i) {
    const b200rt_nif_layer& L = d.layers[i];
    LayerDev& o = m->p.layers[i];
    o.in = (int)L.in_features; o.out = (int)L.out_features; o.relu = L.relu; o.concat = 0;
    if (!L.kernel_f16 || o.in <= 0 || o.out <= 0) { nif_destroy(m); return bad("bad layer"); }
    if (o.in != width) {
      if (o.in == width + feat) o.concat = 1;
      else { nif_destroy(m); return bad("layer input width matches neither the running width nor width + features"); }
    }
    if (o.in > kMaxWidth || o.out > kMaxWidth) { nif_destroy(m); return bad("layer wider than 512"); }
    width = o.out;
    void* w = nullptr;
    const size_t wb = (size_t)o.in * o.out * sizeof(__half);
    if (cudaMalloc(&w, wb) != cudaSuccess || cudaMemcpy(w, L.kernel_f16, wb, cudaMemcpyHostToDevice) != cudaSuccess) {
      if (w) cudaFree(w);
      nif_destroy(m);
      return bad("cudaMalloc/cudaMemcpy of a NIF kernel failed");
    }
    m->allocs.push_back(w);
    o.w = (const __half*)w;
    o.b = nullptr;
    if (L.bias_f16) {
      void* b = nullptr;
      const size_t bb = (size_t)o.out * sizeof(__half);
      if (cudaMalloc(&b, bb) != cudaSuccess || cudaMemcpy(b, L.bias_f16, bb, cudaMemcpyHostToDevice) != cudaSuccess) {
        if (b) cudaFree(b);
        nif_destroy(m);
        return bad("cudaMalloc/cudaMemcpy of a NIF bias failed");
      }
      m->allocs.push_back(b);
      o.b = (const __half*)b;
    }
  }
  if (width != 3) { nif_destroy(m); return bad("last NIF layer must have 3 outputs (b,g,r)"); }
  // Tensor-core path (the product path). Models outside its tiling rules (widths not multiples of 16)
  // run on the CUDA-core kernel below; B200RT_NIF_IMPL=simt forces that kernel for debugging.
  m->tcOk = prepare_tc(m, d);
#ifdef B200RT_NIF_PAIR_KERNEL
  m->tc2Ok = m->tcOk && prepare_tc2(m, d);
#endif
  const char* impl = std::getenv("B200RT_NIF_IMPL");
  if (impl && std::strcmp(impl, "simt") == 0) { m->tcOk = false; m->tcWhyNot = "forced by B200RT_NIF_IMPL=simt"; }
  // Weight uploads came from pageable memory on the default stream; the kernels run on the caller's non-blocking
  // stream, so make sure every DMA has landed before the model is handed out.
  cudaDeviceSynchronize();
  return m;
}

namespace {

// Encode one (u,v) into 4E fp16 features (NifModel.cpp:186-219).
__device__ __forceinline__ void encode_uv(float u, float v, int E, __half* feat) {
  const float un = (u - 1.f) * 2.f, vn = (v - 1.f) * 2.f;
  float c = 1.f;
  for (int j = 0; j < E; ++j, c *= 2.f) {
    const float au = __half2float(__float2half_rn(un * c));
    const float av = __half2float(__float2half_rn(vn * c));
    feat[j] = __float2half_rn(sinf(au));
    feat[j + E] = __float2half_rn(sinf(av));
    feat[j + 2 * E] = __float2half_rn(cosf(au));
    feat[j + 3 * E] = __float2half_rn(cosf(av));
  }
}

// CUDA-core reference implementation of the MLP wavefront: one CTA evaluates kTileRows samples per
// pass, activations ping-pong between two fp16 shared-memory tiles, each thread owns output columns
// tid, tid+256, ... and keeps kTileRows fp32 accumulators per column.
__global__ void __launch_bounds__(kThreads) nif_mlp_kernel(const NifParams p, const float* __restrict__ uvDirect,
                                                           const float* __restrict__ slotEscape,
                                                           const uint32_t* __restrict__ queue,
                                                           const uint32_t* __restrict__ dCount, uint32_t directCount,
                                                           uint32_t first, float* __restrict__ out) {
  __shared__ __half actA[kTileRows][kMaxWidth + 64];
  __shared__ __half actB[kTileRows][kMaxWidth + 64];
  __shared__ __half feat[kTileRows][64];
  __shared__ uint32_t rowSlot[kTileRows];
  const uint32_t count = uvDirect ? directCount : min(*dCount > first ? *dCount - first : 0u, directCount);
  const int F = 4 * p.embed;
  for (uint32_t tile = blockIdx.x; (uint64_t)tile * kTileRows < count; tile += gridDim.x) {
    const uint32_t row0 = tile * kTileRows;
    __syncthreads();
    if (threadIdx.x < kTileRows) {
      const uint32_t r = row0 + threadIdx.x;
      float u = 0.f, v = 0.f;
      uint32_t slot = 0xFFFFFFFFu;
      if (r < count) {
        if (uvDirect) { slot = r; u = uvDirect[2 * r]; v = uvDirect[2 * r + 1]; }
        else { slot = queue[r]; u = slotEscape[5 * (size_t)slot + 3]; v = slotEscape[5 * (size_t)slot + 4]; }
      }
      rowSlot[threadIdx.x] = slot;
      encode_uv(u, v, p.embed, feat[threadIdx.x]);
      for (int k = 0; k < F; ++k) actA[threadIdx.x][k] = feat[threadIdx.x][k];
    }
    __syncthreads();
    __half (*cur)[kMaxWidth + 64] = actA;
    __half (*nxt)[kMaxWidth + 64] = actB;
    int width = F;
    for (int l = 0; l < p.numLayers; ++l) {
      const LayerDev L = p.layers[l];
      if (L.concat) {
        for (int i = threadIdx.x; i < kTileRows * F; i += kThreads) cur[i / F][width + i % F] = feat[i / F][i % F];
        __syncthreads();
      }
      for (int col = threadIdx.x; col < L.out; col += kThreads) {
        float acc[kTileRows];
#pragma unroll
        for (int r = 0; r < kTileRows; ++r) acc[r] = 0.f;
        for (int k = 0; k < L.in; ++k) {
          const float w = __half2float(L.w[(size_t)k * L.out + col]);
#pragma unroll
          for (int r = 0; r < kTileRows; ++r) acc[r] += __half2float(cur[r][k]) * w;
        }
        const float bias = L.b ? __half2float(L.b[col]) : 0.f;
#pragma unroll
        for (int r = 0; r < kTileRows; ++r) {
          float y = acc[r] + bias;
          if (L.relu) y = y > 0.f ? y : 0.f;
          nxt[r][col] = __float2half_rn(y);
        }
      }
      __syncthreads();
      __half (*t)[kMaxWidth + 64] = cur; cur = nxt; nxt = t;
      width = L.out;
    }
    // decode (NifModel.cpp:222-246): fp32 x*max + mean, exp when log-tone-mapped
    if (threadIdx.x < kTileRows * 3) {
      const int r = threadIdx.x / 3, c = threadIdx.x % 3;
      const uint32_t slot = rowSlot[r];
      if (slot != 0xFFFFFFFFu) {
        const float mean = c == 0 ? p.mean0 : (c == 1 ? p.mean1 : p.mean2);
        float y = __half2float(cur[r][c]) * p.maxv + mean;
        if (p.logToneMap) y = expf(y);
        out[3 * (size_t)slot + c] = y;
      }
    }
  }
}

}  // namespace

static int launch(NifModel* m, const float* uvDirect, const float* slotEscape, const uint32_t* queue,
                  const uint32_t* dCount, uint32_t count, uint32_t first, float* out, cudaStream_t stream, int* launches) {
  if (!m) { g_nifError = "no model"; return -1; }
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, m->device);
  if (m->tcOk) {
    const uint32_t tcTiles = (count + tc::kRows - 1) / tc::kRows;
    // one persistent CTA per SM; B200RT_NIF_GRID caps the number of CTAs (measurements: how the kernel scales with SMs)
    static const uint32_t envGrid = [] { const char* e = std::getenv("B200RT_NIF_GRID"); return e ? (uint32_t)std::atoi(e) : 0u; }();
    const uint32_t maxGrid = envGrid ? std::min(envGrid, (uint32_t)sms) : (uint32_t)sms;
    const uint32_t tcGrid = tcTiles < maxGrid ? (tcTiles ? tcTiles : 1u) : maxGrid;
    static const bool profile = [] { const char* e = std::getenv("B200RT_NIF_PROFILE"); return e && e[0] == '1'; }();
    tc::Params params = m->tc;
    unsigned long long* dProf = nullptr;
    if (profile) {
      cudaMalloc(&dProf, (size_t)tcGrid * 16 * sizeof(unsigned long long));
      cudaMemsetAsync(dProf, 0, (size_t)tcGrid * 16 * sizeof(unsigned long long), stream);
      params.prof = dProf;
    }
    cudaError_t te;
#ifdef B200RT_NIF_PAIR_KERNEL
    static const int pairMode = [] { const char* e = std::getenv("B200RT_NIF_PAIR"); return e ? std::atoi(e) : 0; }();
    if (pairMode == 2 && m->tc2Ok && sms >= 2) {  // CTA-pair kernel (nif_tc_pair2.cuh)
      tc::Params params2 = m->tc2;
      params2.prof = params.prof;
      const uint32_t groups = (tcTiles + 1u) / 2u;
      const uint32_t maxPairs = maxGrid / 2u;
      const uint32_t pairs = groups < maxPairs ? (groups ? groups : 1u) : maxPairs;
      cudaLaunchConfig_t cfg{};
      cfg.gridDim = dim3(2u * pairs);
      cfg.blockDim = dim3(tc2::kThreads);
      cfg.dynamicSmemBytes = m->pair2Smem;
      cfg.stream = stream;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      te = cudaLaunchKernelEx(&cfg, tc2::nif_mlp_tc_pair2_kernel, params2, m->maps, uvDirect, slotEscape, queue, dCount, count, first, out);
    } else
#endif
    {
      tc::nif_mlp_tc_kernel<<<tcGrid, tc::kThreads, m->tcSmem, stream>>>(params, uvDirect, slotEscape, queue, dCount, count, first, out);
      te = cudaGetLastError();
    }
    if (te != cudaSuccess) { g_nifError = cudaGetErrorString(te); return -1; }
    if (profile) {  // debugging aid: per-role cycle breakdown of CTA 0 (synchronises!)
      std::vector<unsigned long long> h((size_t)tcGrid * 16);
      cudaStreamSynchronize(stream);
      cudaMemcpy(h.data(), dProf, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
      cudaFree(dProf);
      const char* names[] = {"total", "producer wait empty", "mma wait act", "mma wait weights", "mma phase (issue..commit)",
                             "epi wait acc", "epi encode", "epi drain", "tiles"};
      std::fprintf(stderr, "[nif profile] CTA0 of %u, rows %u:", tcGrid, count);
      const double tiles = h[tc::PF_TILES] ? (double)h[tc::PF_TILES] : 1.0;
      for (int i = 0; i < tc::PF_COUNT; ++i) std::fprintf(stderr, " %s=%.0f/tile", names[i], (double)h[i] / tiles);
      std::fprintf(stderr, "\n");
    }
    if (launches) *launches += 1;
    return 0;
  }
  const uint32_t tiles = (count + kTileRows - 1) / kTileRows;
  const uint32_t grid = tiles < (uint32_t)(sms * 4) ? (tiles ? tiles : 1u) : (uint32_t)(sms * 4);
  nif_mlp_kernel<<<grid, kThreads, 0, stream>>>(m->p, uvDirect, slotEscape, queue, dCount, count, first, out);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { g_nifError = cudaGetErrorString(e); return -1; }
  if (launches) *launches += 1;
  return 0;
}

int nif_eval_uv(NifModel* m, const float* dUv, uint32_t n, float* dBgrOut, cudaStream_t stream, int* launches) {
  return launch(m, dUv, nullptr, nullptr, nullptr, n, 0u, dBgrOut, stream, launches);
}

// maxBatch = IpuScene::setMaxNifBatchSize (src/IpuScene.cpp:265-327, :342-344): the queue is evaluated in serial
// launches of at most that many escaped rays (0 = one launch). The queue length lives on the device, so launches are
// issued for the whole possible range and those past its end find nothing to do.
int nif_eval_queue(NifModel* m, const float* slotEscape, const uint32_t* queue, const uint32_t* dCount,
                   uint32_t maxCount, uint32_t maxBatch, float* slotEnv, cudaStream_t stream, int* launches) {
  if (maxBatch == 0 || maxBatch >= maxCount) return launch(m, nullptr, slotEscape, queue, dCount, maxCount, 0u, slotEnv, stream, launches);
  for (uint32_t first = 0; first < maxCount; first += maxBatch)
    if (int rc = launch(m, nullptr, slotEscape, queue + first, dCount, std::min(maxBatch, maxCount - first), first, slotEnv, stream, launches))
      return rc;
  return 0;
}

}  // namespace rt
