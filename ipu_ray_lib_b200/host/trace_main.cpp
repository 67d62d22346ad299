// `trace` — command line front end with the reference's flags (trace.cpp:338-378), rendering through
// B200Scene on 1..8 B200s. It is the slot `renderIPU` occupies in the reference's main()
// (trace.cpp:270-336, :519-524): build or import the scene, make the ray stream, run the device
// path, visualise an AOV and write `<prefix>_<vis>_b200.exr`.
//
// The reference's main() also renders CPU and Embree images unless --ipu-only is given. This
// binary has no CPU renderer on purpose (the CPU restatement lives under oracle/ as a test-only
// checker), so --ipu-only is accepted and is always in effect.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <regex>
#include <stdexcept>
#include <string>
#include <vector>

#include "B200Scene.hpp"
#include "scene_build.hpp"

using namespace b200rt;

namespace {

struct Options {
  std::string outprefix = "out", crop, meshFile, nifHdri, scene = "box", visualise = "rgb", renderMode = "path-trace",
              logLevel = "info", saveRayStream, builtinMesh = "../assets/monkey_bust.glb";
  std::uint32_t ipus = 4, maxPathLength = 10, rouletteStartDepth = 3, samples = 256;
  std::size_t raysPerWorker = 1, maxNifBatchSize = 0;
  std::int32_t width = 768, height = 432;
  float antiAlias = .25f, hdriRotation = 0.f, availableMemoryProportion = 0.6f;
  std::uint64_t seed = 1442;
  bool loadNormals = false, ipuOnly = false, rayCallback = false, nifSynthetic = false;
};

const char* kHelp =
    "Options (same names and defaults as the reference's trace):\n"
    "  --help                          Show command help.\n"
    "  -o, --outprefix arg (=out)      Set the output filename prefix.\n"
    "  --ipus arg (=4)                 Number of devices (each GPU is a replica; clamped to the GPUs present).\n"
    "  --rays-per-worker arg (=1)      Ray-batch granularity: a batch is rays-per-worker * 8640 rays.\n"
    "  -w, --width arg (=768)          Set rendered image width.\n"
    "  -h, --height arg (=432)         Set rendered image height.\n"
    "  --crop arg                      Window to render, format wxh+c+r.\n"
    "  --anti-alias arg (=0.25)        Width of anti-aliasing noise distribution in pixels.\n"
    "  --mesh-file arg                 Scene file (.dae / .glb with a camera). Default: built-in scene.\n"
    "  --nif-hdri arg                  Path to the 'assets.extra' directory of a NIF model.\n"
    "  --nif-synthetic                 (extension) use fixed-seed synthetic NIF weights shaped by the metadata.\n"
    "  --save-ray-stream arg           (extension) write the raw TraceResult stream (rgb = sum over samples) to a file.\n"
    "  --builtin-mesh arg              (extension) mesh placed in the built-in box scene (=../assets/monkey_bust.glb).\n"
    "  --hdri-rotation arg (=0)        Azimuthal rotation for HDRI environment map (degrees).\n"
    "  --load-normals                  Load and interpolate mesh normals from --mesh-file.\n"
    "  --scene arg (=box)              One of [box-simple, box, spheres].\n"
    "  --visualise arg (=rgb)          One of [rgb, normal, hitpoint, tfar, color, id].\n"
    "  --render-mode arg (=path-trace) One of [shadow-trace, path-trace].\n"
    "  --max-path-length arg (=10)     Max path length for path tracing.\n"
    "  --roulette-start-depth arg (=3) Path length after which rays can be randomly terminated.\n"
    "  --samples arg (=256)            Number of samples per pixel for path tracing.\n"
    "  --seed arg (=1442)              RNG seed.\n"
    "  --available-memory-proportion arg (=0.6)  Accepted for compatibility; ignored.\n"
    "  --max-nif-batch-size arg (=0)   Maximum batch-size for the NIF network (0 = auto).\n"
    "  --ipu-only                      Accepted for compatibility (this binary only renders on the device).\n"
    "  --ipu-ray-callback              Retrieve partial results batch by batch via the callback mechanism.\n"
    "  --log-level arg (=info)         trace, debug, info, warn, err, critical, off.\n";

int logRank(const std::string& l) {
  static const std::map<std::string, int> m = {{"trace", 0}, {"debug", 1}, {"info", 2}, {"warn", 3},
                                                {"err", 4},   {"critical", 5}, {"off", 6}};
  auto it = m.find(l);
  if (it == m.end()) throw std::runtime_error("Invalid log-level: '" + l + "'");
  return it->second;
}
int g_log = 2;
#define LOG(level, ...)                                   \
  do {                                                    \
    if (logRank(level) >= g_log) {                        \
      std::fprintf(stderr, "[%c] ", level[0] - 32);       \
      std::fprintf(stderr, __VA_ARGS__);                  \
      std::fprintf(stderr, "\n");                         \
    }                                                     \
  } while (0)

Options parse(int argc, char** argv) {
  Options o;
  auto need = [&](int& i) -> std::string {
    if (i + 1 >= argc) throw std::runtime_error(std::string("the required argument for option '") + argv[i] + "' is missing");
    return argv[++i];
  };
  for (int i = 1; i < argc; ++i) {
    std::string a = argv[i], v;
    const auto eq = a.find('=');
    bool inlineValue = false;
    if (a.rfind("--", 0) == 0 && eq != std::string::npos) { v = a.substr(eq + 1); a = a.substr(0, eq); inlineValue = true; }
    auto val = [&]() { return inlineValue ? v : need(i); };
    if (a == "--help") { std::fputs(kHelp, stdout); throw std::runtime_error("Show help"); }
    else if (a == "-o" || a == "--outprefix") o.outprefix = val();
    else if (a == "--ipus") o.ipus = (std::uint32_t)std::stoul(val());
    else if (a == "--rays-per-worker") o.raysPerWorker = std::stoul(val());
    else if (a == "-w" || a == "--width") o.width = std::stoi(val());
    else if (a == "-h" || a == "--height") o.height = std::stoi(val());
    else if (a == "--crop") o.crop = val();
    else if (a == "--anti-alias") o.antiAlias = std::stof(val());
    else if (a == "--mesh-file") o.meshFile = val();
    else if (a == "--nif-hdri") o.nifHdri = val();
    else if (a == "--nif-synthetic") o.nifSynthetic = true;
    else if (a == "--save-ray-stream") o.saveRayStream = val();
    else if (a == "--builtin-mesh") o.builtinMesh = val();
    else if (a == "--hdri-rotation") o.hdriRotation = std::stof(val());
    else if (a == "--load-normals") o.loadNormals = true;
    else if (a == "--scene") o.scene = val();
    else if (a == "--visualise") o.visualise = val();
    else if (a == "--render-mode") o.renderMode = val();
    else if (a == "--max-path-length") o.maxPathLength = (std::uint32_t)std::stoul(val());
    else if (a == "--roulette-start-depth") o.rouletteStartDepth = (std::uint32_t)std::stoul(val());
    else if (a == "--samples") o.samples = (std::uint32_t)std::stoul(val());
    else if (a == "--seed") o.seed = std::stoull(val());
    else if (a == "--available-memory-proportion") o.availableMemoryProportion = std::stof(val());
    else if (a == "--max-nif-batch-size") o.maxNifBatchSize = std::stoul(val());
    else if (a == "--ipu-only") o.ipuOnly = true;
    else if (a == "--ipu-ray-callback") o.rayCallback = true;
    else if (a == "--log-level") o.logLevel = val();
    else throw std::runtime_error("unrecognised option '" + a + "'");
  }
  static const std::map<std::string, int> vis = {{"rgb", 0}, {"id", 1}, {"normal", 2}, {"tfar", 3}, {"color", 4}, {"hitpoint", 5}};
  if (!vis.count(o.visualise)) throw std::runtime_error("the argument for option '--visualise' is invalid");
  if (o.renderMode != "shadow-trace" && o.renderMode != "path-trace")
    throw std::runtime_error("the argument for option '--render-mode' is invalid");
  if (o.meshFile.empty() && o.loadNormals)
    throw std::runtime_error("Option 'load-normals' is not valid without the 'mesh-file' option");
  if (o.renderMode == "path-trace" && o.visualise != "rgb")
    throw std::runtime_error("Running path-tracing without visualise=rgb is not advised.");  // app_utils.cpp:241-243
  return o;
}

CropWindow parseCrop(const std::string& fmt, int w, int h) {  // parseCropString (app_utils.cpp:212-233)
  if (fmt.empty()) return CropWindow{w, h, 0, 0};
  std::smatch m;
  if (!std::regex_search(fmt, m, std::regex("(\\d+)x(\\d+)\\+(\\d+)\\+(\\d+)")) || m.size() != 5)
    throw std::runtime_error("Badly formatted string used for --crop.");
  return CropWindow{std::atoi(m.str(1).c_str()), std::atoi(m.str(2).c_str()), std::atoi(m.str(3).c_str()),
                    std::atoi(m.str(4).c_str())};
}

// Fixed-seed synthetic NIF weights (the trained `converted.hdf5` is missing from the reference checkout).
NifWeightsFile syntheticNif(std::uint32_t embedding, std::uint32_t hidden, std::uint64_t seed) {
  std::uint64_t s = seed ? seed : 1;
  auto next = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; };
  auto gauss = [&]() {
    const double u1 = ((next() >> 11) + 1) * (1.0 / 9007199254740993.0), u2 = (next() >> 11) * (1.0 / 9007199254740992.0);
    return std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586 * u2);
  };
  auto toHalf = [](float f) { _Float16 h = (_Float16)f; std::uint16_t b; std::memcpy(&b, &h, 2); return b; };
  NifWeightsFile w;
  w.embedding = embedding;
  const std::uint32_t feat = 4 * embedding;
  std::uint32_t width = feat;
  for (int i = 0; i < 7; ++i) {
    NifWeightsFile::Layer L;
    L.in = width + (i == 3 ? feat : 0);
    L.out = i == 6 ? 3 : hidden;
    L.relu = i == 6 ? 0 : 1;
    const double scale = i == 6 ? 0.5 / std::sqrt((double)L.in) : std::sqrt(2.0 / L.in);
    L.kernel.resize((size_t)L.in * L.out);
    for (auto& k : L.kernel) k = toHalf((float)(gauss() * scale));
    L.bias.resize(L.out);
    for (auto& b : L.bias) b = toHalf((float)(gauss() * (i == 6 ? 0.1 : 0.05)));
    w.layers.push_back(std::move(L));
    width = L.out;
  }
  return w;
}

}  // namespace

int main(int argc, char** argv) {
  Options args;
  try {
    args = parse(argc, argv);
    g_log = logRank(args.logLevel);
  } catch (const std::exception& e) {
    std::fprintf(stderr, "[I] Exiting after: %s.\n", e.what());
    return EXIT_FAILURE;
  }

  try {
    LOG("trace", "HitRecord size: %zu", sizeof(HitRecord));
    LOG("trace", "TraceResult size: %zu", sizeof(TraceResult));
    LOG("trace", "CompactBVH2Node size: %zu", sizeof(BvhNode));

    // ===== Scene setup (buildSceneDescription + buildSceneData) =====
    SceneParts parts;
    if (args.meshFile.empty()) {
      if (args.scene == "box" || args.scene == "box-simple") parts = makeCornellBoxScene(args.builtinMesh, args.scene == "box-simple");
      else if (args.scene == "spheres") parts = makePrimitiveScene();
      else throw std::runtime_error("Invalid scene selection: '" + args.scene + "'");
    } else {
      parts = importScene(args.meshFile, args.loadNormals);
    }
    HostScene scene;
    finaliseScene(parts, scene);
    LOG("debug", "Max leaf depth in BVH: %u", scene.bvhMaxDepth);

    const CropWindow window = parseCrop(args.crop, args.width, args.height);
    LOG("info", "Rendering window: width: %d, height: %d, start col: %d, start row: %d", window.w, window.h, window.c, window.r);
    const bool pathTrace = args.renderMode == "path-trace";
    SceneRef sceneRef;
    sceneRef.data = &scene;
    sceneRef.imageWidth = (float)args.width; sceneRef.imageHeight = (float)args.height;
    sceneRef.fovRadians = scene.horizontalFov; sceneRef.antiAliasScale = args.antiAlias;
    sceneRef.maxPathLength = args.maxPathLength; sceneRef.rouletteStartDepth = args.rouletteStartDepth;
    sceneRef.samplesPerPixel = args.samples; sceneRef.rngSeed = args.seed;
    sceneRef.window = window; sceneRef.pathTrace = pathTrace;

    // ===== Rendering (renderIPU, trace.cpp:270-336) =====
    std::vector<TraceResult> rayStream((size_t)window.w * window.h);
    initPerspectiveRayStream(rayStream.data(), args.width, args.height, window, scene.horizontalFov);

    B200Scene::RayCallbackFn rayCallback;
    B200Scene::RayCallbackFn* rayCallbackPtr = nullptr;
    if (args.rayCallback) {
      rayCallback = [](std::size_t idx, const std::vector<TraceResult>&) { LOG("debug", "Application callback received batch %zu", idx); };
      rayCallbackPtr = &rayCallback;
    }
    B200Scene device(scene.spheres, scene.discs, sceneRef, rayStream, args.raysPerWorker, rayCallbackPtr);
    device.setRuntimeConfig(RuntimeConfig{args.ipus, args.ipus});
    if (!args.nifHdri.empty()) {
      if (args.nifSynthetic) {
        b200rt_nif_metadata md{};
        if (b200rt_read_nif_metadata((args.nifHdri + "/nif_metadata.txt").c_str(), &md) != 0)
          throw std::runtime_error(b200rt_scene_last_error());
        NifWeightsFile w = syntheticNif(md.embedding_dimension, md.hidden_size ? md.hidden_size : 320, args.seed);
        w.max = md.max; std::copy(md.mean, md.mean + 3, w.mean); w.logToneMap = (std::uint32_t)md.log_tone_map;
        device.setNifWeights(w);
        LOG("warn", "Using synthetic NIF weights (seed %llu) shaped by '%s/nif_metadata.txt'", (unsigned long long)args.seed, args.nifHdri.c_str());
      } else {
        device.loadNifModel(args.nifHdri);
      }
    }
    device.setHdriRotation(args.hdriRotation);
    device.setAvailableMemoryProportion(args.availableMemoryProportion);
    device.setMaxNifBatchSize(args.maxNifBatchSize);

    LOG("info", "B200 Rendering started.");
    if (device.run() != EXIT_SUCCESS) return EXIT_FAILURE;
    LOG("info", "B200 Rendering finished.");
    if (!args.saveRayStream.empty()) {
      FILE* f = std::fopen(args.saveRayStream.c_str(), "wb");
      if (!f || std::fwrite(rayStream.data(), sizeof(TraceResult), rayStream.size(), f) != rayStream.size())
        throw std::runtime_error("could not write " + args.saveRayStream);
      std::fclose(f);
    }
    if (pathTrace) b200rt_scale_rgb(rayStream.data(), rayStream.size(), 1.f / (float)args.samples);

    const double secs = device.getTraceTimeSecs();
    const double casts = pathTrace ? (double)args.samples : 1.0;
    LOG("info", "B200 time: %f", secs);
    LOG("info", "B200 %s per second: %g", pathTrace ? "paths" : "rays", (double)rayStream.size() * casts / secs);
    const auto& st = device.getStats();
    LOG("info", "B200 BVH queries per second: %g (kernel time %.3f ms)",
        (double)(st.closest_hit_queries + st.occlusion_queries) / secs, st.kernel_ms);

    static const std::map<std::string, int> vis = {{"rgb", 0}, {"id", 1}, {"normal", 2}, {"tfar", 3}, {"color", 4}, {"hitpoint", 5}};
    std::vector<float> image((size_t)args.width * args.height * 3, 0.f);
    b200rt_host_scene* unused = nullptr;
    (void)unused;
    b200rt_scene_desc d{};
    d.mat_ids = scene.matIDs.data(); d.num_mat_ids = (std::uint32_t)scene.matIDs.size();
    d.materials = scene.materials.data(); d.num_materials = (std::uint32_t)scene.materials.size();
    const long hits = visualiseHits(rayStream.data(), rayStream.size(), d, vis.at(args.visualise), image.data(), args.width, args.height);
    const std::string out = args.outprefix + "_" + args.visualise + "_b200.exr";
    writeExr(out, image.data(), args.width, args.height);
    LOG("debug", "B200 hit count: %ld", hits);
    LOG("info", "Wrote %s", out.c_str());
    LOG("info", "Done.");
    return EXIT_SUCCESS;
  } catch (const std::exception& e) {
    std::fprintf(stderr, "[E] %s\n", e.what());
    return EXIT_FAILURE;
  }
}
