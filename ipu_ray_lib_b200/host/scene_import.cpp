// importScene placeholder; the COLLADA/GLB-with-camera importer lands in scene_import.cpp (N2).
#include <stdexcept>
#include "scene_build.hpp"
namespace b200rt {
SceneParts importScene(const std::string& file, bool) {
  throw std::runtime_error("importScene: no importer for '" + file + "' yet");
}
}  // namespace b200rt
