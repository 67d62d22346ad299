// Image writers for the AOV images: uncompressed scanline OpenEXR (fp32 B, G, R channels) and PFM.
// Replaces cv::imwrite("<prefix>_<vis>_<backend>.exr") in the reference (trace.cpp:503-523).
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "scene_build.hpp"

namespace b200rt {

namespace {
struct Out {
  std::vector<unsigned char> d;
  void bytes(const void* p, size_t n) { d.insert(d.end(), (const unsigned char*)p, (const unsigned char*)p + n); }
  void str(const char* s) { bytes(s, std::strlen(s) + 1); }
  void i32(std::int32_t v) { bytes(&v, 4); }
  void u64(std::uint64_t v) { bytes(&v, 8); }
  void f32(float v) { bytes(&v, 4); }
  void u8(unsigned char v) { d.push_back(v); }
  void attr(const char* name, const char* type, std::int32_t size) { str(name); str(type); i32(size); }
};
}  // namespace

void writeExr(const std::string& path, const float* bgr, int w, int h) {
  Out o;
  o.i32(20000630);  // magic
  o.i32(2);         // version 2, single-part scanline
  // channels: alphabetical B, G, R; each: name\0, pixelType(2=FLOAT), pLinear(1)+pad(3), xSampling, ySampling
  o.attr("channels", "chlist", 3 * (2 + 4 + 4 + 4 + 4) + 1);
  for (const char* c : {"B", "G", "R"}) { o.str(c); o.i32(2); o.u8(0); o.u8(0); o.u8(0); o.u8(0); o.i32(1); o.i32(1); }
  o.u8(0);
  o.attr("compression", "compression", 1); o.u8(0);
  o.attr("dataWindow", "box2i", 16); o.i32(0); o.i32(0); o.i32(w - 1); o.i32(h - 1);
  o.attr("displayWindow", "box2i", 16); o.i32(0); o.i32(0); o.i32(w - 1); o.i32(h - 1);
  o.attr("lineOrder", "lineOrder", 1); o.u8(0);
  o.attr("pixelAspectRatio", "float", 4); o.f32(1.f);
  o.attr("screenWindowCenter", "v2f", 8); o.f32(0.f); o.f32(0.f);
  o.attr("screenWindowWidth", "float", 4); o.f32(1.f);
  o.u8(0);  // end of header
  const std::uint64_t rowBytes = 8 + 3ull * 4 * (std::uint64_t)w;
  const std::uint64_t tableEnd = o.d.size() + 8ull * (std::uint64_t)h;
  for (int y = 0; y < h; ++y) o.u64(tableEnd + rowBytes * (std::uint64_t)y);
  std::vector<float> plane((size_t)w);
  for (int y = 0; y < h; ++y) {
    o.i32(y);
    o.i32(3 * 4 * w);
    for (int c = 0; c < 3; ++c) {  // stored B,G,R == image channel order
      for (int x = 0; x < w; ++x) plane[(size_t)x] = bgr[3 * ((size_t)y * w + x) + c];
      o.bytes(plane.data(), 4 * (size_t)w);
    }
  }
  FILE* f = std::fopen(path.c_str(), "wb");
  if (!f) throw std::runtime_error("Could not open '" + path + "' for writing");
  const size_t n = std::fwrite(o.d.data(), 1, o.d.size(), f);
  std::fclose(f);
  if (n != o.d.size()) throw std::runtime_error("Short write to '" + path + "'");
}

void writePfm(const std::string& path, const float* bgr, int w, int h) {
  FILE* f = std::fopen(path.c_str(), "wb");
  if (!f) throw std::runtime_error("Could not open '" + path + "' for writing");
  std::fprintf(f, "PF\n%d %d\n-1.0\n", w, h);
  std::vector<float> row(3 * (size_t)w);
  for (int y = h - 1; y >= 0; --y) {  // PFM stores bottom row first, RGB
    for (int x = 0; x < w; ++x) {
      const float* px = bgr + 3 * ((size_t)y * w + x);
      row[3 * (size_t)x + 0] = px[2];
      row[3 * (size_t)x + 1] = px[1];
      row[3 * (size_t)x + 2] = px[0];
    }
    std::fwrite(row.data(), 4, row.size(), f);
  }
  std::fclose(f);
}

}  // namespace b200rt
