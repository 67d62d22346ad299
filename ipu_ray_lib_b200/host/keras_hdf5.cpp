// Minimal HDF5 reader for the one file layout the trace path needs: the Keras "h5" model the reference's
// IpuScene::loadNifModel opens (`<assets.extra>/converted.hdf5`, src/IpuScene.cpp:177) through Hdf5Model
// (src/keras/Hdf5Model.cpp:8-133). libhdf5 is not part of this build, so the subset of the HDF5 File Format
// Specification (version 3.0) that h5py/Keras emit for such a model is read directly:
//
//   superblock version 0 (and 2/3)                       spec III.A
//   version 1 object headers with continuation blocks    spec IV.A.1.a, message 0x0010
//   old-style groups: symbol table message 0x0011 -> v1 B-tree ('TREE') -> symbol nodes ('SNOD') + local heap ('HEAP');
//     compact new-style groups (link messages 0x0006) are accepted too
//   attributes (message 0x000C, versions 1-3) holding a scalar string: fixed length, or variable length through the
//     global heap ('GCOL')                                spec IV.A.2.m, III.E
//   datasets: simple dataspace (0x0001), IEEE little-endian float16/float32 datatype (0x0003), contiguous or compact
//     layout (0x0008 version 3)                           spec IV.A.2.b/d/i
//
// What is extracted is exactly what Hdf5Model extracts: the root attributes keras_version / backend / model_config, the
// Dense layers of the "Functional" model_config in order (InputLayer and Concatenate are ignored, anything else is an
// error, Hdf5Model.cpp:16-53), and per layer /model_weights/<name>/<name>/kernel:0 [in, out] (+ bias:0 [out] when
// use_bias). float32 weights are rounded to fp16 (RNE) here, as the reference does when it uploads them.
#include <cstdint>
#include <cstring>
#include <fstream>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/b200rt_scene.h"
#include "mini_json.hpp"

namespace {

[[noreturn]] void bad(const std::string& what) { throw std::runtime_error("hdf5: " + what); }

struct Message {
  uint16_t type;
  uint8_t flags;
  size_t off, size;  // payload within the file image
};

class H5File {
 public:
  explicit H5File(const std::string& path) {
    std::ifstream f(path, std::ios::binary);
    if (!f) bad("could not open '" + path + "'");
    buf_.assign(std::istreambuf_iterator<char>(f), std::istreambuf_iterator<char>());
    static const unsigned char sig[8] = {0x89, 'H', 'D', 'F', '\r', '\n', 0x1a, '\n'};
    // the superblock sits at 0, 512, 1024, ... (a user block may precede it)
    size_t at = 0;
    for (;; at = at ? at * 2 : 512) {
      if (at + 8 > buf_.size()) bad("'" + path + "' is not an HDF5 file");
      if (std::memcmp(&buf_[at], sig, 8) == 0) break;
    }
    const uint8_t version = u8(at + 8);
    if (version == 0 || version == 1) {
      so_ = u8(at + 13); sl_ = u8(at + 14);
      size_t p = at + 24 + (version == 1 ? 4 : 0);
      base_ = rdOff(p);
      p += 4 * (size_t)so_;               // base, free-space, end-of-file, driver-info addresses
      rootHeader_ = rdOff(p + so_);       // root symbol table entry: link name offset, then object header address
    } else if (version == 2 || version == 3) {
      so_ = u8(at + 9); sl_ = u8(at + 10);
      base_ = rdOff(at + 12);
      rootHeader_ = rdOff(at + 12 + 3 * (size_t)so_);
    } else {
      bad("unsupported superblock version " + std::to_string(version));
    }
    if ((so_ != 4 && so_ != 8) || (sl_ != 4 && sl_ != 8)) bad("unsupported offset/length size");
  }

  uint64_t root() const { return rootHeader_; }

  // All messages of the version-1 object header at `addr`, continuation blocks followed.
  std::vector<Message> messages(uint64_t addr) const {
    const size_t h = abs(addr);
    need(h, 16);
    if (std::memcmp(&buf_[h], "OHDR", 4) == 0) bad("version 2 object headers are not supported (file written with libver='latest')");
    if (u8(h) != 1) bad("unsupported object header version");
    const uint16_t total = u16(h + 2);
    std::vector<Message> out;
    std::vector<std::pair<size_t, size_t>> blocks = {{h + 16, u32(h + 8)}};
    for (size_t b = 0; b < blocks.size(); ++b) {
      size_t p = blocks[b].first;
      const size_t end = p + blocks[b].second;
      need(p, blocks[b].second);
      while (p + 8 <= end && out.size() < total) {
        Message m{u16(p), u8(p + 4), p + 8, u16(p + 2)};
        need(m.off, m.size);
        if (m.type == 0x0010) blocks.push_back({abs(rdOff(m.off)), (size_t)rdLen(m.off + so_)});
        out.push_back(m);
        p = m.off + m.size;
      }
    }
    return out;
  }

  // Object header address of child `name` of the group whose header is at `group`.
  uint64_t child(uint64_t group, const std::string& name) const {
    for (const Message& m : messages(group)) {
      if (m.type == 0x0011) {  // symbol table: B-tree + local heap
        const uint64_t btree = rdOff(m.off), heap = rdOff(m.off + so_);
        const size_t hp = abs(heap);
        need(hp, 8 + 2 * (size_t)sl_ + so_);
        if (std::memcmp(&buf_[hp], "HEAP", 4) != 0) bad("bad local heap signature");
        const size_t names = abs(rdOff(hp + 8 + 2 * (size_t)sl_));
        uint64_t found = 0;
        if (searchTree(btree, names, name, found)) return found;
      } else if (m.type == 0x0006) {  // link message (compact new-style group)
        size_t p = m.off;
        const uint8_t ver = u8(p), flags = u8(p + 1);
        if (ver != 1) continue;
        p += 2;
        uint8_t linkType = 0;
        if (flags & 0x08) linkType = u8(p++);
        if (flags & 0x04) p += 8;
        if (flags & 0x10) p += 1;
        const int lenSize = 1 << (flags & 3);
        uint64_t len = 0;
        for (int i = 0; i < lenSize; ++i) len |= (uint64_t)u8(p + i) << (8 * i);
        p += lenSize;
        need(p, len);
        const std::string linkName((const char*)&buf_[p], (size_t)len);
        p += len;
        if (linkType == 0 && linkName == name) return rdOff(p);
      }
    }
    bad("no object named '" + name + "'");
  }

  uint64_t resolve(const std::string& path) const {
    uint64_t at = root();
    size_t i = 0;
    while (i < path.size()) {
      while (i < path.size() && path[i] == '/') ++i;
      size_t j = path.find('/', i);
      if (j == std::string::npos) j = path.size();
      if (j > i) at = child(at, path.substr(i, j - i));
      i = j;
    }
    return at;
  }

  // Scalar string attribute `name` of the object at `addr`.
  std::string stringAttribute(uint64_t addr, const std::string& name) const {
    for (const Message& m : messages(addr)) {
      if (m.type != 0x000C) continue;
      size_t p = m.off;
      const uint8_t ver = u8(p);
      if (ver < 1 || ver > 3) bad("unsupported attribute message version");
      const uint16_t nameSize = u16(p + 2), typeSize = u16(p + 4), spaceSize = u16(p + 6);
      p += 8 + (ver == 3 ? 1 : 0);
      auto pad = [&](size_t n) { return ver == 1 ? (n + 7) / 8 * 8 : n; };
      need(p, nameSize);
      const std::string attrName((const char*)&buf_[p], strnlen((const char*)&buf_[p], nameSize));
      p += pad(nameSize);
      const size_t typeAt = p;
      p += pad(typeSize) + pad(spaceSize);
      if (attrName != name) continue;
      const uint32_t cls = u8(typeAt) & 0x0f, size = u32(typeAt + 4);
      if (cls == 3) {  // fixed-length string
        need(p, size);
        return std::string((const char*)&buf_[p], strnlen((const char*)&buf_[p], size));
      }
      if (cls == 9) {  // variable length: {length u32, global heap collection address, object index u32}
        need(p, 4 + (size_t)so_ + 4);
        return globalHeapObject(rdOff(p + 4), u32(p + 4 + so_));
      }
      bad("attribute '" + name + "' is not a string");
    }
    bad("no attribute named '" + name + "'");
  }

  struct Dataset {
    std::vector<uint64_t> dims;
    uint32_t elemSize = 0;
    const uint8_t* data = nullptr;
    size_t bytes = 0;
  };
  Dataset dataset(uint64_t addr) const {
    Dataset d;
    bool haveLayout = false;
    for (const Message& m : messages(addr)) {
      if (m.type == 0x0001) {  // dataspace
        const uint8_t ver = u8(m.off), rank = u8(m.off + 1);
        size_t p = m.off + (ver == 1 ? 8 : 4);
        if (ver != 1 && ver != 2) bad("unsupported dataspace version");
        for (int i = 0; i < rank; ++i) d.dims.push_back(rdLen(p + (size_t)i * sl_));
      } else if (m.type == 0x0003) {  // datatype
        const uint8_t cls = u8(m.off) & 0x0f, bits0 = u8(m.off + 1);
        d.elemSize = u32(m.off + 4);
        if (cls != 1) bad("dataset is not floating point");
        if (bits0 & 1) bad("big-endian floats are not supported");
        if (d.elemSize != 2 && d.elemSize != 4) bad("Only float32 and float16 weights are supported.");  // Hdf5Model.cpp:118-120
      } else if (m.type == 0x0008) {  // data layout
        const uint8_t ver = u8(m.off), cls = u8(m.off + 1);
        if (ver != 3) bad("unsupported data layout message version");
        if (cls == 1) {  // contiguous
          const uint64_t at = rdOff(m.off + 2);
          d.bytes = (size_t)rdLen(m.off + 2 + so_);
          if (d.bytes) { need(abs(at), d.bytes); d.data = &buf_[abs(at)]; }
        } else if (cls == 0) {  // compact
          d.bytes = u16(m.off + 2);
          need(m.off + 4, d.bytes);
          d.data = &buf_[m.off + 4];
        } else {
          bad("chunked datasets are not supported (save the model without compression)");
        }
        haveLayout = true;
      }
    }
    uint64_t n = 1;
    for (uint64_t x : d.dims) n *= x;
    if (!haveLayout || !d.elemSize || d.bytes != n * d.elemSize) bad("dataset header is incomplete or its size does not match its shape");
    return d;
  }

 private:
  std::vector<uint8_t> buf_;
  uint8_t so_ = 8, sl_ = 8;
  uint64_t base_ = 0, rootHeader_ = 0;

  void need(size_t off, size_t n) const { if (off > buf_.size() || n > buf_.size() - off) bad("truncated file"); }
  size_t abs(uint64_t addr) const {
    if (addr == ~0ull || (so_ == 4 && addr == 0xffffffffull)) bad("undefined address");
    return (size_t)(base_ + addr);
  }
  uint8_t u8(size_t p) const { need(p, 1); return buf_[p]; }
  uint16_t u16(size_t p) const { need(p, 2); return (uint16_t)(buf_[p] | buf_[p + 1] << 8); }
  uint32_t u32(size_t p) const { need(p, 4); uint32_t v; std::memcpy(&v, &buf_[p], 4); return v; }
  uint64_t rdN(size_t p, int n) const { need(p, (size_t)n); uint64_t v = 0; std::memcpy(&v, &buf_[p], (size_t)n); return v; }
  uint64_t rdOff(size_t p) const { return rdN(p, so_); }
  uint64_t rdLen(size_t p) const { return rdN(p, sl_); }

  bool searchTree(uint64_t node, size_t names, const std::string& name, uint64_t& found) const {
    const size_t p = abs(node);
    need(p, 8 + 2 * (size_t)so_);
    if (std::memcmp(&buf_[p], "TREE", 4) == 0) {
      if (u8(p + 4) != 0) bad("not a group B-tree");
      const uint16_t used = u16(p + 6);
      size_t q = p + 8 + 2 * (size_t)so_ + sl_;  // skip key 0
      for (uint16_t i = 0; i < used; ++i, q += (size_t)so_ + sl_)
        if (searchTree(rdOff(q), names, name, found)) return true;
      return false;
    }
    if (std::memcmp(&buf_[p], "SNOD", 4) != 0) bad("bad symbol table node signature");
    const uint16_t n = u16(p + 6);
    size_t q = p + 8;
    for (uint16_t i = 0; i < n; ++i, q += 2 * (size_t)so_ + 24) {
      const size_t nm = names + (size_t)rdOff(q);
      need(nm, 1);
      const size_t len = strnlen((const char*)&buf_[nm], buf_.size() - nm);
      if (name.size() == len && std::memcmp(&buf_[nm], name.data(), len) == 0) { found = rdOff(q + so_); return true; }
    }
    return false;
  }

  std::string globalHeapObject(uint64_t collection, uint32_t index) const {
    const size_t p = abs(collection);
    need(p, 8 + (size_t)sl_);
    if (std::memcmp(&buf_[p], "GCOL", 4) != 0) bad("bad global heap signature");
    const size_t end = p + (size_t)rdLen(p + 8);
    size_t q = p + 8 + sl_;
    while (q + 8 + sl_ <= end) {
      const uint16_t idx = u16(q);
      const uint64_t size = rdLen(q + 8);
      if (idx == 0) break;
      if (idx == index) { need(q + 8 + sl_, size); return std::string((const char*)&buf_[q + 8 + sl_], (size_t)size); }
      q += 8 + sl_ + (size_t)((size + 7) / 8 * 8);
    }
    bad("global heap object not found");
  }
};

uint16_t float_to_half(float f) {
  const _Float16 h = (_Float16)f;  // round to nearest even
  uint16_t b;
  std::memcpy(&b, &h, 2);
  return b;
}

thread_local std::string g_h5Error;

}  // namespace

struct b200rt_keras_model {
  struct Layer {
    std::string name, activation, dtype;
    uint32_t in = 0, out = 0;
    bool useBias = false;
    std::vector<uint16_t> kernel, bias;
  };
  std::vector<Layer> layers;
  std::string kerasVersion, backend;
};

extern "C" {

const char* b200rt_keras_last_error(void) { return g_h5Error.c_str(); }

int b200rt_keras_hdf5_open(const char* path, b200rt_keras_model** out) {
  if (!path || !out) { g_h5Error = "null argument"; return B200RT_ERR_INVALID_ARG; }
  *out = nullptr;
  try {
    const H5File f(path);
    auto model = std::make_unique<b200rt_keras_model>();
    model->kerasVersion = f.stringAttribute(f.root(), "keras_version");
    model->backend = f.stringAttribute(f.root(), "backend");
    const mini_json::Value cfg = mini_json::parse(f.stringAttribute(f.root(), "model_config"));
    if (cfg.at("class_name").str != "Functional") bad("Expected a Keras 'Functional' Model");  // Hdf5Model.cpp:17-20
    for (const mini_json::Value& l : cfg.at("config").at("layers").arr) {
      const std::string cn = l.at("class_name").str;
      if (cn == "Dense") {
        const mini_json::Value& c = l.at("config");
        b200rt_keras_model::Layer L;
        L.name = c.at("name").str; L.activation = c.at("activation").str; L.dtype = c.at("dtype").str;
        L.out = (uint32_t)c.at("units").num; L.useBias = c.at("use_bias").b;
        model->layers.push_back(std::move(L));
      } else if (cn != "InputLayer" && cn != "Concatenate") {
        bad("Layer class: '" + cn + "' not supported by Hdf5Model loader.");  // Hdf5Model.cpp:44-50
      }
    }
    auto to_half = [](const H5File::Dataset& d, std::vector<uint16_t>& dst) {
      const size_t n = d.bytes / d.elemSize;
      dst.resize(n);
      if (d.elemSize == 2) { std::memcpy(dst.data(), d.data, n * 2); return; }
      for (size_t i = 0; i < n; ++i) { float v; std::memcpy(&v, d.data + 4 * i, 4); dst[i] = float_to_half(v); }
    };
    for (auto& L : model->layers) {
      const std::string base = "/model_weights/" + L.name + "/" + L.name + "/";  // Hdf5Model.cpp:71-82
      const H5File::Dataset k = f.dataset(f.resolve(base + "kernel:0"));
      if (k.dims.size() != 2 || k.dims[1] != L.out) bad("kernel of '" + L.name + "' is not [in, units]");
      L.in = (uint32_t)k.dims[0];
      to_half(k, L.kernel);
      if (L.useBias) {
        const H5File::Dataset b = f.dataset(f.resolve(base + "bias:0"));
        if (b.dims.size() != 1 || b.dims[0] != L.out) bad("bias of '" + L.name + "' is not [units]");
        to_half(b, L.bias);
      }
    }
    *out = model.release();
    return B200RT_OK;
  } catch (const std::exception& e) {
    g_h5Error = e.what();
    return B200RT_ERR_IO;
  }
}

void b200rt_keras_hdf5_close(b200rt_keras_model* m) { delete m; }

uint32_t b200rt_keras_hdf5_num_layers(const b200rt_keras_model* m) { return m ? (uint32_t)m->layers.size() : 0u; }

int b200rt_keras_hdf5_layer(const b200rt_keras_model* m, uint32_t i, b200rt_keras_layer* out) {
  if (!m || !out || i >= m->layers.size()) { g_h5Error = "bad layer index"; return B200RT_ERR_INVALID_ARG; }
  const auto& L = m->layers[i];
  *out = b200rt_keras_layer{};
  std::strncpy(out->name, L.name.c_str(), sizeof(out->name) - 1);
  std::strncpy(out->activation, L.activation.c_str(), sizeof(out->activation) - 1);
  out->layer.in_features = L.in;
  out->layer.out_features = L.out;
  out->layer.kernel_f16 = L.kernel.data();
  out->layer.bias_f16 = L.useBias ? L.bias.data() : nullptr;
  out->layer.relu = L.activation == "relu" ? 1 : 0;
  return B200RT_OK;
}

const char* b200rt_keras_hdf5_version(const b200rt_keras_model* m) { return m ? m->kerasVersion.c_str() : ""; }

}  // extern "C"
