// Minimal JSON reader (objects keep insertion order). Enough for glTF's JSON chunk and the NIF
// metadata file; replaces boost::property_tree in the reference (src/neural_networks/NifMetaData.cpp:11-16).
#pragma once
#include <cstdlib>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace mini_json {

struct Value {
  enum Kind { Null, Bool, Number, String, Array, Object } kind = Null;
  bool b = false;
  double num = 0.0;
  std::string str;
  std::vector<Value> arr;
  std::vector<std::pair<std::string, Value>> obj;

  const Value* find(const std::string& key) const {
    for (auto& kv : obj)
      if (kv.first == key) return &kv.second;
    return nullptr;
  }
  const Value& at(const std::string& key) const {
    const Value* v = find(key);
    if (!v) throw std::runtime_error("json: missing key '" + key + "'");
    return *v;
  }
  const Value& at(size_t i) const {
    if (i >= arr.size()) throw std::runtime_error("json: index out of range");
    return arr[i];
  }
  bool has(const std::string& key) const { return find(key) != nullptr; }
  double number(double dflt) const { return kind == Number ? num : dflt; }
  size_t size() const { return kind == Array ? arr.size() : obj.size(); }
};

class Parser {
 public:
  explicit Parser(const std::string& text) : s(text), i(0) {}
  Value parse() {
    Value v = value();
    ws();
    return v;
  }

 private:
  const std::string& s;
  size_t i;
  [[noreturn]] void fail(const char* what) const {
    throw std::runtime_error(std::string("json: ") + what + " at offset " + std::to_string(i));
  }
  void ws() {
    while (i < s.size() && (s[i] == ' ' || s[i] == '\n' || s[i] == '\t' || s[i] == '\r')) ++i;
  }
  bool eat(char c) {
    ws();
    if (i < s.size() && s[i] == c) { ++i; return true; }
    return false;
  }
  Value value() {
    ws();
    if (i >= s.size()) fail("unexpected end");
    Value v;
    const char c = s[i];
    if (c == '{') {
      ++i;
      v.kind = Value::Object;
      if (eat('}')) return v;
      do {
        ws();
        std::string k = string();
        if (!eat(':')) fail("expected ':'");
        v.obj.emplace_back(std::move(k), value());
      } while (eat(','));
      if (!eat('}')) fail("expected '}'");
    } else if (c == '[') {
      ++i;
      v.kind = Value::Array;
      if (eat(']')) return v;
      do { v.arr.push_back(value()); } while (eat(','));
      if (!eat(']')) fail("expected ']'");
    } else if (c == '"') {
      v.kind = Value::String;
      v.str = string();
    } else if (s.compare(i, 4, "true") == 0) {
      v.kind = Value::Bool; v.b = true; i += 4;
    } else if (s.compare(i, 5, "false") == 0) {
      v.kind = Value::Bool; v.b = false; i += 5;
    } else if (s.compare(i, 4, "null") == 0) {
      i += 4;
    } else {
      char* end = nullptr;
      v.num = std::strtod(s.c_str() + i, &end);
      if (end == s.c_str() + i) fail("bad token");
      v.kind = Value::Number;
      i = (size_t)(end - s.c_str());
    }
    return v;
  }
  std::string string() {
    if (i >= s.size() || s[i] != '"') fail("expected string");
    ++i;
    std::string out;
    while (i < s.size() && s[i] != '"') {
      char c = s[i++];
      if (c == '\\' && i < s.size()) {
        char e = s[i++];
        switch (e) {
          case 'n': out += '\n'; break;
          case 't': out += '\t'; break;
          case 'r': out += '\r'; break;
          case 'b': out += '\b'; break;
          case 'f': out += '\f'; break;
          case 'u': {  // keep BMP code points as UTF-8
            unsigned cp = (unsigned)std::strtoul(s.substr(i, 4).c_str(), nullptr, 16);
            i += 4;
            if (cp < 0x80) out += (char)cp;
            else if (cp < 0x800) { out += (char)(0xC0 | (cp >> 6)); out += (char)(0x80 | (cp & 0x3F)); }
            else { out += (char)(0xE0 | (cp >> 12)); out += (char)(0x80 | ((cp >> 6) & 0x3F)); out += (char)(0x80 | (cp & 0x3F)); }
            break;
          }
          default: out += e;
        }
      } else {
        out += c;
      }
    }
    if (i >= s.size()) fail("unterminated string");
    ++i;
    return out;
  }
};

inline Value parse(const std::string& text) { return Parser(text).parse(); }

}  // namespace mini_json
