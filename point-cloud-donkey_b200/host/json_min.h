// json_min.h — the small subset of jsoncpp the reference's JSONObject uses (utils/json_object.cpp:41-178,
// utils/json_parameter.h:26-33): parse (with // and /* */ comments, as jsoncpp's default reader allows; on duplicate
// keys the last one wins), typed access with defaults, and a styled writer.  No dependencies.
#pragma once
#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace jsonmin {

struct Value {
  enum Type { Null, Bool, Number, String, Array, Object } type = Null;
  bool b = false;
  double num = 0;
  bool is_int = false;
  std::string str;
  std::vector<Value> arr;
  std::vector<std::pair<std::string, Value>> obj;  // insertion order kept; lookups are linear (configs are small)

  Value() = default;
  static Value object() { Value v; v.type = Object; return v; }
  static Value array() { Value v; v.type = Array; return v; }
  static Value of(bool x) { Value v; v.type = Bool; v.b = x; return v; }
  static Value of(int x) { Value v; v.type = Number; v.num = x; v.is_int = true; return v; }
  static Value of(double x) { Value v; v.type = Number; v.num = x; return v; }
  static Value of(const std::string& x) { Value v; v.type = String; v.str = x; return v; }

  bool isNull() const { return type == Null; }
  bool isObject() const { return type == Object; }
  bool isString() const { return type == String; }
  bool isNumber() const { return type == Number; }
  bool isBool() const { return type == Bool; }
  bool isMember(const std::string& k) const { return find(k) != nullptr; }
  const Value* find(const std::string& k) const {
    const Value* hit = nullptr;
    if (type == Object)
      for (auto& kv : obj)
        if (kv.first == k) hit = &kv.second;  // last duplicate wins
    return hit;
  }
  const Value& operator[](const std::string& k) const {
    static const Value null_value;
    const Value* v = find(k);
    return v ? *v : null_value;
  }
  Value& set(const std::string& k, const Value& v) {
    for (auto& kv : obj)
      if (kv.first == k) { kv.second = v; return kv.second; }
    type = Object;
    obj.push_back({k, v});
    return obj.back().second;
  }
};

class Parser {
 public:
  explicit Parser(const std::string& s) : s_(s) {}
  Value parse() {
    Value v = value();
    skip();
    if (p_ != s_.size()) fail("trailing characters");
    return v;
  }

 private:
  const std::string& s_;
  size_t p_ = 0;
  [[noreturn]] void fail(const std::string& m) const {
    throw std::runtime_error("JSON parse error at offset " + std::to_string(p_) + ": " + m);
  }
  void skip() {
    for (;;) {
      while (p_ < s_.size() && std::isspace((unsigned char)s_[p_])) ++p_;
      if (p_ + 1 < s_.size() && s_[p_] == '/' && s_[p_ + 1] == '/') {
        while (p_ < s_.size() && s_[p_] != '\n') ++p_;
      } else if (p_ + 1 < s_.size() && s_[p_] == '/' && s_[p_ + 1] == '*') {
        p_ += 2;
        while (p_ + 1 < s_.size() && !(s_[p_] == '*' && s_[p_ + 1] == '/')) ++p_;
        p_ = std::min(s_.size(), p_ + 2);
      } else
        return;
    }
  }
  Value value() {
    skip();
    if (p_ >= s_.size()) fail("unexpected end");
    char c = s_[p_];
    if (c == '{') return object();
    if (c == '[') return array();
    if (c == '"') return Value::of(string());
    if (s_.compare(p_, 4, "true") == 0) { p_ += 4; return Value::of(true); }
    if (s_.compare(p_, 5, "false") == 0) { p_ += 5; return Value::of(false); }
    if (s_.compare(p_, 4, "null") == 0) { p_ += 4; return Value(); }
    return number();
  }
  Value object() {
    Value v = Value::object();
    ++p_;
    skip();
    if (p_ < s_.size() && s_[p_] == '}') { ++p_; return v; }
    for (;;) {
      skip();
      if (p_ >= s_.size() || s_[p_] != '"') fail("expected a key");
      std::string k = string();
      skip();
      if (p_ >= s_.size() || s_[p_] != ':') fail("expected ':'");
      ++p_;
      Value x = value();
      v.obj.push_back({k, x});
      skip();
      if (p_ < s_.size() && s_[p_] == ',') { ++p_; continue; }
      if (p_ < s_.size() && s_[p_] == '}') { ++p_; return v; }
      fail("expected ',' or '}'");
    }
  }
  Value array() {
    Value v = Value::array();
    ++p_;
    skip();
    if (p_ < s_.size() && s_[p_] == ']') { ++p_; return v; }
    for (;;) {
      v.arr.push_back(value());
      skip();
      if (p_ < s_.size() && s_[p_] == ',') { ++p_; continue; }
      if (p_ < s_.size() && s_[p_] == ']') { ++p_; return v; }
      fail("expected ',' or ']'");
    }
  }
  std::string string() {
    std::string out;
    ++p_;
    while (p_ < s_.size() && s_[p_] != '"') {
      char c = s_[p_++];
      if (c == '\\' && p_ < s_.size()) {
        char e = s_[p_++];
        switch (e) {
          case 'n': out += '\n'; break;
          case 't': out += '\t'; break;
          case 'r': out += '\r'; break;
          case 'b': out += '\b'; break;
          case 'f': out += '\f'; break;
          case 'u': {
            unsigned cp = (unsigned)std::strtoul(s_.substr(p_, 4).c_str(), nullptr, 16);
            p_ += 4;
            if (cp < 0x80) out += (char)cp;
            else if (cp < 0x800) { out += (char)(0xC0 | (cp >> 6)); out += (char)(0x80 | (cp & 0x3F)); }
            else { out += (char)(0xE0 | (cp >> 12)); out += (char)(0x80 | ((cp >> 6) & 0x3F)); out += (char)(0x80 | (cp & 0x3F)); }
            break;
          }
          default: out += e;
        }
      } else
        out += c;
    }
    if (p_ >= s_.size()) fail("unterminated string");
    ++p_;
    return out;
  }
  Value number() {
    size_t st = p_;
    bool is_int = true;
    if (p_ < s_.size() && (s_[p_] == '-' || s_[p_] == '+')) ++p_;
    while (p_ < s_.size() && (std::isdigit((unsigned char)s_[p_]) || s_[p_] == '.' || s_[p_] == 'e' || s_[p_] == 'E' ||
                              s_[p_] == '-' || s_[p_] == '+')) {
      if (!std::isdigit((unsigned char)s_[p_])) is_int = false;
      ++p_;
    }
    if (st == p_) fail("unexpected character");
    Value v = Value::of(std::strtod(s_.substr(st, p_ - st).c_str(), nullptr));
    v.is_int = is_int;
    return v;
  }
};

inline Value parse_file(const std::string& path) {
  std::ifstream f(path);
  if (!f) return Value();
  std::stringstream ss;
  ss << f.rdbuf();
  std::string s = ss.str();
  return Parser(s).parse();
}

inline void write_value(std::ostream& os, const Value& v, int indent) {
  auto pad = [&](int n) { for (int i = 0; i < n; ++i) os << "   "; };
  switch (v.type) {
    case Value::Null: os << "null"; break;
    case Value::Bool: os << (v.b ? "true" : "false"); break;
    case Value::Number:
      if (v.is_int && std::fabs(v.num) < 9e15) os << (long long)v.num;
      else { char buf[64]; std::snprintf(buf, sizeof(buf), "%.17g", v.num); os << buf; }
      break;
    case Value::String: {
      os << '"';
      for (char c : v.str) {
        if (c == '"' || c == '\\') os << '\\' << c;
        else if (c == '\n') os << "\\n";
        else if (c == '\t') os << "\\t";
        else os << c;
      }
      os << '"';
      break;
    }
    case Value::Array:
      os << "[";
      for (size_t i = 0; i < v.arr.size(); ++i) { if (i) os << ", "; write_value(os, v.arr[i], indent + 1); }
      os << "]";
      break;
    case Value::Object:
      if (v.obj.empty()) { os << "{}"; break; }
      os << "{\n";
      for (size_t i = 0; i < v.obj.size(); ++i) {
        pad(indent + 1);
        os << '"' << v.obj[i].first << "\" : ";
        write_value(os, v.obj[i].second, indent + 1);
        os << (i + 1 < v.obj.size() ? ",\n" : "\n");
      }
      pad(indent);
      os << "}";
      break;
  }
}

inline bool write_file(const Value& v, const std::string& path) {
  std::ofstream f(path);
  if (!f) return false;
  write_value(f, v, 0);
  f << "\n";
  return (bool)f;
}

}  // namespace jsonmin
