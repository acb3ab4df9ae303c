// io_formats.h — on-disk formats either side of the hot path:
//   * boost::archive::binary_oarchive / binary_iarchive framing for the .ismd model data
//     (utils/json_object.cpp:84-86,156-161; written for archive library version 17 = Boost 1.71 of the reference's
//     Ubuntu 20.04; the reader accepts versions >= 8, which share the layout).  "Format parity unpinned": the reference
//     ships no .ismd file to compare with.
//   * PCD point clouds (ascii and binary; implicit_shape_model.cpp:213-249 uses pcl::io::loadPCDFile into
//     PointXYZRGBNormal).  binary_compressed and PLY are "next" rows.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <fstream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace ism3d {
namespace io {

// ---- boost binary archive -----------------------------------------------------------------------------------------
class BinaryOArchive {
 public:
  explicit BinaryOArchive(std::ostream& os) : os_(os) {
    const std::string sig = "serialization::archive";
    put<uint64_t>(sig.size());
    os_.write(sig.data(), (std::streamsize)sig.size());
    put<uint16_t>(17);          // library_version_type
    put<uint8_t>(sizeof(int));  // basic_binary_oprimitive::init: native sizes + endianness probe
    put<uint8_t>(sizeof(long));
    put<uint8_t>(sizeof(float));
    put<uint8_t>(sizeof(double));
    put<int32_t>(1);
  }
  template <typename T>
  void put(T v) { os_.write(reinterpret_cast<const char*>(&v), sizeof(T)); }
  void put_string(const std::string& s) {
    put<uint64_t>(s.size());
    os_.write(s.data(), (std::streamsize)s.size());
  }
  // std::vector<arithmetic>: class info (tracking byte + 4-byte version) on the first object of each element type,
  // then collection_size_type (8 bytes) and the raw array (array optimisation of binary archives)
  void put_vector(const std::vector<float>& v) { vec(v, seen_f_); }
  void put_vector(const std::vector<uint32_t>& v) { vec(v, seen_u_); }

 private:
  template <typename T>
  void vec(const std::vector<T>& v, bool& seen) {
    if (!seen) {
      put<uint8_t>(0);
      put<uint32_t>(0);
      seen = true;
    }
    put<uint64_t>(v.size());
    if (!v.empty()) os_.write(reinterpret_cast<const char*>(v.data()), (std::streamsize)(sizeof(T) * v.size()));
  }
  std::ostream& os_;
  bool seen_f_ = false, seen_u_ = false;
};

class BinaryIArchive {
 public:
  explicit BinaryIArchive(std::istream& is) : is_(is) {
    uint64_t n = get<uint64_t>();
    if (n != 22) throw std::runtime_error("not a boost binary archive (signature length)");
    std::string sig(n, '\0');
    is_.read(&sig[0], (std::streamsize)n);
    if (sig != "serialization::archive") throw std::runtime_error("not a boost binary archive (signature)");
    version_ = get<uint16_t>();
    if (version_ < 8) throw std::runtime_error("boost archive version < 8 is not supported");
    uint8_t si = get<uint8_t>(), sl = get<uint8_t>(), sf = get<uint8_t>(), sd = get<uint8_t>();
    if (si != 4 || sl != 8 || sf != 4 || sd != 8) throw std::runtime_error("archive written on an incompatible platform");
    if (get<int32_t>() != 1) throw std::runtime_error("archive endianness mismatch");
  }
  template <typename T>
  T get() {
    T v;
    is_.read(reinterpret_cast<char*>(&v), sizeof(T));
    if (!is_) throw std::runtime_error("unexpected end of .ismd archive");
    return v;
  }
  std::string get_string() {
    uint64_t n = get<uint64_t>();
    if (n > (1ull << 30)) throw std::runtime_error("corrupt string length in archive");
    std::string s(n, '\0');
    if (n) is_.read(&s[0], (std::streamsize)n);
    return s;
  }
  std::vector<float> get_vector_f() { return vec<float>(seen_f_); }
  std::vector<uint32_t> get_vector_u() { return vec<uint32_t>(seen_u_); }
  int version() const { return version_; }

 private:
  template <typename T>
  std::vector<T> vec(bool& seen) {
    if (!seen) {
      get<uint8_t>();
      get<uint32_t>();
      seen = true;
    }
    uint64_t n = get<uint64_t>();
    if (n > (1ull << 32)) throw std::runtime_error("corrupt vector length in archive");
    std::vector<T> v(n);
    if (n) is_.read(reinterpret_cast<char*>(v.data()), (std::streamsize)(sizeof(T) * n));
    if (!is_) throw std::runtime_error("unexpected end of .ismd archive");
    return v;
  }
  std::istream& is_;
  int version_ = 0;
  bool seen_f_ = false, seen_u_ = false;
};

// ---- PCD ----------------------------------------------------------------------------------------------------------
struct Cloud {  // pcl::PointCloud<pcl::PointXYZRGBNormal> as flat arrays
  std::vector<float> xyz, normals;
  std::vector<uint32_t> rgb;
  bool has_normals = false, has_rgb = false;
  size_t size() const { return rgb.size(); }
};

inline bool load_pcd(const std::string& path, Cloud& out, std::string& err) {
  std::ifstream f(path, std::ios::binary);
  if (!f) { err = "could not open " + path; return false; }
  std::vector<std::string> fields;
  std::vector<int> sizes, counts;
  std::vector<char> types;
  size_t points = 0, width = 0, height = 1;
  std::string data_mode, line;
  while (std::getline(f, line)) {
    if (!line.empty() && line.back() == '\r') line.pop_back();
    if (line.empty() || line[0] == '#') continue;
    std::istringstream ls(line);
    std::string key;
    ls >> key;
    if (key == "FIELDS") { std::string s; while (ls >> s) fields.push_back(s); }
    else if (key == "SIZE") { int s; while (ls >> s) sizes.push_back(s); }
    else if (key == "TYPE") { char c; while (ls >> c) types.push_back(c); }
    else if (key == "COUNT") { int c; while (ls >> c) counts.push_back(c); }
    else if (key == "WIDTH") ls >> width;
    else if (key == "HEIGHT") ls >> height;
    else if (key == "POINTS") ls >> points;
    else if (key == "DATA") { ls >> data_mode; break; }
  }
  if (fields.empty() || data_mode.empty()) { err = "malformed PCD header in " + path; return false; }
  if (counts.empty()) counts.assign(fields.size(), 1);
  if (sizes.size() != fields.size() || types.size() != fields.size() || counts.size() != fields.size()) {
    err = "inconsistent PCD header in " + path;
    return false;
  }
  if (points == 0) points = width * height;
  auto idx_of = [&](const char* n) { for (size_t i = 0; i < fields.size(); ++i) if (fields[i] == n) return (int)i; return -1; };
  const int ix = idx_of("x"), iy = idx_of("y"), iz = idx_of("z");
  int irgb = idx_of("rgb");
  if (irgb < 0) irgb = idx_of("rgba");
  const int inx = idx_of("normal_x"), iny = idx_of("normal_y"), inz = idx_of("normal_z");
  if (ix < 0 || iy < 0 || iz < 0) { err = "PCD file has no x/y/z fields: " + path; return false; }
  out.has_normals = inx >= 0 && iny >= 0 && inz >= 0;
  out.has_rgb = irgb >= 0;
  out.xyz.assign(points * 3, 0.f);
  out.normals.assign(points * 3, 0.f);
  out.rgb.assign(points, 0u);
  std::vector<size_t> offs(fields.size());
  size_t stride = 0;
  for (size_t i = 0; i < fields.size(); ++i) { offs[i] = stride; stride += (size_t)sizes[i] * counts[i]; }
  auto store = [&](size_t p, int fi, double v, uint32_t bits) {
    if (fi == ix) out.xyz[3 * p] = (float)v;
    else if (fi == iy) out.xyz[3 * p + 1] = (float)v;
    else if (fi == iz) out.xyz[3 * p + 2] = (float)v;
    else if (fi == inx) out.normals[3 * p] = (float)v;
    else if (fi == iny) out.normals[3 * p + 1] = (float)v;
    else if (fi == inz) out.normals[3 * p + 2] = (float)v;
    else if (fi == irgb) out.rgb[p] = bits & 0x00ffffffu;
  };
  if (data_mode == "ascii") {
    for (size_t p = 0; p < points; ++p) {
      if (!std::getline(f, line)) { err = "PCD file truncated: " + path; return false; }
      std::istringstream ls(line);
      for (size_t fi = 0; fi < fields.size(); ++fi)
        for (int c = 0; c < counts[fi]; ++c) {
          std::string tok;
          if (!(ls >> tok)) { err = "PCD row too short: " + path; return false; }
          if (c > 0) continue;
          if ((int)fi == irgb) {
            uint32_t bits;
            if (types[fi] == 'F') { float fv = std::strtof(tok.c_str(), nullptr); std::memcpy(&bits, &fv, 4); }
            else bits = (uint32_t)std::strtoul(tok.c_str(), nullptr, 10);
            store(p, (int)fi, 0, bits);
          } else
            store(p, (int)fi, tok == "nan" ? NAN : std::strtod(tok.c_str(), nullptr), 0);
        }
    }
  } else if (data_mode == "binary") {
    std::vector<char> buf(stride * points);
    f.read(buf.data(), (std::streamsize)buf.size());
    if ((size_t)f.gcount() != buf.size()) { err = "PCD binary payload truncated: " + path; return false; }
    for (size_t p = 0; p < points; ++p)
      for (size_t fi = 0; fi < fields.size(); ++fi) {
        const char* src = buf.data() + p * stride + offs[fi];
        if ((int)fi == irgb) { uint32_t bits = 0; std::memcpy(&bits, src, std::min(4, sizes[fi])); store(p, (int)fi, 0, bits); continue; }
        double v = 0;
        if (types[fi] == 'F' && sizes[fi] == 4) { float t; std::memcpy(&t, src, 4); v = t; }
        else if (types[fi] == 'F' && sizes[fi] == 8) { std::memcpy(&v, src, 8); }
        else if (sizes[fi] == 4) { int32_t t; std::memcpy(&t, src, 4); v = types[fi] == 'U' ? (double)(uint32_t)t : (double)t; }
        else if (sizes[fi] == 2) { int16_t t; std::memcpy(&t, src, 2); v = types[fi] == 'U' ? (double)(uint16_t)t : (double)t; }
        else if (sizes[fi] == 1) { int8_t t; std::memcpy(&t, src, 1); v = types[fi] == 'U' ? (double)(uint8_t)t : (double)t; }
        store(p, (int)fi, v, 0);
      }
  } else {
    err = "PCD DATA mode '" + data_mode + "' is not supported (binary_compressed is a SURVEY 8f-2 'next' row): " + path;
    return false;
  }
  return true;
}

inline bool save_pcd_binary(const std::string& path, const float* xyz, const float* normals, const uint32_t* rgb, size_t n) {
  std::ofstream f(path, std::ios::binary);
  if (!f) return false;
  f << "# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS x y z rgb normal_x normal_y normal_z curvature\n"
    << "SIZE 4 4 4 4 4 4 4 4\nTYPE F F F U F F F F\nCOUNT 1 1 1 1 1 1 1 1\nWIDTH " << n << "\nHEIGHT 1\n"
    << "VIEWPOINT 0 0 0 1 0 0 0\nPOINTS " << n << "\nDATA binary\n";
  for (size_t i = 0; i < n; ++i) {
    float row[8] = {xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], 0.f, normals ? normals[3 * i] : 0.f,
                    normals ? normals[3 * i + 1] : 0.f, normals ? normals[3 * i + 2] : 0.f, 0.f};
    uint32_t c = rgb ? rgb[i] : 0u;
    std::memcpy(&row[3], &c, 4);
    f.write(reinterpret_cast<const char*>(row), sizeof(row));
  }
  return (bool)f;
}

}  // namespace io
}  // namespace ism3d
