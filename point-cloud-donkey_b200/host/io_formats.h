// io_formats.h — on-disk formats either side of the hot path:
//   * boost::archive::binary_oarchive / binary_iarchive framing for the .ismd model data
//     (utils/json_object.cpp:84-86,156-161; written for archive library version 17 = Boost 1.71 of the reference's
//     Ubuntu 20.04; the reader accepts versions >= 8, which share the layout).  "Format parity unpinned": the reference
//     ships no .ismd file to compare with.
//   * point clouds as ImplicitShapeModel::loadPointCloud reads them (implicit_shape_model.cpp:213-249, by the last
//     four characters of the file name): PCD ascii / binary / binary_compressed (LZF, the format of the reference's
//     vendored third_party/liblzf-3.6; fields stored structure-of-arrays) and PLY ascii / binary_little_endian /
//     binary_big_endian (pcl::io::loadPLYFile into PointXYZRGBNormal: x y z, nx ny nz, red green blue).
#pragma once
#include <algorithm>
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
  // pcl::PointCloud::width / height: height > 1 = an organized cloud (isOrganized(), a PCD with HEIGHT > 1);
  // 0 x 0 = unorganized / unknown
  size_t width = 0, height = 0;
  size_t size() const { return rgb.size(); }
  bool organized() const { return height > 1 && width * height == size(); }
};

// LZF decompression (the stream format of liblzf 3.6): a control byte < 32 starts a literal run of ctrl+1 bytes,
// otherwise a back reference of length (ctrl >> 5) + 2 (7 = extended by the next byte) at distance
// ((ctrl & 31) << 8 | next byte) + 1.  Returns the number of bytes produced, 0 on a corrupt stream.
inline size_t lzf_decompress(const unsigned char* in, size_t in_len, unsigned char* out, size_t out_len) {
  size_t ip = 0, op = 0;
  while (ip < in_len) {
    unsigned ctrl = in[ip++];
    if (ctrl < 32) {
      const size_t run = ctrl + 1;
      if (op + run > out_len || ip + run > in_len) return 0;
      std::memcpy(out + op, in + ip, run);
      op += run;
      ip += run;
    } else {
      size_t len = ctrl >> 5;
      if (ip >= in_len) return 0;
      size_t dist = (size_t)(ctrl & 0x1f) << 8;
      if (len == 7) {
        len += in[ip++];
        if (ip >= in_len) return 0;
      }
      dist += in[ip++];
      dist += 1;
      len += 2;
      if (dist > op || op + len > out_len) return 0;
      for (size_t i = 0; i < len; ++i, ++op) out[op] = out[op - dist];  // overlapping copies are the point
    }
  }
  return op;
}

// PCD scalar (TYPE I/U/F, SIZE 1/2/4/8) -> double
inline double scalar_as_double(const unsigned char* src, char type, int size) {
  if (type == 'F' && size == 4) { float t; std::memcpy(&t, src, 4); return t; }
  if (type == 'F' && size == 8) { double t; std::memcpy(&t, src, 8); return t; }
  if (size == 8) { int64_t t; std::memcpy(&t, src, 8); return type == 'U' ? (double)(uint64_t)t : (double)t; }
  if (size == 4) { int32_t t; std::memcpy(&t, src, 4); return type == 'U' ? (double)(uint32_t)t : (double)t; }
  if (size == 2) { int16_t t; std::memcpy(&t, src, 2); return type == 'U' ? (double)(uint16_t)t : (double)t; }
  if (size == 1) { int8_t t; std::memcpy(&t, src, 1); return type == 'U' ? (double)(uint8_t)t : (double)t; }
  return 0.0;
}

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
  out.width = width;
  out.height = height;
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
        store(p, (int)fi, scalar_as_double(reinterpret_cast<const unsigned char*>(src), types[fi], sizes[fi]), 0);
      }
  } else if (data_mode == "binary_compressed") {
    // uint32 compressed size, uint32 uncompressed size, LZF stream; the payload is structure-of-arrays:
    // all values of field 0, then all values of field 1, ...
    uint32_t csize = 0, usize = 0;
    f.read(reinterpret_cast<char*>(&csize), 4);
    f.read(reinterpret_cast<char*>(&usize), 4);
    if (!f || (size_t)usize != stride * points) { err = "PCD compressed header does not match the field table: " + path; return false; }
    std::vector<unsigned char> cbuf(csize), buf(usize);
    f.read(reinterpret_cast<char*>(cbuf.data()), (std::streamsize)csize);
    if ((size_t)f.gcount() != cbuf.size()) { err = "PCD compressed payload truncated: " + path; return false; }
    if (lzf_decompress(cbuf.data(), cbuf.size(), buf.data(), buf.size()) != buf.size()) {
      err = "PCD LZF stream is corrupt: " + path;
      return false;
    }
    size_t col = 0;
    for (size_t fi = 0; fi < fields.size(); ++fi) {
      const size_t fsz = (size_t)sizes[fi] * counts[fi];
      for (size_t p = 0; p < points; ++p) {
        const unsigned char* src = buf.data() + col + p * fsz;
        if ((int)fi == irgb) { uint32_t bits = 0; std::memcpy(&bits, src, std::min(4, sizes[fi])); store(p, (int)fi, 0, bits); continue; }
        store(p, (int)fi, scalar_as_double(src, types[fi], sizes[fi]), 0);
      }
      col += fsz * points;
    }
  } else {
    err = "PCD DATA mode '" + data_mode + "' is not supported: " + path;
    return false;
  }
  return true;
}


// ---- PLY ------------------------------------------------------------------------------------------------------------
// Vertex element only (faces and other elements are skipped), scalar properties of any PLY type; lists are skipped.
inline bool load_ply(const std::string& path, Cloud& out, std::string& err) {
  std::ifstream f(path, std::ios::binary);
  if (!f) { err = "could not open " + path; return false; }
  struct Prop { std::string name; int size; char kind; bool list; int cnt_size; char cnt_kind; };  // kind: F float, I int, U uint
  struct Elem { std::string name; size_t count; std::vector<Prop> props; };
  auto type_of = [](const std::string& t, int& size, char& kind) -> bool {
    if (t == "float" || t == "float32") { size = 4; kind = 'F'; }
    else if (t == "double" || t == "float64") { size = 8; kind = 'F'; }
    else if (t == "uchar" || t == "uint8") { size = 1; kind = 'U'; }
    else if (t == "char" || t == "int8") { size = 1; kind = 'I'; }
    else if (t == "ushort" || t == "uint16") { size = 2; kind = 'U'; }
    else if (t == "short" || t == "int16") { size = 2; kind = 'I'; }
    else if (t == "uint" || t == "uint32") { size = 4; kind = 'U'; }
    else if (t == "int" || t == "int32") { size = 4; kind = 'I'; }
    else return false;
    return true;
  };
  std::string line, format;
  std::vector<Elem> elems;
  if (!std::getline(f, line) || line.substr(0, 3) != "ply") { err = "not a PLY file: " + path; return false; }
  bool header_done = false;
  while (std::getline(f, line)) {
    if (!line.empty() && line.back() == '\r') line.pop_back();
    std::istringstream ls(line);
    std::string key;
    ls >> key;
    if (key == "format") ls >> format;
    else if (key == "element") { Elem e; ls >> e.name >> e.count; elems.push_back(e); }
    else if (key == "property") {
      if (elems.empty()) { err = "PLY property before any element: " + path; return false; }
      Prop p{};
      std::string t;
      ls >> t;
      if (t == "list") {
        std::string ct, vt;
        ls >> ct >> vt >> p.name;
        p.list = true;
        if (!type_of(ct, p.cnt_size, p.cnt_kind) || !type_of(vt, p.size, p.kind)) { err = "unknown PLY type in " + path; return false; }
      } else {
        ls >> p.name;
        if (!type_of(t, p.size, p.kind)) { err = "unknown PLY type '" + t + "' in " + path; return false; }
      }
      elems.back().props.push_back(p);
    } else if (key == "end_header") { header_done = true; break; }
  }
  if (!header_done || (format != "ascii" && format != "binary_little_endian" && format != "binary_big_endian")) {
    err = "malformed PLY header in " + path;
    return false;
  }
  const bool ascii = format == "ascii", swap = format == "binary_big_endian";
  auto read_scalar = [&](int size, char kind, double& v) -> bool {
    if (ascii) { return (bool)(f >> v); }
    unsigned char b[8];
    f.read(reinterpret_cast<char*>(b), size);
    if (!f) return false;
    if (swap) std::reverse(b, b + size);
    v = scalar_as_double(b, kind, size);
    return true;
  };
  bool have_vertex = false;
  for (const Elem& e : elems) {
    const bool is_vertex = e.name == "vertex";
    int ix = -1, iy = -1, iz = -1, inx = -1, iny = -1, inz = -1, ir = -1, ig = -1, ib = -1, irgb = -1;
    if (is_vertex) {
      for (size_t i = 0; i < e.props.size(); ++i) {
        const std::string& n = e.props[i].name;
        if (n == "x") ix = (int)i; else if (n == "y") iy = (int)i; else if (n == "z") iz = (int)i;
        else if (n == "nx" || n == "normal_x") inx = (int)i; else if (n == "ny" || n == "normal_y") iny = (int)i;
        else if (n == "nz" || n == "normal_z") inz = (int)i;
        else if (n == "red" || n == "r" || n == "diffuse_red") ir = (int)i;
        else if (n == "green" || n == "g" || n == "diffuse_green") ig = (int)i;
        else if (n == "blue" || n == "b" || n == "diffuse_blue") ib = (int)i;
        else if (n == "rgb" || n == "rgba") irgb = (int)i;
      }
      if (ix < 0 || iy < 0 || iz < 0) { err = "PLY vertex element has no x/y/z: " + path; return false; }
      out.has_normals = inx >= 0 && iny >= 0 && inz >= 0;
      out.has_rgb = (ir >= 0 && ig >= 0 && ib >= 0) || irgb >= 0;
      out.xyz.assign(e.count * 3, 0.f);
      out.normals.assign(e.count * 3, 0.f);
      out.rgb.assign(e.count, 0u);
      have_vertex = true;
    }
    for (size_t p = 0; p < e.count; ++p) {
      unsigned r = 0, g = 0, b = 0;
      for (size_t i = 0; i < e.props.size(); ++i) {
        const Prop& pr = e.props[i];
        double v = 0;
        if (pr.list) {
          double n = 0;
          if (!read_scalar(pr.cnt_size, pr.cnt_kind, n)) { err = "PLY payload truncated: " + path; return false; }
          for (int k = 0; k < (int)n; ++k)
            if (!read_scalar(pr.size, pr.kind, v)) { err = "PLY payload truncated: " + path; return false; }
          continue;
        }
        if (!read_scalar(pr.size, pr.kind, v)) { err = "PLY payload truncated: " + path; return false; }
        if (!is_vertex) continue;
        const int ii = (int)i;
        if (ii == ix) out.xyz[3 * p] = (float)v; else if (ii == iy) out.xyz[3 * p + 1] = (float)v;
        else if (ii == iz) out.xyz[3 * p + 2] = (float)v;
        else if (ii == inx) out.normals[3 * p] = (float)v; else if (ii == iny) out.normals[3 * p + 1] = (float)v;
        else if (ii == inz) out.normals[3 * p + 2] = (float)v;
        else if (ii == ir) r = (unsigned)v & 0xffu; else if (ii == ig) g = (unsigned)v & 0xffu;
        else if (ii == ib) b = (unsigned)v & 0xffu;
        else if (ii == irgb) {
          uint32_t bits;
          if (pr.kind == 'F') { float fv = (float)v; std::memcpy(&bits, &fv, 4); } else bits = (uint32_t)v;
          out.rgb[p] = bits & 0x00ffffffu;
        }
      }
      if (is_vertex && irgb < 0 && out.has_rgb) out.rgb[p] = (r << 16) | (g << 8) | b;
    }
    if (is_vertex) break;  // nothing after the vertices is needed
  }
  if (!have_vertex) { err = "PLY file has no vertex element: " + path; return false; }
  return true;
}

// ImplicitShapeModel::loadPointCloud (implicit_shape_model.cpp:213-249): the last four characters pick the reader
inline bool load_cloud(const std::string& path, Cloud& out, std::string& err) {
  if (path.size() < 5) { err = "invalid filename: " + path; return false; }
  const std::string ext = path.substr(path.size() - 4, 4);
  if (ext == ".pcd") return load_pcd(path, out, err);
  if (ext == ".ply") return load_ply(path, out, err);
  err = "Unknown extension: " + ext;
  return false;
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
