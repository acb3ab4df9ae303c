// ismd_roundtrip.cpp — CPU-only checks of the on-disk formats (no GPU call): boost-archive framing, JSON config
// parse/write, PCD ascii/binary read.  Driven by tests/test_host_formats.py.
#include <cmath>
#include <cstdio>
#include <iostream>
#include <sstream>

#include "io_formats.h"
#include "json_min.h"

using namespace ism3d;

static int fail(const char* what) {
  std::fprintf(stderr, "FAIL: %s\n", what);
  return 1;
}

int main(int argc, char** argv) {
  if (argc < 2) return fail("usage: ismd_roundtrip <archive|json FILE|pcd FILE>");
  std::string cmd = argv[1];
  if (cmd == "archive") {
    std::stringstream ss(std::ios::in | std::ios::out | std::ios::binary);
    {
      io::BinaryOArchive oa(ss);
      oa.put<uint32_t>(7);
      oa.put<float>(1.5f);
      oa.put_vector(std::vector<float>{1.f, 2.f, 3.f});
      oa.put_vector(std::vector<uint32_t>{4u, 5u});
      oa.put_vector(std::vector<float>{});
      oa.put_string("hello");
    }
    const std::string bytes = ss.str();
    // header: u64 22 | "serialization::archive" | u16 17 | sizes 4 8 4 8 | i32 1  = 8 + 22 + 2 + 4 + 4 = 40 bytes
    if (bytes.size() != 40 + 4 + 4 + (5 + 8 + 12) + (5 + 8 + 8) + 8 + (8 + 5)) return fail("archive size");
    if ((unsigned char)bytes[30] != 17 || bytes[31] != 0) return fail("library version bytes");
    if (bytes[32] != 4 || bytes[33] != 8 || bytes[34] != 4 || bytes[35] != 8) return fail("native sizes");
    std::stringstream in(bytes, std::ios::in | std::ios::binary);
    io::BinaryIArchive ia(in);
    if (ia.version() != 17) return fail("version");
    if (ia.get<uint32_t>() != 7 || ia.get<float>() != 1.5f) return fail("primitives");
    auto f = ia.get_vector_f();
    auto u = ia.get_vector_u();
    auto e = ia.get_vector_f();
    if (f.size() != 3 || f[2] != 3.f || u.size() != 2 || u[1] != 5u || !e.empty()) return fail("vectors");
    if (ia.get_string() != "hello") return fail("string");
    std::puts("archive ok");
    return 0;
  }
  if (cmd == "json" && argc > 2) {
    jsonmin::Value v = jsonmin::parse_file(argv[2]);
    if (!v.isObject() || !v.isMember("ObjectConfig")) return fail("ObjectConfig missing");
    const jsonmin::Value& oc = v["ObjectConfig"];
    std::cout << "DistanceType=" << oc["Parameters"]["DistanceType"].str << " Features=" << oc["Children"]["Features"]["Type"].str
              << " Radius=" << oc["Children"]["Features"]["Parameters"]["Radius"].num
              << " K=" << oc["Children"]["Codebook"]["Children"]["ActivationStrategy"]["Parameters"]["K"].num
              << " Bandwidth=" << oc["Children"]["Voting"]["Parameters"]["Bandwidth"].num << std::endl;
    if (argc > 3) {
      if (!jsonmin::write_file(v, argv[3])) return fail("write");
      jsonmin::Value w = jsonmin::parse_file(argv[3]);
      if (w["ObjectConfig"]["Children"]["Voting"]["Parameters"]["Bandwidth"].num !=
          oc["Children"]["Voting"]["Parameters"]["Bandwidth"].num)
        return fail("json round trip");
    }
    return 0;
  }
  if (cmd == "pcd" && argc > 2) {
    io::Cloud c;
    std::string err;
    if (!io::load_cloud(argv[2], c, err)) return fail(err.c_str());
    double sx = 0, sn = 0;
    unsigned long long sc = 0;
    for (float v : c.xyz) sx += v;
    for (float v : c.normals) sn += v;
    for (uint32_t v : c.rgb) sc += v;
    std::printf("points=%zu normals=%d rgb=%d sum_xyz=%.6f sum_n=%.6f sum_rgb=%llu\n", c.size(), (int)c.has_normals,
                (int)c.has_rgb, sx, sn, sc);
    return 0;
  }
  return fail("unknown command");
}
