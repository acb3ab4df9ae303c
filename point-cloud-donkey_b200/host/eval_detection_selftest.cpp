// CPU-only known-answer test of host/eval_detection.h (run by tests/test_host_formats.py): matching, AP, parsers.
#include <cstdio>
#include <cstdlib>

#include "eval_detection.h"

#define CHECK(c)                                                         \
  do {                                                                   \
    if (!(c)) {                                                          \
      std::fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #c); \
      return 1;                                                          \
    }                                                                    \
  } while (0)

static evaldet::DetectionObject obj(const char* cls, float x, float y, float z, float conf, const char* file) {
  evaldet::DetectionObject o;
  o.class_label = o.instance_label = cls;
  o.position[0] = x; o.position[1] = y; o.position[2] = z;
  o.confidence = conf;
  o.filepath = file;
  return o;
}

int main(int argc, char** argv) {
  using namespace evaldet;
  // two chairs and a table in scene A, one chair in scene B
  std::vector<DetectionObject> gt = {obj("chair", 0, 0, 0, 1, "A"), obj("chair", 2, 0, 0, 1, "A"), obj("table", 5, 5, 0, 1, "A"),
                                     obj("chair", 1, 1, 1, 1, "B")};
  // chair detections by confidence: hit, miss (too far), duplicate of the first gt (already used -> next gt is too far),
  // hit in scene B, right position but wrong scene
  std::vector<DetectionObject> det = {obj("chair", 0.05f, 0, 0, 0.9f, "A"), obj("chair", 9, 9, 9, 0.8f, "A"),
                                      obj("chair", 0.02f, 0, 0, 0.7f, "A"), obj("chair", 1, 1.1f, 1, 0.6f, "B"),
                                      obj("chair", 2, 0, 0, 0.5f, "B"), obj("table", 5.1f, 5, 0, 0.4f, "A"),
                                      obj("lamp", 0, 0, 0, 0.99f, "A")};
  DatasetMetrics m = evaluate(gt, det, {{"chair", 0.3f}, {"table", 0.3f}});
  const ClassMetrics& c = m.per_class.at("chair");
  CHECK(c.num_gt == 3 && c.tp == 2 && c.fp == 3);
  CHECK(c.tps == std::vector<int>({1, 0, 0, 1, 0}));
  CHECK(std::fabs(c.precision - 0.4f) < 1e-6f && std::fabs(c.recall - 2.f / 3.f) < 1e-6f);
  CHECK(std::fabs(c.ap - (1.0f / 1 + 2.0f / 4) / 3) < 1e-6f);  // hits at ranks 1 and 4
  const ClassMetrics& t = m.per_class.at("table");
  CHECK(t.tp == 1 && t.fp == 0 && std::fabs(t.ap - 1.0f) < 1e-6f);
  CHECK(std::fabs(m.mAP - (c.ap + 1.0f) / 2) < 1e-6f);
  CHECK(m.num_gt == 4 && m.tp == 3 && m.fp == 3);
  // dataset sweep: confidences 0.9 T, 0.8 F, 0.7 F, 0.6 T, 0.5 F, 0.4 T (the lamp has no ground truth: tp = fp = 0 at the end)
  CHECK(std::fabs(m.overall_ap - (1.f / 1 + 2.f / 4 + 3.f / 6) / 4) < 1e-6f);
  CHECK(m.recalls.size() == 7 && std::fabs(m.recalls.back() - 0.75f) < 1e-6f);
  if (argc > 2) {  // parsers: <annotation file> <list file>
    std::vector<DetectionObject> a = parseAnnotationFile(argv[1]);
    CHECK(a.size() == 2 && a[0].class_label == "chair" && std::fabs(a[0].occlusion - 0.25f) < 1e-6f);
    CHECK(std::fabs(a[1].position[2] - 3.5f) < 1e-6f);
    std::vector<std::string> clouds, annots;
    parseFileListDetectionTest(argv[2], clouds, annots);
    CHECK(clouds.size() == 2 && annots.size() == 2 && clouds[1] == "scene2.pcd" && annots[0] == "scene1.txt");
  }
  std::printf("eval_detection selftest ok\n");
  return 0;
}
