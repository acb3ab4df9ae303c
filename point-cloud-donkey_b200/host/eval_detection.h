// eval_detection.h — detection metrics of the reference's eval_tool_detection (src/eval_tool/eval_helpers_detection.h):
// annotation / list-file parsing (:441-500, :612-700), greedy ground-truth matching by confidence within a per-class
// distance threshold (:225-291), per-class precision / recall / AP (:125-141, :308-340) and the dataset-wide
// precision-recall sweep with its overall AP (:147-222).  Host-only (no GPU call): the detections come from
// ImplicitShapeModel::detect.  The occlusion field and the optional box of an annotation are parsed and ignored, as
// the reference's metrics ignore them.
#pragma once
#include <algorithm>
#include <cmath>
#include <fstream>
#include <limits>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <tuple>
#include <vector>

namespace evaldet {

struct DetectionObject {
  std::string class_label, instance_label;
  float position[3] = {0, 0, 0};
  float occlusion = 0.f;
  float confidence = 1.f;
  std::string filepath;  // the annotation file: detections and ground truth of one scene share it
};

// "class (occlusion) x y z [sx sy sz qw qx qy qz]" per line (eval_helpers_detection.h:441-500)
inline std::vector<DetectionObject> parseAnnotationFile(const std::string& filename) {
  std::vector<DetectionObject> objects;
  std::ifstream file(filename);
  if (!file) throw std::runtime_error("could not open annotation file " + filename);
  std::string line;
  while (std::getline(file, line)) {
    if (line.empty()) continue;
    std::stringstream iss(line);
    std::string item;
    std::vector<std::string> tokens;
    while (std::getline(iss, item, ' '))
      if (!item.empty()) tokens.push_back(item);
    if (tokens.empty()) continue;
    if (tokens.size() != 5 && tokens.size() != 12)
      throw std::runtime_error("annotation line with " + std::to_string(tokens.size()) + " tokens (expected 5 or 12) in " + filename);
    DetectionObject o;
    o.class_label = tokens[0];
    if (o.class_label == "book" || o.class_label == "books" || o.class_label == "dress") continue;  // the reference's SUN RGB-D fix
    o.instance_label = o.class_label;
    std::string occ = tokens[1];
    occ = occ.substr(1, occ.find_first_of(')') - 1);
    o.occlusion = std::stof(occ);
    for (int a = 0; a < 3; ++a) o.position[a] = std::stof(tokens[2 + a]);
    o.confidence = 1.0f;
    o.filepath = filename;
    objects.push_back(o);
  }
  return objects;
}

// "# test detection [inst]" header, then "cloud annotation" pairs (eval_helpers_detection.h:647-700)
inline void parseFileListDetectionTest(const std::string& input, std::vector<std::string>& clouds,
                                       std::vector<std::string>& annotations) {
  std::ifstream infile(input);
  if (!infile) throw std::runtime_error("File " + input + " does not exist!");
  std::string file, gt_file, flag, flag2;
  infile >> file >> gt_file >> flag >> flag2;
  if (file != "#" || gt_file != "test") throw std::runtime_error("the input file must start with '# test detection'");
  if (flag != "detection")
    throw std::runtime_error("You are using a classification data set with the detection eval_tool! Use 'eval_tool' instead.");
  if (flag2 == "inst") {
    if (!(infile >> file)) return;
  } else {
    file = flag2;  // without "inst" the first file name was read into the flag
  }
  if (!(infile >> gt_file)) return;
  clouds.push_back(file);
  annotations.push_back(gt_file);
  while (infile >> file >> gt_file) {
    if (file[0] == '#') continue;
    clouds.push_back(file);
    annotations.push_back(gt_file);
  }
}

// eval_helpers_detection.h:225-276 for one class: detections sorted by confidence (descending); each takes the closest
// unused ground-truth object of the same scene and class; a match farther than the threshold (or none) is a false positive
inline std::pair<std::vector<int>, std::vector<int>> match_gt_objects(const std::vector<DetectionObject>& gt,
                                                                      std::vector<DetectionObject>& det,
                                                                      float dist_threshold) {
  std::stable_sort(det.begin(), det.end(), [](const DetectionObject& a, const DetectionObject& b) { return a.confidence > b.confidence; });
  std::vector<bool> used(gt.size(), false);
  std::vector<int> tp(det.size(), 0), fp(det.size(), 0);
  for (size_t d = 0; d < det.size(); ++d) {
    float best = std::numeric_limits<float>::max();
    int best_i = -1;
    for (size_t g = 0; g < gt.size(); ++g) {
      if (det[d].filepath != gt[g].filepath || det[d].class_label != gt[g].class_label) continue;
      const float dx = gt[g].position[0] - det[d].position[0], dy = gt[g].position[1] - det[d].position[1],
                  dz = gt[g].position[2] - det[d].position[2];
      const float dist = std::sqrt(dx * dx + dy * dy + dz * dz);
      if (dist < best && !used[g]) {
        best = dist;
        best_i = (int)g;
      }
    }
    if (best > dist_threshold || best_i == -1) {
      fp[d] = 1;
    } else {
      tp[d] = 1;
      used[(size_t)best_i] = true;
    }
  }
  return {tp, fp};
}

struct ClassMetrics {
  float precision = 0, recall = 0, ap = 0;
  int num_gt = 0, tp = 0, fp = 0;
  std::vector<int> tps, fps;
};

// computeAllMetrics (eval_helpers_detection.h:308-340)
inline ClassMetrics computeAllMetrics(const std::vector<DetectionObject>& gt, std::vector<DetectionObject>& det,
                                      float dist_threshold) {
  ClassMetrics m;
  std::tie(m.tps, m.fps) = match_gt_objects(gt, det, dist_threshold);
  m.num_gt = (int)gt.size();
  for (int v : m.tps) m.tp += v;
  for (int v : m.fps) m.fp += v;
  m.precision = m.tp / float(m.fp + m.tp);  // 0/0 -> NaN as in the reference
  m.recall = m.num_gt == 0 ? 0 : float(m.tp) / m.num_gt;
  int cumul = 0;
  for (size_t i = 0; i < m.tps.size(); ++i)
    if (m.tps[i] == 1) {
      cumul += 1;
      m.ap += (float(cumul) / (i + 1)) * (1.0 / m.num_gt);
    }
  return m;
}

struct DatasetMetrics {
  std::map<std::string, ClassMetrics> per_class;
  float mAP = 0, mPrecision = 0, mRecall = 0;      // means over the classes that have ground truth
  float overall_ap = 0;                            // dataset-wide sweep over all detections
  std::vector<float> precisions, recalls;          // the precision-recall curve of that sweep
  int num_gt = 0, tp = 0, fp = 0;
};

// the per-class loop of eval_detection.cpp:427-440 + computePrecisionRecallForPlotting (:147-222)
inline DatasetMetrics evaluate(const std::vector<DetectionObject>& gt_objects, const std::vector<DetectionObject>& detections,
                               const std::map<std::string, float>& dist_threshold_per_class) {
  std::map<std::string, std::vector<DetectionObject>> gt_map, det_map;
  for (const auto& o : gt_objects) gt_map[o.class_label].push_back(o);
  for (const auto& o : detections) det_map[o.class_label].push_back(o);
  DatasetMetrics out;
  struct Summary { float confidence; int tp, fp; };
  std::vector<Summary> all;
  for (auto& kv : gt_map) {
    auto thr = dist_threshold_per_class.find(kv.first);
    const float t = thr == dist_threshold_per_class.end() ? 0.f : thr->second;
    std::vector<DetectionObject>& det = det_map[kv.first];
    ClassMetrics m = computeAllMetrics(kv.second, det, t);
    out.num_gt += m.num_gt;
    out.tp += m.tp;
    out.fp += m.fp;
    out.mAP += m.ap;
    out.mPrecision += std::isnan(m.precision) ? 0.f : m.precision;
    out.mRecall += m.recall;
    for (size_t i = 0; i < det.size(); ++i) all.push_back({det[i].confidence, m.tps[i], m.fps[i]});
    out.per_class[kv.first] = std::move(m);
  }
  // detections of classes without any ground truth count neither way in the reference's sweep (tp = fp = 0)
  for (auto& kv : det_map)
    if (!gt_map.count(kv.first))
      for (size_t i = 0; i < kv.second.size(); ++i) all.push_back({0.0f, 0, 0});
  const float n = (float)std::max<size_t>(1, gt_map.size());
  out.mAP /= n;
  out.mPrecision /= n;
  out.mRecall /= n;
  std::stable_sort(all.begin(), all.end(), [](const Summary& a, const Summary& b) { return a.confidence > b.confidence; });
  int tp_sum = 0, fp_sum = 0;
  for (const Summary& d : all) {
    tp_sum += d.tp;
    fp_sum += d.fp;
    out.precisions.push_back(tp_sum / float(fp_sum + tp_sum));
    out.recalls.push_back(float(tp_sum) / out.num_gt);
    if (d.tp == 1) out.overall_ap += (float(tp_sum) / (tp_sum + fp_sum)) * (1.0 / out.num_gt);
  }
  return out;
}

}  // namespace evaldet
