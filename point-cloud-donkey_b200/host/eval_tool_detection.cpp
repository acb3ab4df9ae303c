// eval_tool_detection.cpp — the reference's detection front-end (src/eval_tool/eval_detection.cpp:298-462) on top of
// the B200 path: "-d model.ism -f test_list.txt -o outdir" runs ImplicitShapeModel::detect on every scene of a
// "# test detection" list, matches the maxima with the ground-truth annotations and writes summary.txt with per-class
// precision / recall / AP, mAP and the dataset-wide AP (metrics: eval_detection.h).  Training of detection models goes
// through eval_tool -t (same model format).
#include <sys/stat.h>

#include <chrono>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <map>
#include <string>
#include <vector>

#include "eval_detection.h"
#include "ism3d_b200.h"

int main(int argc, char** argv) {
  std::string model, list, outdir;
  int device = 0;
  for (int i = 1; i < argc; ++i) {
    const std::string a = argv[i];
    auto next = [&]() -> std::string { return i + 1 < argc ? argv[++i] : std::string(); };
    if (a == "-d" || a == "--detect") model = next();
    else if (a == "-f" || a == "--inputfile") list = next();
    else if (a == "-o" || a == "--output") outdir = next();
    else if (a == "--device") device = std::stoi(next());
    else if (a == "-h" || a == "--help") {
      std::cout << "eval_tool_detection -d trained.ism -f test_list.txt -o outdir [--device N]\n"
                   "  test_list.txt: '# test detection' header, then '<cloud> <annotation file>' per line\n"
                   "  annotation:    '<class> (<occlusion>) x y z [sx sy sz qw qx qy qz]' per object\n";
      return 0;
    }
  }
  if (model.empty() || list.empty()) {
    std::cerr << "No input file provided! You need -d <model> and -f <list>" << std::endl;
    return 1;
  }
  try {
    std::vector<std::string> clouds, annots;
    evaldet::parseFileListDetectionTest(list, clouds, annots);
    if (clouds.empty() || clouds.size() != annots.size()) {
      std::cerr << "number of point clouds does not match the number of groundtruth files or is zero" << std::endl;
      return 1;
    }
    std::cout << "starting the detection process" << std::endl;
    ism3d::ImplicitShapeModel ism(device);
    ism.setSignalsState(false);
    if (!ism.readObject(model)) {
      std::cerr << "could not read ism from file, detection stopped: " << model << std::endl;
      return 1;
    }
    const std::map<unsigned, std::string> class_names = ism.getClassLabels(), inst_names = ism.getInstanceLabels();
    const std::map<unsigned, unsigned> inst_to_class = ism.getInstanceClassMap();
    const bool inst_primary = ism.isInstancePrimaryLabel() && !inst_names.empty();
    auto name_of = [](const std::map<unsigned, std::string>& m, unsigned id) {
      auto it = m.find(id);
      return it == m.end() ? std::to_string(id) : it->second;
    };
    if (!outdir.empty()) mkdir(outdir.c_str(), 0755);
    std::vector<evaldet::DetectionObject> gt_objects, detections;
    std::map<std::string, double> times;
    const auto t0 = std::chrono::steady_clock::now();
    for (size_t i = 0; i < clouds.size(); ++i) {
      std::cout << "Processing file: " << clouds[i] << std::endl;
      std::vector<ism3d::VotingMaximum> maxima;
      if (!ism.detect(clouds[i], maxima, times)) {
        std::cerr << "detection failed" << std::endl;
        return 1;
      }
      std::vector<evaldet::DetectionObject> g = evaldet::parseAnnotationFile(annots[i]);
      gt_objects.insert(gt_objects.end(), g.begin(), g.end());
      for (const ism3d::VotingMaximum& m : maxima) {  // convertMaxToObj (eval_helpers_detection.h:415-438)
        evaldet::DetectionObject o;
        if (inst_primary) {
          auto c = inst_to_class.find(m.classId);
          o.class_label = name_of(class_names, c == inst_to_class.end() ? m.classId : c->second);
          o.instance_label = name_of(inst_names, m.classId);
        } else {
          o.class_label = name_of(class_names, m.classId);
          o.instance_label = name_of(inst_names, m.instanceId);
        }
        o.position[0] = m.position[0];
        o.position[1] = m.position[1];
        o.position[2] = m.position[2];
        o.confidence = m.weight;
        o.filepath = annots[i];
        detections.push_back(o);
      }
    }
    const double total_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    std::map<std::string, float> thr;  // ImplicitShapeModel::getDetectionThreshold, by class name
    for (const auto& it : ism.getDetectionThreshold()) thr[name_of(class_names, it.first)] = it.second;
    const evaldet::DatasetMetrics dm = evaldet::evaluate(gt_objects, detections, thr);
    std::ostringstream s;
    s << std::fixed << std::setprecision(4);
    s << "class            num gt   tp   fp   precision   recall   AP\n";
    for (const auto& kv : dm.per_class)
      s << std::left << std::setw(16) << kv.first << std::right << std::setw(7) << kv.second.num_gt << std::setw(5)
        << kv.second.tp << std::setw(5) << kv.second.fp << std::setw(12) << kv.second.precision << std::setw(9)
        << kv.second.recall << std::setw(9) << kv.second.ap << "\n";
    s << "\nscenes: " << clouds.size() << ", ground-truth objects: " << dm.num_gt << ", detections: " << detections.size()
      << ", tp: " << dm.tp << ", fp: " << dm.fp << "\n";
    s << "mAP: " << dm.mAP << "\nmean precision: " << dm.mPrecision << "\nmean recall: " << dm.mRecall
      << "\noverall AP (all detections, one sweep): " << dm.overall_ap << "\n";
    s << "total processing time: " << total_s << " [s]\n";
    for (const auto& t : times) s << "time " << t.first << ": " << t.second / 1000.0 << " [s]\n";
    std::cout << s.str();
    if (!outdir.empty()) {
      std::ofstream(outdir + "/summary.txt") << s.str();
      std::ofstream pr(outdir + "/precision_recall.txt");
      for (size_t i = 0; i < dm.precisions.size(); ++i) pr << dm.recalls[i] << " " << dm.precisions[i] << "\n";
    } else {
      std::cerr << "no output file specified, detected maxima will not be saved" << std::endl;
    }
  } catch (const std::exception& e) {
    std::cerr << e.what() << std::endl;
    return 1;
  }
  return 0;
}
