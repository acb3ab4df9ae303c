// eval_tool.cpp — the reference's classification front-end (src/eval_tool/eval_classification.cpp) on top of the
// B200 path: same flags (-t/-d/-f/-o/-i/-m/-c/-p/-g, eval_classification.cpp:49-63), same list-file format
// (src/eval_tool/eval_helpers.h:100-177), same summary.txt lines.  The test list is classified in batches through
// pcdb_classify_batch instead of one detect() per cloud (eval_classification.cpp:347-356); per-cloud outputs are the same.
#include <sys/stat.h>

#include <chrono>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <thread>
#include <memory>
#include <map>
#include <sstream>
#include <string>
#include <vector>

#include "ism3d_b200.h"

namespace {

enum class LabelUsage { CLASS_ONLY, BOTH_GIVEN, CLASS_PRIMARY, INSTANCE_PRIMARY };

std::map<std::string, unsigned> class_labels_map, instance_labels_map;
std::map<unsigned, std::string> class_labels_rmap, instance_labels_rmap;
std::map<unsigned, unsigned> instance_to_class_map;

unsigned convertLabel(const std::string& label, std::map<std::string, unsigned>& m, std::map<unsigned, std::string>& r) {
  auto it = m.find(label);
  if (it != m.end()) return it->second;
  unsigned id = (unsigned)m.size();
  m.insert({label, id});
  r.insert({id, label});
  return id;
}

// eval_helpers.h:100-177
LabelUsage parseFileList(const std::string& input, std::vector<std::string>& filenames, std::vector<unsigned>& class_labels,
                         std::vector<unsigned>& instance_labels, std::string& mode) {
  std::ifstream infile(input);
  if (!infile) throw ism3d::RuntimeException("could not open input file " + input);
  std::string file, class_label, instance_label;
  bool using_instances = false;
  infile >> file >> class_label >> instance_label;
  if (file == "#" && (class_label == "train" || class_label == "test")) {
    mode = class_label;
    if (instance_label == "inst") using_instances = true;
    if (instance_label == "detection") {
      std::cerr << "ERROR: You are using a detection data set with the classification eval_tool!" << std::endl;
      std::exit(1);
    }
  }
  if (using_instances) {
    while (infile >> file >> class_label >> instance_label) {
      if (file[0] == '#') continue;
      filenames.push_back(file);
      unsigned c = convertLabel(class_label, class_labels_map, class_labels_rmap);
      unsigned i = convertLabel(instance_label, instance_labels_map, instance_labels_rmap);
      instance_to_class_map.insert({i, c});
      class_labels.push_back(c);
      instance_labels.push_back(i);
    }
    return LabelUsage::BOTH_GIVEN;
  }
  file = instance_label;
  infile >> class_label;
  filenames.push_back(file);
  class_labels.push_back(convertLabel(class_label, class_labels_map, class_labels_rmap));
  while (infile >> file >> class_label) {
    if (file[0] == '#') continue;
    filenames.push_back(file);
    unsigned c = convertLabel(class_label, class_labels_map, class_labels_rmap);
    class_labels.push_back(c);
    instance_to_class_map.insert({c, c});
  }
  return LabelUsage::CLASS_ONLY;
}

struct Options {
  std::map<std::string, std::vector<std::string>> v;
  bool has(const std::string& k) const { return v.count(k) > 0; }
  const std::string& one(const std::string& k) const { return v.at(k).at(0); }
};

Options parse(int argc, char** argv) {
  static const std::map<std::string, std::string> alias = {
      {"-h", "help"},  {"--help", "help"},   {"-o", "output"},  {"--output", "output"}, {"-f", "inputfile"},
      {"--inputfile", "inputfile"}, {"-t", "train"}, {"--train", "train"}, {"-i", "inplace"}, {"--inplace", "inplace"},
      {"-m", "models"}, {"--models", "models"}, {"-c", "classes"}, {"--classes", "classes"}, {"-d", "detect"},
      {"--detect", "detect"}, {"-p", "pointclouds"}, {"--pointclouds", "pointclouds"}, {"-g", "groundtruth"},
      {"--groundtruth", "groundtruth"}, {"--batch", "batch"}, {"--device", "device"}, {"--gpus", "gpus"},
      {"--log", "log"}};
  Options o;
  std::string cur;
  for (int i = 1; i < argc; ++i) {
    std::string a = argv[i];
    auto it = alias.find(a);
    if (it != alias.end()) {
      cur = it->second;
      o.v[cur];
    } else if (!cur.empty())
      o.v[cur].push_back(a);
    else
      throw ism3d::BadParamException("unexpected argument " + a);
  }
  return o;
}

}  // namespace

int main(int argc, char** argv) {
  try {
    Options opt = parse(argc, argv);
    if (opt.has("help") || argc == 1) {
      std::cout << "Usage: eval_tool [options]\n"
                   "  -h  help\n  -o <dir|file>  output folder (classification) or output .ism (training)\n"
                   "  -f <list>      input file with clouds and labels (first line '# train|test [inst]')\n"
                   "  -t <ism>       train an implicit shape model      -i  overwrite the loaded ism file\n"
                   "  -m <pcd...> -c <ids...>   training models and class ids on the command line\n"
                   "  -d <ism>       classify with a trained model\n"
                   "  -p <pcd...> -g <ids...>   test clouds and ground-truth ids on the command line\n"
                   "  --batch N  clouds per GPU batch (default 256)   --device N  CUDA device\n"
                   "  --gpus N   classification only: shard the test list over N GPUs (devices device..device+N-1),\n"
                   "             one context and one host thread per GPU, no communication\n";
      return 0;
    }
    const int device = opt.has("device") ? std::stoi(opt.one("device")) : 0;
    const int batch = opt.has("batch") ? std::stoi(opt.one("batch")) : 256;
    std::vector<std::string> filenames;
    std::vector<unsigned> class_labels, instance_labels;
    std::string mode;
    LabelUsage label_usage = LabelUsage::CLASS_ONLY;
    if (opt.has("inputfile")) label_usage = parseFileList(opt.one("inputfile"), filenames, class_labels, instance_labels, mode);

    if ((opt.has("train") && mode.empty()) || mode == "train") {
      std::cout << "starting the training process" << std::endl;
      const std::string ismFile = opt.has("train") ? opt.one("train") : opt.one("detect");
      ism3d::ImplicitShapeModel ism(device);
      ism.setLogging(opt.has("log"));
      ism.setSignalsState(false);
      if (!ism.readObject(ismFile, true)) {
        std::cerr << "could not read ism from file, training stopped: " << ismFile << std::endl;
        return 1;
      }
      if (label_usage == LabelUsage::BOTH_GIVEN)
        label_usage = ism.isInstancePrimaryLabel() ? LabelUsage::INSTANCE_PRIMARY : LabelUsage::CLASS_PRIMARY;
      if (opt.has("output")) ism.setOutputFilename(opt.one("output"));
      std::vector<std::string> models;
      std::vector<unsigned> class_ids, instance_ids;
      if (opt.has("models") && opt.has("classes")) {
        models = opt.v["models"];
        for (auto& s : opt.v["classes"]) class_ids.push_back((unsigned)std::stoul(s));
        instance_ids = class_ids;
      } else if (!filenames.empty()) {
        models = filenames;
        if (label_usage == LabelUsage::CLASS_ONLY) { class_ids = class_labels; instance_ids = class_labels; }
        else if (label_usage == LabelUsage::CLASS_PRIMARY) { class_ids = class_labels; instance_ids = instance_labels; }
        else { class_ids = instance_labels; instance_ids = instance_labels; }
      }
      if (models.size() != class_ids.size()) {
        std::cerr << "number of models does not match the number of class ids" << std::endl;
        return 1;
      }
      for (size_t i = 0; i < models.size(); ++i)
        if (!ism.addTrainingModel(models[i], class_ids[i], instance_ids[i])) {
          std::cerr << "could not add training model: " << models[i] << ", class " << class_ids[i] << std::endl;
          return 1;
        }
      ism.train();
      ism.setLabels(class_labels_rmap, instance_labels_rmap, instance_to_class_map);
      if (opt.has("inplace")) {
        if (!ism.writeObject(ismFile, ismFile + "d")) { std::cerr << "could not write ism" << std::endl; return 1; }
      } else if (opt.has("output")) {
        if (!ism.writeObject(opt.one("output"))) { std::cerr << "could not write ism" << std::endl; return 1; }
      } else {
        std::cerr << "the trained ism is not saved" << std::endl;
        return 1;
      }
    }

    if ((opt.has("detect") && mode.empty()) || mode == "test") {
      std::cout << "starting the classification process" << std::endl;
      const std::string ismFile = opt.has("detect") ? opt.one("detect") : opt.one("train");
      ism3d::ImplicitShapeModel ism(device);
      ism.setLogging(opt.has("log"));
      ism.setSignalsState(false);
      if (!ism.readObject(ismFile)) {
        std::cerr << "could not read ism from file, classification stopped: " << ismFile << std::endl;
        return 1;
      }
      std::vector<std::string> pointClouds;
      std::vector<unsigned> gt_class_ids, gt_instance_ids;
      class_labels_rmap = ism.getClassLabels();
      instance_labels_rmap = ism.getInstanceLabels();
      instance_to_class_map = ism.getInstanceClassMap();
      if (label_usage == LabelUsage::BOTH_GIVEN)
        label_usage = ism.isInstancePrimaryLabel() ? LabelUsage::INSTANCE_PRIMARY : LabelUsage::CLASS_PRIMARY;
      if (opt.has("pointclouds") && opt.has("groundtruth")) {
        pointClouds = opt.v["pointclouds"];
        for (auto& s : opt.v["groundtruth"]) gt_class_ids.push_back((unsigned)std::stoul(s));
        gt_instance_ids = gt_class_ids;
      } else if (!filenames.empty()) {
        pointClouds = filenames;
        gt_class_ids = class_labels;
        gt_instance_ids = label_usage == LabelUsage::CLASS_ONLY ? class_labels : instance_labels;
      }
      if (pointClouds.size() != gt_class_ids.size() || pointClouds.empty()) {
        std::cerr << "number of point clouds does not match the number of groundtruth ids" << std::endl;
        return 1;
      }
      std::ofstream summaryFile;
      if (opt.has("output")) {
        mkdir(opt.one("output").c_str(), 0755);
        summaryFile.open(opt.one("output") + "/summary.txt", std::ios::out);
      } else
        std::cerr << "no output file specified, detected maxima will not be saved" << std::endl;
      auto t0 = std::chrono::steady_clock::now();
      std::vector<std::vector<ism3d::VotingMaximum>> all;
      std::map<std::string, double> times;
      const int n_gpus = opt.has("gpus") ? std::max(1, std::stoi(opt.one("gpus"))) : 1;
      if (n_gpus <= 1) {
        if (!ism.detectBatch(pointClouds, all, times, batch)) {
          std::cerr << "classification failed" << std::endl;
          return 1;
        }
      } else {
        // test clouds are independent: contiguous shards, codebook replicated, results concatenated
        all.assign(pointClouds.size(), {});
        std::vector<std::map<std::string, double>> shard_times((size_t)n_gpus);
        std::vector<int> ok((size_t)n_gpus, 1);
        std::vector<std::string> errors((size_t)n_gpus);
        std::vector<std::thread> workers;
        const size_t n = pointClouds.size();
        for (int g = 0; g < n_gpus; ++g) {
          const size_t lo = n * (size_t)g / (size_t)n_gpus, hi = n * (size_t)(g + 1) / (size_t)n_gpus;
          workers.emplace_back([&, g, lo, hi]() {
            try {
              if (hi <= lo) return;
              ism3d::ImplicitShapeModel* model = &ism;
              std::unique_ptr<ism3d::ImplicitShapeModel> own;
              if (g > 0) {  // rank 0 reuses the model already loaded on the first device
                own.reset(new ism3d::ImplicitShapeModel(device + g));
                own->setLogging(false);
                if (!own->readObject(ismFile)) { ok[(size_t)g] = 0; errors[(size_t)g] = "could not read " + ismFile; return; }
                model = own.get();
              }
              std::vector<std::string> part(pointClouds.begin() + (long)lo, pointClouds.begin() + (long)hi);
              std::vector<std::vector<ism3d::VotingMaximum>> res;
              if (!model->detectBatch(part, res, shard_times[(size_t)g], batch)) { ok[(size_t)g] = 0; return; }
              for (size_t i = 0; i < res.size(); ++i) all[lo + i] = std::move(res[i]);
            } catch (const std::exception& e) {
              ok[(size_t)g] = 0;
              errors[(size_t)g] = e.what();
            }
          });
        }
        for (auto& w : workers) w.join();
        for (int g = 0; g < n_gpus; ++g)
          if (!ok[(size_t)g]) {
            std::cerr << "classification failed on GPU " << device + g << ": " << errors[(size_t)g] << std::endl;
            return 1;
          }
        // the stage buckets of the GPUs overlap in time: report the slowest shard per bucket
        for (auto& st : shard_times)
          for (auto& kv : st) times[kv.first] = std::max(times[kv.first], kv.second);
      }
      int numCorrectClasses = 0, numCorrectInstances = 0;
      std::map<unsigned, std::pair<unsigned, unsigned>> acc;
      for (size_t i = 0; i < pointClouds.size(); ++i) {
        const auto& maxima = all[i];
        int classId = -1, instanceId = -1;
        if (!maxima.empty()) {  // eval_classification.cpp:412-427
          classId = (int)maxima[0].classId;
          instanceId = (int)maxima[0].instanceId;
          if (label_usage == LabelUsage::INSTANCE_PRIMARY) {
            instanceId = classId;
            classId = (int)instance_to_class_map[(unsigned)classId];
          }
        }
        if (summaryFile.is_open())
          summaryFile << "file: " << pointClouds[i] << ", ground truth class: " << gt_class_ids[i]
                      << ", classified class: " << classId << std::endl;
        auto& a = acc[gt_class_ids[i]];
        a.second++;
        if ((int)gt_class_ids[i] == classId) { numCorrectClasses++; a.first++; }
        if ((int)gt_instance_ids[i] == instanceId) numCorrectInstances++;
      }
      const double total_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
      float avg_pc_acc = 0;
      for (auto& e : acc) avg_pc_acc += (float)e.second.first / e.second.second;
      avg_pc_acc /= acc.size();
      const float accuracy = ((float)numCorrectClasses / pointClouds.size()) * 100.0f;
      if (summaryFile.is_open()) {
        summaryFile << "\n\nclass id to class name mapping:" << std::endl;
        for (auto& e : class_labels_rmap) summaryFile << e.first << ": " << e.second << std::endl;
        double time_sum = 0;
        for (auto& it : times) if (it.first != "complete") time_sum += it.second / 1000;
        summaryFile << "\n\n\ncomplete time: " << times["complete"] / 1000 << " [s], sum all steps: " << time_sum << " [s]\n";
        summaryFile << "times per step:\n";
        const std::pair<const char*, const char*> rows[] = {{"create flann index: ", "flann"}, {"compute normals:    ", "normals"},
                                                            {"compute keypoints:  ", "keypoints"}, {"compute features:   ", "features"},
                                                            {"cast votes:         ", "voting"}, {"find maxima:        ", "maxima"}};
        for (auto& r : rows) summaryFile << r.first << std::setw(10) << std::setfill(' ') << times[r.second] / 1000 << " [s]" << std::endl;
        summaryFile << "\n\n Accuracy: " << accuracy << " %, Average per Class Accuracy: " << avg_pc_acc * 100.0f << " %\n\n";
        summaryFile << " result: " << numCorrectClasses << " of " << pointClouds.size() << " clouds classified correctly (" << accuracy << " %)\n";
        summaryFile << " result: " << numCorrectInstances << " of " << pointClouds.size() << " instances recognized correctly ("
                    << ((float)numCorrectInstances / pointClouds.size()) * 100.0f << " %)\n\n";
        summaryFile << " Total processing time: " << total_s << " seconds \n";
      }
      std::cout << "Accuracy: " << accuracy << " % (" << numCorrectClasses << " of " << pointClouds.size() << "), "
                << pointClouds.size() / total_s << " clouds/s incl. file IO" << std::endl;
    }
  } catch (const ism3d::Exception& e) {
    std::cerr << e.what() << std::endl;
    return 1;
  } catch (const std::exception& e) {
    std::cerr << "an exception occurred: " << e.what() << std::endl;
    return 1;
  }
  return 0;
}
