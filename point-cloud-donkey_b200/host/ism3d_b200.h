// ism3d_b200.h — host-side C++ mirror of the reference's public interface for the classification hot path, on top of
// the C-ABI of include/pcdb200.h.  Same class / method names and argument meaning as
// src/implicit_shape_model/implicit_shape_model.h:82-253 so that eval_tool-style front-ends compile against it
// without PCL: readObject / writeObject (.ism JSON + .ismd boost archive, utils/json_object.cpp:41-178),
// addTrainingModel / train (implicit_shape_model.cpp:165-500), detect (:564-712), labels, getCodebook / getVoting.
// Point clouds are flat arrays (io::Cloud) instead of pcl::PointCloud<PointXYZRGBNormal>.
#pragma once
#include <cstdint>
#include <exception>
#include <map>
#include <memory>
#include <string>
#include <tuple>
#include <vector>

#include "../../include/pcdb200.h"
#include "io_formats.h"
#include "json_min.h"

namespace ism3d {

// utils/exception.h:12-80
class Exception : public std::exception {
 public:
  explicit Exception(const std::string& m) : msg_(m) {}
  const char* what() const noexcept override { return msg_.c_str(); }
 private:
  std::string msg_;
};
class RuntimeException : public Exception { public: using Exception::Exception; };
class BadParamException : public Exception { public: using Exception::Exception; };
class JSONException : public Exception { public: using Exception::Exception; };

namespace Utils {
struct BoundingBox {  // utils/utils.h:50-64
  float position[3] = {0, 0, 0};
  float rotQuat[4] = {1, 0, 0, 0};  // w x y z
  float size[3] = {0, 0, 0};
};
BoundingBox computeAABB(const io::Cloud& cloud);     // utils/utils.cpp:221-233
float computeCloudRadius(const io::Cloud& cloud);    // utils/utils.cpp:301-321
}  // namespace Utils

struct Vote {  // voting/voting_maximum.h:25-42
  float position[3];
  float weight;
  unsigned classId, instanceId;
  float keypoint[3], keypoint_training[3];
  Utils::BoundingBox boundingBox;
  int codewordId;
};

struct VotingMaximum {  // voting/voting_maximum.h:51-88
  float position[3] = {0, 0, 0};
  float weight = 0;
  unsigned classId = (unsigned)-1, instanceId = (unsigned)-1;
  float instanceWeight = 0;
  Utils::BoundingBox boundingBox;
  std::vector<Vote> votes;
  struct GlobalHypothesis { unsigned classId = (unsigned)-1; float classWeight = -1; unsigned instanceId = 0; float instanceWeight = 0; } globalHypothesis;
};

// Flat codebook: what Codebook + CodewordDistribution + Codeword hold (codebook/codebook.h, codeword_distribution.h)
class Codebook {
 public:
  int64_t getSize() const { return (int64_t)ids.size(); }
  int getDim() const { return dim; }
  bool isEmpty() const { return ids.empty(); }
  int dim = 0;
  std::vector<float> words;               // N x D, id order
  std::vector<int32_t> ids, numFeatures, classIds;
  std::vector<float> weights, keypoints;  // N, N x 3
  std::vector<int64_t> vote_off;          // N + 1
  std::vector<float> vote_xyz, vote_weight, vote_bbox, vote_class_weight;
  std::vector<uint32_t> vote_class, vote_instance;
  std::vector<int64_t> cw_off;            // per distribution: statistical class weights (class, weight)
  std::vector<int32_t> cw_class;
  std::vector<float> cw_weight;
  std::map<unsigned, float> classSigmas;  // sigma^2 per class
};

class Voting {
 public:
  const std::map<unsigned, std::vector<Vote>>& getVotes() const { return m_votes; }  // read by training_gui.cpp:1012
  std::map<unsigned, std::vector<Vote>> m_votes;
  std::map<unsigned, std::pair<float, float>> m_dimensions_map, m_variance_map;  // voting.cpp:497-557
};

class ImplicitShapeModel {
 public:
  explicit ImplicitShapeModel(int device = 0);
  ~ImplicitShapeModel();
  // What readObject(file, training = true) makes of a configuration file, WITHOUT creating a device context (host-only:
  // config checks and tests run where there is no GPU).  Throws what readObject throws for options outside the built path.
  static pcdb_params paramsOfConfigFile(const std::string& file, std::string* bounding_box_type = nullptr);
  ImplicitShapeModel(const ImplicitShapeModel&) = delete;

  // JSONObject (utils/json_object.h:49-66)
  bool readObject(std::string file, bool training = false);
  bool writeObject(std::string file);
  bool writeObject(std::string file, std::string fileData);

  // implicit_shape_model.h:86-253
  void clear();
  bool addTrainingModel(const std::string& filename, unsigned classId, unsigned instanceId);
  void train();
  bool detect(const std::string& filename, std::vector<VotingMaximum>& maxima, std::map<std::string, double>& times);
  std::tuple<std::vector<VotingMaximum>, std::map<std::string, double>> detect(const io::Cloud& points, bool hasNormals);
  // the throughput entry: the whole test list in batches through pcdb_classify_batch; per-cloud outputs unchanged
  bool detectBatch(const std::vector<std::string>& filenames, std::vector<std::vector<VotingMaximum>>& maxima,
                   std::map<std::string, double>& times, int batch = 256, bool with_votes = false);

  void setSignalsState(bool) {}  // boost::signals2 hooks of the GUI: not part of this path
  void setLogging(bool l) { m_logging = l; }
  void setOutputFilename(const std::string& f) { m_output_file_name = f; }
  void setLabels(const std::map<unsigned, std::string>& class_labels, const std::map<unsigned, std::string>& instance_labels,
                 const std::map<unsigned, unsigned>& instance_to_class_map);
  const std::map<unsigned, std::string>& getClassLabels() const { return m_class_labels; }
  const std::map<unsigned, std::string>& getInstanceLabels() const { return m_instance_labels; }
  const std::map<unsigned, unsigned>& getInstanceClassMap() const { return m_instance_to_class_map; }
  bool isInstancePrimaryLabel() const { return m_instance_labels_primary; }
  bool isUsingGlobalFeatures() const { return false; }
  // size hint per class for detection (implicit_shape_model.h:215-247): DistanceThresholdDetection, scaled by the class's
  // average object radius ("ObjectRadius") or median bounding-box dimension ("BoundingBoxMedian") learned in training
  std::map<unsigned, float> getDetectionThreshold() const;
  const Codebook* getCodebook() const { return &m_codebook; }
  const Voting* getVoting() const { return &m_voting; }
  const pcdb_params& params() const { return m_params; }
  pcdb_ctx* context() { return m_ctx; }

 private:
  struct NoDevice {};
  explicit ImplicitShapeModel(NoDevice);
  void configFromJson(const jsonmin::Value& objectConfig);
  void saveData(std::ostream& os) const;
  void loadData(std::istream& is);
  void uploadCodebook();
  void check(int rc) const;
  void log(const char* level, const std::string& msg) const;
  std::vector<VotingMaximum> toMaxima(const pcdb_maximum* mx, int64_t n, const std::vector<pcdb_vote>* votes,
                                      const std::vector<int64_t>* member_idx, const std::vector<float>* member_w) const;

  pcdb_ctx* m_ctx = nullptr;
  pcdb_params m_params;
  jsonmin::Value m_config;  // "ObjectConfig" as read (unknown keys are kept for writeObject)
  Codebook m_codebook;
  Voting m_voting;
  bool m_codebook_uploaded = false;
  bool m_logging = true;
  bool m_instance_labels_primary = true;
  float m_distance_detection_thresh = 0.05f;
  std::string m_distance_thresh_type = "Fixed";
  std::string m_bb_type = "MVBB", m_output_file_name, m_input_config_file;
  std::map<unsigned, std::string> m_class_labels, m_instance_labels;
  std::map<unsigned, unsigned> m_instance_to_class_map;
  std::map<std::string, double> m_processing_times;
  // training set (class id -> models), implicit_shape_model.cpp:165-210
  std::map<unsigned, std::vector<std::string>> m_training_objects_filenames;
  std::map<unsigned, std::vector<unsigned>> m_training_objects_instance_ids;
};

}  // namespace ism3d
