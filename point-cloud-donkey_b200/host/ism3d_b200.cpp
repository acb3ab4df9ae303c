// ism3d_b200.cpp — see ism3d_b200.h.  Every computation on the path is a call into libpcdb200 (CUDA); what runs here
// is IO, configuration and the reference's std::map bookkeeping around it.
#include "ism3d_b200.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <ctime>
#include <iostream>
#include <numeric>

namespace ism3d {

// ---- Utils ------------------------------------------------------------------------------------------------------
namespace Utils {
BoundingBox computeAABB(const io::Cloud& c) {
  BoundingBox b;
  float mn[3], mx[3];
  bool any = false;
  for (size_t i = 0; i < c.size(); ++i) {
    const float* p = &c.xyz[3 * i];
    if (!std::isfinite(p[0]) || !std::isfinite(p[1]) || !std::isfinite(p[2])) continue;
    for (int a = 0; a < 3; ++a) {
      mn[a] = any ? std::min(mn[a], p[a]) : p[a];
      mx[a] = any ? std::max(mx[a], p[a]) : p[a];
    }
    any = true;
  }
  if (!any) throw RuntimeException("computeAABB: cloud has no finite point");
  for (int a = 0; a < 3; ++a) {
    b.size[a] = mx[a] - mn[a];
    b.position[a] = mn[a] + (b.size[a] / 2);
  }
  return b;
}
float computeCloudRadius(const io::Cloud& c) {
  // pcl::compute3DCentroid accumulates in the scalar type of the output (float here), skipping non-finite points
  float acc[3] = {0, 0, 0};
  unsigned n = 0;
  for (size_t i = 0; i < c.size(); ++i) {
    const float* p = &c.xyz[3 * i];
    if (!std::isfinite(p[0]) || !std::isfinite(p[1]) || !std::isfinite(p[2])) continue;
    acc[0] += p[0]; acc[1] += p[1]; acc[2] += p[2];
    ++n;
  }
  if (n == 0) return 0.f;
  float cen[3] = {acc[0] / n, acc[1] / n, acc[2] / n};
  float radius = 0.f;
  for (size_t i = 0; i < c.size(); ++i) {
    float d0 = c.xyz[3 * i] - cen[0], d1 = c.xyz[3 * i + 1] - cen[1], d2 = c.xyz[3 * i + 2] - cen[2];
    float r = std::sqrt(d0 * d0 + d1 * d1 + d2 * d2);
    if (r > radius) radius = r;
  }
  return radius;
}
}  // namespace Utils

namespace {
// boost::math::quaternion<float> arithmetic (w x y z) in the reference's operation order (utils/utils.cpp:136-178,
// 342-394,560-574); this file is compiled with -ffp-contract=off
struct Quat { float a, b, c, d; };
Quat qmul(const Quat& l, const Quat& r) {
  Quat o;
  o.a = +l.a * r.a - l.b * r.b - l.c * r.c - l.d * r.d;
  o.b = +l.a * r.b + l.b * r.a + l.c * r.d - l.d * r.c;
  o.c = +l.a * r.c - l.b * r.d + l.c * r.a + l.d * r.b;
  o.d = +l.a * r.d + l.b * r.c - l.c * r.b + l.d * r.a;
  return o;
}
Quat qconj(const Quat& q) { return Quat{q.a, -q.b, -q.c, -q.d}; }
Quat lrf_quat(const float* rf) {
  float m[3][3] = {{rf[0], rf[1], rf[2]}, {rf[3], rf[4], rf[5]}, {rf[6], rf[7], rf[8]}};
  float quat[4] = {0, 0, 0, 0};
  float trace = m[0][0] + m[1][1] + m[2][2];
  float root;
  if (trace > 0.0f) {
    root = sqrtf(trace + 1.0f);
    quat[3] = 0.5f * root;
    root = 0.5f / root;
    quat[0] = (m[2][1] - m[1][2]) * root;
    quat[1] = (m[0][2] - m[2][0]) * root;
    quat[2] = (m[1][0] - m[0][1]) * root;
  } else {
    static const size_t next[3] = {1, 2, 0};
    size_t i = 0;
    if (m[1][1] > m[0][0]) i = 1;
    if (m[2][2] > m[i][i]) i = 2;
    size_t j = next[i], k = next[j];
    root = sqrtf(m[i][i] - m[j][j] - m[k][k] + 1.0);
    quat[i] = 0.5f * root;
    root = 0.5f / root;
    quat[3] = (m[k][j] - m[j][k]) * root;
    quat[j] = (m[j][i] + m[i][j]) * root;
    quat[k] = (m[k][i] + m[i][k]) * root;
  }
  return Quat{quat[3], quat[0], quat[1], quat[2]};
}
void quat_rotate(const Quat& q, const float* p, float* out) {
  Quat t = qmul(qmul(q, Quat{0, p[0], p[1], p[2]}), qconj(q));
  out[0] = t.b; out[1] = t.c; out[2] = t.d;
}
void quat_rotate_inv(const Quat& q, const float* p, float* out) {
  Quat t = qmul(qmul(qconj(q), Quat{0, p[0], p[1], p[2]}), q);
  out[0] = t.b; out[1] = t.c; out[2] = t.d;
}

std::string dirname_of(const std::string& f) {
  size_t p = f.find_last_of('/');
  return p == std::string::npos ? std::string() : f.substr(0, p + 1);
}
bool has_extension(const std::string& f) {
  size_t s = f.find_last_of('/');
  size_t d = f.find_last_of('.');
  return d != std::string::npos && (s == std::string::npos || d > s);
}
}  // namespace

// ---- lifecycle ----------------------------------------------------------------------------------------------------
ImplicitShapeModel::ImplicitShapeModel(int device) {
  pcdb_default_params(&m_params);
  int rc = pcdb_create(&m_ctx, device);
  if (rc != PCDB_OK) throw RuntimeException(std::string("pcdb_create failed: ") + pcdb_last_error(nullptr));
  m_processing_times = {{"complete", 0}, {"features", 0}, {"keypoints", 0}, {"normals", 0},
                        {"flann", 0},    {"voting", 0},   {"maxima", 0}};  // implicit_shape_model.cpp:160
}
ImplicitShapeModel::ImplicitShapeModel(NoDevice) { pcdb_default_params(&m_params); }
ImplicitShapeModel::~ImplicitShapeModel() {
  if (m_ctx) pcdb_destroy(m_ctx);
}

pcdb_params ImplicitShapeModel::paramsOfConfigFile(const std::string& file, std::string* bounding_box_type) {
  ImplicitShapeModel m{NoDevice{}};
  m.setLogging(false);
  jsonmin::Value cfg = jsonmin::parse_file(file);
  if (!cfg.isObject() || !cfg.isMember("ObjectConfig")) throw RuntimeException("no ObjectConfig in " + file);
  m.configFromJson(cfg["ObjectConfig"]);
  if (bounding_box_type) *bounding_box_type = m.m_bb_type;
  return m.m_params;
}

void ImplicitShapeModel::check(int rc) const {
  if (rc == PCDB_OK) return;
  std::string m = pcdb_last_error(m_ctx);
  if (rc == PCDB_E_INVALID) throw BadParamException(m);
  throw RuntimeException(m);
}
void ImplicitShapeModel::log(const char* level, const std::string& msg) const {
  if (!m_logging && std::string(level) == "INFO") return;
  std::time_t t = std::time(nullptr);
  char buf[16];
  std::strftime(buf, sizeof(buf), "%H:%M:%S", std::localtime(&t));
  std::cerr << "[" << buf << "] " << level << ": " << msg << std::endl;  // log4cxx pattern of :82-89
}
void ImplicitShapeModel::clear() {
  m_training_objects_filenames.clear();
  m_training_objects_instance_ids.clear();
  m_codebook = Codebook();
  m_codebook_uploaded = false;
}

// ---- configuration ---------------------------------------------------------------------------------------------------
namespace {
// JSONParameter semantics (utils/json_parameter.h:26-33, json_parameter_base.cpp:35-45): missing -> default (+ warning),
// wrong JSON type -> JSONException
struct ParamReader {
  const jsonmin::Value& p;
  const ImplicitShapeModel* owner;
  std::vector<std::string>* warnings;
  double num(const char* name, double def) const {
    const jsonmin::Value* v = p.find(name);
    if (!v) { warnings->push_back(std::string("parameter \"") + name + "\" not found, using default"); return def; }
    if (v->isBool()) return v->b ? 1 : 0;
    if (!v->isNumber()) throw JSONException(std::string("invalid JSON type for parameter ") + name);
    return v->num;
  }
  bool boolean(const char* name, bool def) const {
    const jsonmin::Value* v = p.find(name);
    if (!v) { warnings->push_back(std::string("parameter \"") + name + "\" not found, using default"); return def; }
    if (!v->isBool()) throw JSONException(std::string("invalid JSON type for parameter ") + name);
    return v->b;
  }
  std::string str(const char* name, const std::string& def) const {
    const jsonmin::Value* v = p.find(name);
    if (!v) { warnings->push_back(std::string("parameter \"") + name + "\" not found, using default"); return def; }
    if (!v->isString()) throw JSONException(std::string("invalid JSON type for parameter ") + name);
    return v->str;
  }
};
std::string type_of(const jsonmin::Value& child, const char* what) {
  if (!child.isObject() || !child["Type"].isString())
    throw BadParamException(std::string("config has no valid \"") + what + "\" child (implicit_shape_model.cpp:1095-1103)");
  return child["Type"].str;
}
}  // namespace

void ImplicitShapeModel::configFromJson(const jsonmin::Value& oc) {
  std::vector<std::string> warn;
  m_config = oc;
  pcdb_params P;
  pcdb_default_params(&P);
  ParamReader top{oc["Parameters"], this, &warn};
  P.normal_radius = (float)top.num("NormalRadius", 0.05);                       // implicit_shape_model.cpp:110
  P.consistent_normals_method = (int)top.num("ConsistentNormalsMethod", 2);      // :112
  if (P.consistent_normals_method < 0 || P.consistent_normals_method > 2)
    throw BadParamException("ConsistentNormalsMethod must be 0, 1 or 2 (3 needs the optional VCG library)");
  const std::string dist = top.str("DistanceType", "Euclidean");
  if (dist == "Euclidean") P.distance_type = PCDB_DIST_EUCLIDEAN;
  else if (dist == "ChiSquared") P.distance_type = PCDB_DIST_CHISQUARED;
  else throw RuntimeException("invalid distance type: " + dist);  // implicit_shape_model.cpp:1245-1252
  m_bb_type = top.str("BoundingBoxType", "MVBB");
  m_instance_labels_primary = top.boolean("InstanceLabelsPrimary", true);
  m_distance_detection_thresh = (float)top.num("DistanceThresholdDetection", 0.05);
  m_distance_thresh_type = top.str("DistanceThresholdType", "Fixed");
  if (top.boolean("SingleObjectMode", false))
    throw RuntimeException("The parameter for \"single object mode\" must be set inside the \"Voting\" section of the config file. You are using the \"Parameters\" section.");
  if (!top.boolean("FLANNExactMatch", false))
    log("INFO", "FLANNExactMatch=false: this implementation always runs the exact search (the approximate kd-forest "
                "is not reproducible; SURVEY.md section 0-4)");
  // options outside the built hot path are collected and reported together
  std::vector<std::string> outside;
  for (const char* k : {"UseSmoothing", "UseStatisticalOutlierRemoval", "UseRadiusOutlierRemoval", "UseVoxelFiltering", "UseSvmTraining"})
    if (oc["Parameters"].isMember(k) && top.boolean(k, false)) outside.push_back(std::string("Parameters.") + k + "=true");

  const jsonmin::Value& ch = oc["Children"];
  // Features
  const jsonmin::Value& feat = ch["Features"];
  const std::string ftype = type_of(feat, "Features");
  if (ftype == "SHOT") P.feature_type = PCDB_FEATURE_SHOT;
  else if (ftype == "CSHOT") P.feature_type = PCDB_FEATURE_CSHOT;
  else throw BadParamException("Features.Type \"" + ftype + "\" is not built on this path (SHOT and CSHOT are)");
  ParamReader fr{feat["Parameters"], this, &warn};
  P.feature_radius = fr.num("Radius", 0.1);
  P.lrf_radius = (double)(float)fr.num("ReferenceFrameRadius", 0.2);  // float member promoted to double
  const std::string lrf_type = fr.str("ReferenceFrameType", "SHOT");
  if (lrf_type != "SHOT") throw BadParamException("invalid reference frame type for this path: " + lrf_type);
  // Keypoints
  const jsonmin::Value& kp = ch["Keypoints"];
  if (type_of(kp, "Keypoints") != "VoxelGrid") throw BadParamException("Keypoints.Type must be \"VoxelGrid\" on this path");
  P.leaf_size = (float)ParamReader{kp["Parameters"], this, &warn}.num("LeafSize", 0.1);
  // Clustering / ranking: identity behaviour only
  if (type_of(ch["Clustering"], "Clustering") != "None") throw BadParamException("Clustering.Type must be \"None\" on this path");
  if (type_of(ch["FeatureWeighting"], "FeatureWeighting") != "Uniform") throw BadParamException("FeatureWeighting.Type must be \"Uniform\" on this path");
  // Codebook + activation
  const jsonmin::Value& cb = ch["Codebook"];
  if (!cb.isObject()) throw BadParamException("config has no \"Codebook\" child");
  ParamReader cr{cb["Parameters"], this, &warn};
  P.use_class_weight = cr.boolean("UseClassWeight", false);
  P.use_vote_weight = cr.boolean("UseVoteWeight", false);
  P.use_matching_weight = cr.boolean("UseMatchingWeight", false);
  P.use_codeword_weight = cr.boolean("UseCodewordWeight", false);
  if (cb["Parameters"].isMember("UsePartialShot") && cr.boolean("UsePartialShot", false))
    throw BadParamException("UsePartialShot=true is outside the built hot path");
  const jsonmin::Value& act = cb["Children"]["ActivationStrategy"];
  if (type_of(act, "ActivationStrategy") != "KNN") throw BadParamException("ActivationStrategy.Type must be \"KNN\" on this path");
  ParamReader ar{act["Parameters"], this, &warn};
  P.knn_k = (int)ar.num("K", 1);
  P.use_distance_ratio = ar.boolean("UseDistanceRatio", false);
  P.distance_ratio_threshold = (float)ar.num("DistanceRatioThreshold", 0.95);
  // Voting
  const jsonmin::Value& vo = ch["Voting"];
  if (type_of(vo, "Voting") != "MeanShift") throw BadParamException("Voting.Type must be \"MeanShift\" on this path");
  ParamReader vr{vo["Parameters"], this, &warn};
  P.bandwidth = (float)vr.num("Bandwidth", 0.2);
  P.ms_threshold = (float)vr.num("Threshold", 1e-3);
  P.ms_max_iter = (int)vr.num("MaxIter", 1000);
  const std::string kernel = vr.str("Kernel", "Gaussian");
  if (kernel == "Gaussian") P.ms_kernel = PCDB_KERNEL_GAUSSIAN;
  else if (kernel == "Uniform") P.ms_kernel = PCDB_KERNEL_UNIFORM;
  else throw BadParamException("invalid Voting.Kernel: " + kernel);
  const std::string sup = vr.str("MaximaSuppression", "Average");
  if (sup == "Average") P.maxima_suppression = PCDB_SUPPRESS_AVERAGE;
  else if (sup == "Suppress") P.maxima_suppression = PCDB_SUPPRESS_SUPPRESS;
  else throw BadParamException("Voting.MaximaSuppression \"" + sup + "\" is not built on this path");
  P.min_threshold = (float)vr.num("MinThreshold", 0.0);
  P.min_votes_threshold = (int)vr.num("MinVotesThreshold", 1);
  P.best_k = (int)vr.num("BestK", -1);
  P.average_rotation = vr.boolean("AverageRotation", false);
  P.single_object_mode = vr.boolean("SingleObjectMode", false);
  // MaximaHandler::getSearchDistForClass (maxima_handler.cpp:509-521); an unknown type logs an error and uses the config value
  const std::string rtype = vr.str("BinOrBandwidthType", "Config");
  if (rtype == "Config" || rtype == "Fixed") P.radius_type = PCDB_RADIUS_CONFIG;
  else if (rtype == "FirstDim" || rtype == "ObjectRadius") P.radius_type = PCDB_RADIUS_FIRST_DIM;
  else if (rtype == "SecondDim" || rtype == "BoundingBoxMedian") P.radius_type = PCDB_RADIUS_SECOND_DIM;
  else {
    log("ERROR", "Invalid radius type: " + rtype + "! Using config value instead.");
    P.radius_type = PCDB_RADIUS_CONFIG;
  }
  P.radius_factor = (float)vr.num("BinOrBandwidthFactor", 1.0);
  // MaximaHandler::setSingleObjectMaxType (maxima_handler.h:42-58)
  const std::string mtype = vr.str("SingleObjectMaxType", "Default");
  if (mtype == "None" || mtype == "Default") P.single_object_max_type = PCDB_SOMAX_DEFAULT;
  else if (mtype == "BandwidthVotes") P.single_object_max_type = PCDB_SOMAX_BANDWIDTH;
  else if (mtype == "VotingSpaceVotes") P.single_object_max_type = PCDB_SOMAX_VOTING_SPACE;
  else if (mtype == "ModelRadiusVotes") P.single_object_max_type = PCDB_SOMAX_MODEL_RADIUS;
  else {
    log("WARN", "Invalid single object maximum type: " + mtype + "! Using default instead.");
    P.single_object_max_type = PCDB_SOMAX_DEFAULT;
  }
  // MaximaHandler::filterMaxima (maxima_handler.cpp:272-295): an unknown type logs an error and filters nothing
  const std::string ftr = vr.str("MaxFilterType", "None");
  if (ftr == "None") P.max_filter_type = PCDB_MAXFILTER_NONE;
  else if (ftr == "Simple") P.max_filter_type = PCDB_MAXFILTER_SIMPLE;
  else if (ftr == "Merge") P.max_filter_type = PCDB_MAXFILTER_MERGE;
  else {
    log("ERROR", "Invalid maxima filter type specified: " + ftr + "! No filtering is performed!");
    P.max_filter_type = PCDB_MAXFILTER_NONE;
  }
  if (vr.boolean("UseGlobalFeatures", false))
    outside.push_back("Voting.UseGlobalFeatures=true (global-descriptor classifiers: GlobalFeatures child, SVM / KNN merge of "
                      "global and local hypotheses, voting.cpp:217-300)");
  // Voting::filterVotesWithRansac (voting.cpp:47-50,110-127,356-433)
  P.ransac_vote_filtering = vr.boolean("RansacVoteFiltering", false);
  P.ransac_refine_model = vr.boolean("RansacRefineModel", false);
  P.ransac_inlier_threshold = (float)vr.num("RansacInlierThreshold", 0.1);
  const std::string rth = vr.str("RansacInlierThresholdType", "Fixed");
  if (rth == "ObjectRadius") P.ransac_threshold_type = PCDB_RANSAC_OBJECT_RADIUS;
  else if (rth == "BoundingBoxMedian") P.ransac_threshold_type = PCDB_RANSAC_BBOX_MEDIAN;
  else P.ransac_threshold_type = PCDB_RANSAC_FIXED;  // voting.cpp:112-122: anything else keeps the configured value
  if (P.ransac_vote_filtering && P.ransac_refine_model)
    outside.push_back("Voting.RansacRefineModel=true (PCL's SampleConsensus::refineModel, voting.cpp:363)");
  if (!outside.empty()) {
    std::string msg = "this configuration asks for parts of the reference outside the built hot path; set them to false: ";
    for (size_t i = 0; i < outside.size(); ++i) msg += (i ? "; " : "") + outside[i];
    throw BadParamException(msg);
  }
  for (const std::string& w : warn) log("WARN", w);
  m_params = P;
  if (m_ctx) check(pcdb_set_params(m_ctx, &m_params));
}

std::map<unsigned, float> ImplicitShapeModel::getDetectionThreshold() const {
  std::map<unsigned, float> out;
  for (const auto& it : m_voting.m_dimensions_map) {
    float v = m_distance_detection_thresh;
    if (m_distance_thresh_type == "ObjectRadius") v *= it.second.first;
    if (m_distance_thresh_type == "BoundingBoxMedian") v *= it.second.second;
    out.insert({it.first, v});
  }
  return out;
}

// ---- .ism / .ismd ----------------------------------------------------------------------------------------------------
bool ImplicitShapeModel::readObject(std::string file, bool training) {
  log("INFO", "reading object configuration from file: " + file);
  m_input_config_file = file;
  jsonmin::Value cfg;
  try {
    cfg = jsonmin::parse_file(file);
  } catch (const std::exception& e) {
    log("ERROR", e.what());
    return false;
  }
  if (cfg.isNull()) { log("ERROR", "Json Config is NULL!"); return false; }
  if (cfg.isMember("ObjectConfig")) configFromJson(cfg["ObjectConfig"]);
  if (cfg.isMember("ObjectData") && !training) {
    if (!cfg["ObjectData"].isString()) return false;
    std::string fileData = dirname_of(file) + cfg["ObjectData"].str;
    std::ifstream ifs(fileData, std::ios::binary);
    if (!ifs) { log("ERROR", "Error opening file: " + fileData); return false; }
    try {
      loadData(ifs);
    } catch (const std::exception& e) {
      log("ERROR", std::string("could not load child objects: ") + e.what());
      return false;
    }
  } else if (!training) {
    log("ERROR", "Config file " + file + " has only parameters, but no trained object data!");
    return false;
  }
  log("INFO", "reading successful");
  return true;
}

bool ImplicitShapeModel::writeObject(std::string file) {
  if (!has_extension(file)) file += ".ism";
  return writeObject(file, file + "d");
}

bool ImplicitShapeModel::writeObject(std::string file, std::string fileData) {
  log("INFO", "writing object to files: " + file + ", " + fileData);
  size_t pos = fileData.find_last_of('/');
  std::string noPath = pos != std::string::npos ? fileData.substr(pos + 1) : fileData;
  jsonmin::Value cfg = jsonmin::Value::object();
  cfg.set("ObjectConfig", m_config);
  cfg.set("ObjectData", jsonmin::Value::of(noPath));
  std::ofstream ofs(fileData, std::ios::binary);
  if (!ofs) return false;
  saveData(ofs);
  ofs.close();
  if (!jsonmin::write_file(cfg, file)) return false;
  log("INFO", "writing successful");
  return true;
}

void ImplicitShapeModel::saveData(std::ostream& os) const {  // implicit_shape_model.cpp:1144-1179
  io::BinaryOArchive oa(os);
  oa.put<uint32_t>((uint32_t)m_instance_to_class_map.size());
  for (auto& it : m_instance_to_class_map) { oa.put<uint32_t>(it.first); oa.put<uint32_t>(it.second); }
  // Codebook::iSaveData codebook.cpp:740-761
  const Codebook& c = m_codebook;
  const int64_t N = c.getSize();
  oa.put<int32_t>((int32_t)N);
  for (int64_t i = 0; i < N; ++i) {
    // Codeword::iSaveData codeword.cpp:71-83
    oa.put<int32_t>(c.ids[i]);
    oa.put<int32_t>(c.numFeatures[i]);
    oa.put<float>(c.weights[i]);
    oa.put_vector(std::vector<float>(c.words.begin() + i * c.dim, c.words.begin() + (i + 1) * c.dim));
    oa.put<int32_t>(c.classIds[i]);
    for (int a = 0; a < 3; ++a) oa.put<float>(c.keypoints[3 * i + a]);
    // CodewordDistribution::iSaveData codeword_distribution.cpp:349-393
    const int64_t v0 = c.vote_off[i], v1 = c.vote_off[i + 1];
    oa.put<int32_t>((int32_t)(v1 - v0));
    for (int64_t v = v0; v < v1; ++v) for (int a = 0; a < 3; ++a) oa.put<float>(c.vote_xyz[3 * v + a]);
    oa.put_vector(std::vector<float>(c.vote_weight.begin() + v0, c.vote_weight.begin() + v1));
    oa.put_vector(std::vector<uint32_t>(c.vote_class.begin() + v0, c.vote_class.begin() + v1));
    oa.put_vector(std::vector<uint32_t>(c.vote_instance.begin() + v0, c.vote_instance.begin() + v1));
    const int64_t w0 = c.cw_off[i], w1 = c.cw_off[i + 1];
    oa.put<int32_t>((int32_t)(w1 - w0));
    for (int64_t w = w0; w < w1; ++w) { oa.put<int32_t>(c.cw_class[w]); oa.put<float>(c.cw_weight[w]); }
    oa.put<int32_t>((int32_t)(v1 - v0));
    for (int64_t v = v0; v < v1; ++v) for (int a = 0; a < 7; ++a) oa.put<float>(c.vote_bbox[7 * v + a]);
  }
  oa.put<int32_t>((int32_t)c.classSigmas.size());
  for (auto& it : c.classSigmas) { oa.put<int32_t>((int32_t)it.first); oa.put<float>(it.second); }
  // ActivationStrategy, Keypoints, Features, GlobalFeatures, Clustering: nothing (json_object.cpp:256-265)
  // Voting::iSaveData voting.cpp:559-614
  oa.put<uint32_t>((uint32_t)m_voting.m_dimensions_map.size());
  for (auto& it : m_voting.m_dimensions_map) { oa.put<uint32_t>(it.first); oa.put<float>(it.second.first); oa.put<float>(it.second.second); }
  oa.put<uint32_t>((uint32_t)m_voting.m_variance_map.size());
  for (auto& it : m_voting.m_variance_map) { oa.put<uint32_t>(it.first); oa.put<float>(it.second.first); oa.put<float>(it.second.second); }
  oa.put<uint32_t>(0);  // global features: none on this path (GlobalFeatures.Type "Dummy")
  // FeatureRanking: nothing; label maps
  oa.put<uint32_t>((uint32_t)m_class_labels.size());
  for (auto& it : m_class_labels) oa.put_string(it.second);
  oa.put<uint32_t>((uint32_t)m_instance_labels.size());
  for (auto& it : m_instance_labels) oa.put_string(it.second);
}

void ImplicitShapeModel::loadData(std::istream& is) {  // implicit_shape_model.cpp:1181-1237
  io::BinaryIArchive ia(is);
  m_instance_to_class_map.clear();
  for (uint32_t n = ia.get<uint32_t>(), i = 0; i < n; ++i) {
    uint32_t a = ia.get<uint32_t>(), b = ia.get<uint32_t>();
    m_instance_to_class_map.insert({a, b});
  }
  struct Entry {
    int32_t id, numFeatures, classId;
    float weight, kp[3];
    std::vector<float> data, votes, vweights, bbox;
    std::vector<uint32_t> cls, inst;
    std::vector<std::pair<int32_t, float>> cw;
  };
  std::map<int, Entry> dist;  // keyed by stored id (codebook.cpp:857-859: id order)
  const int32_t n_dist = ia.get<int32_t>();
  log("INFO", "Loading codebook with size: " + std::to_string(n_dist));
  for (int32_t i = 0; i < n_dist; ++i) {
    Entry e;
    e.id = ia.get<int32_t>();
    e.numFeatures = ia.get<int32_t>();
    e.weight = ia.get<float>();
    e.data = ia.get_vector_f();
    e.classId = ia.get<int32_t>();
    for (int a = 0; a < 3; ++a) e.kp[a] = ia.get<float>();
    const int32_t nv = ia.get<int32_t>();
    e.votes.resize((size_t)nv * 3);
    for (float& v : e.votes) v = ia.get<float>();
    e.vweights = ia.get_vector_f();
    e.cls = ia.get_vector_u();
    e.inst = ia.get_vector_u();
    for (int32_t nw = ia.get<int32_t>(), w = 0; w < nw; ++w) {
      int32_t c = ia.get<int32_t>();
      float f = ia.get<float>();
      e.cw.push_back({c, f});
    }
    const int32_t nb = ia.get<int32_t>();
    e.bbox.resize((size_t)nb * 7);
    for (float& v : e.bbox) v = ia.get<float>();
    if ((int)e.cls.size() != nv || (int)e.inst.size() != nv || (int)e.vweights.size() != nv || nb != nv)
      throw RuntimeException("inconsistent codeword distribution in .ismd");
    dist[e.id] = std::move(e);
  }
  Codebook c;
  c.vote_off.push_back(0);
  c.cw_off.push_back(0);
  for (auto& kv : dist) {
    Entry& e = kv.second;
    if (c.dim == 0) c.dim = (int)e.data.size();
    if ((int)e.data.size() != c.dim) throw RuntimeException("codewords of different lengths in .ismd");
    c.ids.push_back(e.id);
    c.numFeatures.push_back(e.numFeatures);
    c.weights.push_back(e.weight);
    c.classIds.push_back(e.classId);
    c.keypoints.insert(c.keypoints.end(), e.kp, e.kp + 3);
    c.words.insert(c.words.end(), e.data.begin(), e.data.end());
    c.vote_xyz.insert(c.vote_xyz.end(), e.votes.begin(), e.votes.end());
    c.vote_weight.insert(c.vote_weight.end(), e.vweights.begin(), e.vweights.end());
    c.vote_class.insert(c.vote_class.end(), e.cls.begin(), e.cls.end());
    c.vote_instance.insert(c.vote_instance.end(), e.inst.begin(), e.inst.end());
    c.vote_bbox.insert(c.vote_bbox.end(), e.bbox.begin(), e.bbox.end());
    for (uint32_t cl : e.cls) {  // statistical class weight of the vote's class (1 with a warning if missing, :92-99)
      float w = 1.0f;
      for (auto& p : e.cw) if ((uint32_t)p.first == cl) w = p.second;
      c.vote_class_weight.push_back(w);
    }
    for (auto& p : e.cw) { c.cw_class.push_back(p.first); c.cw_weight.push_back(p.second); }
    c.vote_off.push_back((int64_t)c.vote_class.size());
    c.cw_off.push_back((int64_t)c.cw_class.size());
  }
  for (int32_t n = ia.get<int32_t>(), i = 0; i < n; ++i) {
    int32_t cl = ia.get<int32_t>();
    float s = ia.get<float>();
    c.classSigmas[(unsigned)cl] = s;
  }
  m_codebook = std::move(c);
  m_codebook_uploaded = false;
  // Voting::iLoadData voting.cpp:616-734
  m_voting.m_dimensions_map.clear();
  m_voting.m_variance_map.clear();
  for (uint32_t n = ia.get<uint32_t>(), i = 0; i < n; ++i) {
    uint32_t cl = ia.get<uint32_t>();
    float a = ia.get<float>(), b = ia.get<float>();
    m_voting.m_dimensions_map.insert({cl, {a, b}});
  }
  for (uint32_t n = ia.get<uint32_t>(), i = 0; i < n; ++i) {
    uint32_t cl = ia.get<uint32_t>();
    float a = ia.get<float>(), b = ia.get<float>();
    m_voting.m_variance_map.insert({cl, {a, b}});
  }
  // global features are always fully deserialised (and ignored: UseGlobalFeatures=false on this path)
  for (uint32_t n = ia.get<uint32_t>(), i = 0; i < n; ++i) {
    ia.get<uint32_t>();
    for (uint32_t nc = ia.get<uint32_t>(), j = 0; j < nc; ++j)
      for (uint32_t nf = ia.get<uint32_t>(), k = 0; k < nf; ++k) {
        for (int r = 0; r < 9; ++r) ia.get<float>();
        ia.get_vector_f();
        ia.get<float>();
        ia.get<uint32_t>();
      }
  }
  m_class_labels.clear();
  for (uint32_t n = ia.get<uint32_t>(), i = 0; i < n; ++i) m_class_labels.insert({i, ia.get_string()});
  m_instance_labels.clear();
  for (uint32_t n = ia.get<uint32_t>(), i = 0; i < n; ++i) m_instance_labels.insert({i, ia.get_string()});
}

void ImplicitShapeModel::setLabels(const std::map<unsigned, std::string>& cl, const std::map<unsigned, std::string>& il,
                                   const std::map<unsigned, unsigned>& icm) {
  m_class_labels = cl;
  m_instance_labels = il;
  m_instance_to_class_map = icm;
}

void ImplicitShapeModel::uploadCodebook() {
  const Codebook& c = m_codebook;
  if (c.isEmpty()) throw RuntimeException("the codebook is empty (train or load a model first)");
  unsigned max_class = 0;
  for (uint32_t v : c.vote_class) max_class = std::max(max_class, v);
  for (auto& it : c.classSigmas) max_class = std::max(max_class, it.first);
  const int n_classes = (int)max_class + 1;
  std::vector<float> sigma((size_t)n_classes, 1.0f);  // missing sigma => 1 (codeword_distribution.cpp:117-121)
  for (auto& it : c.classSigmas) sigma[it.first] = it.second;
  check(pcdb_set_codebook(m_ctx, c.words.data(), c.getSize(), c.dim, c.vote_off.data(), c.vote_xyz.data(),
                          c.vote_weight.data(), c.vote_class.data(), c.vote_instance.data(), c.vote_bbox.data(),
                          c.vote_class_weight.data(), c.keypoints.data(), c.ids.data(), c.weights.data(), sigma.data(),
                          n_classes, 0));
  // Voting::iLoadData -> MaximaHandler::setBoundingBoxMaps (voting.cpp:619-650): the table behind BinOrBandwidthType
  std::vector<float> d1((size_t)n_classes, 0.f), d2((size_t)n_classes, 0.f);
  for (const auto& it : m_voting.m_dimensions_map)
    if ((int)it.first < n_classes) {
      d1[it.first] = it.second.first;
      d2[it.first] = it.second.second;
    }
  check(pcdb_set_class_dimensions(m_ctx, d1.data(), d2.data(), n_classes));
  m_codebook_uploaded = true;
}

// ---- training ----------------------------------------------------------------------------------------------------------
bool ImplicitShapeModel::addTrainingModel(const std::string& filename, unsigned classId, unsigned instanceId) {
  log("INFO", "adding training model with class id " + std::to_string(classId) + " and instance id " + std::to_string(instanceId));
  m_training_objects_filenames[classId].push_back(filename);
  m_training_objects_instance_ids[classId].push_back(instanceId);
  return true;
}

void ImplicitShapeModel::train() {
  m_codebook = Codebook();
  m_codebook_uploaded = false;
  if (m_training_objects_filenames.empty()) { log("WARN", "no training models found"); return; }
  if (m_bb_type != "AABB" && m_bb_type != "MVBB") throw BadParamException("invalid bounding box type: " + m_bb_type);
  if (m_bb_type == "MVBB") {
    // Utils::computeMVBB (utils/utils.cpp:242-293) is libgdiam's approximate minimum-volume box, a training-side
    // dependency that is not built.  Decision: train with the axis-aligned box (the reference's other BoundingBoxType),
    // say so, and record "AABB" in the model that is written, so that the saved .ism describes what its votes were
    // computed from.  Labels are unaffected in practice (votes still point at a box centre); vote vectors and stored
    // box sizes differ from an MVBB-trained model.
    log("WARN", "BoundingBoxType MVBB (libgdiam approximate minimum-volume box) is not built: training uses AABB and the "
                "written model records BoundingBoxType = AABB");
    m_bb_type = "AABB";
    if (m_config.isObject() && m_config.isMember("Parameters")) {
      jsonmin::Value params = m_config["Parameters"];
      params.set("BoundingBoxType", jsonmin::Value::of(std::string("AABB")));
      m_config.set("Parameters", params);
    }
  }
  const int D = m_params.feature_type == PCDB_FEATURE_CSHOT ? PCDB_CSHOT_DIM : PCDB_SHOT_DIM;
  // per class (std::map order), per model: features through the GPU path
  std::vector<float> fxyz, flrf, fdesc;
  std::vector<int64_t> foff{0};
  std::vector<unsigned> cloud_class, cloud_inst;
  std::vector<Utils::BoundingBox> boxes;
  std::map<unsigned, std::vector<Utils::BoundingBox>> class_boxes;
  std::map<unsigned, std::vector<float>> class_radii;
  for (auto& it : m_training_objects_filenames) {
    const unsigned class_id = it.first;
    log("INFO", "training class " + std::to_string(class_id) + " with " + std::to_string(it.second.size()) + " objects");
    for (size_t j = 0; j < it.second.size(); ++j) {
      io::Cloud cloud;
      std::string err;
      if (!io::load_cloud(it.second[j], cloud, err)) throw RuntimeException("could not load training model: " + err);
      if (cloud.size() == 0) throw RuntimeException("point cloud is empty: " + it.second[j]);
      // implicit_shape_model.cpp:374-390: clouds without (usable) normals get them estimated
      const bool has_normals = cloud.has_normals && !(cloud.normals[0] == 0 && cloud.normals[1] == 0 && cloud.normals[2] == 0) &&
                               !std::isnan(cloud.normals[0]);
      Utils::BoundingBox bb = Utils::computeAABB(cloud);
      const int64_t off[2] = {0, (int64_t)cloud.size()};
      const int64_t cap = (int64_t)cloud.size();
      std::vector<float> x((size_t)cap * 3), l((size_t)cap * 9), d((size_t)cap * D);
      int64_t out_off[2];
      check(pcdb_compute_features(m_ctx, cloud.xyz.data(), has_normals ? cloud.normals.data() : nullptr,
                                  cloud.has_rgb ? cloud.rgb.data() : nullptr, off, 1, x.data(), l.data(), d.data(), out_off,
                                  cap));
      const int64_t nf = out_off[1];
      fxyz.insert(fxyz.end(), x.begin(), x.begin() + nf * 3);
      flrf.insert(flrf.end(), l.begin(), l.begin() + nf * 9);
      fdesc.insert(fdesc.end(), d.begin(), d.begin() + nf * D);
      foff.push_back(foff.back() + nf);
      cloud_class.push_back(class_id);
      cloud_inst.push_back(m_training_objects_instance_ids[class_id][j]);
      boxes.push_back(bb);
      class_boxes[class_id].push_back(bb);
      class_radii[class_id].push_back(Utils::computeCloudRadius(cloud));
    }
  }
  // Voting::forwardBoxesAndRadii voting.cpp:497-557
  m_voting.m_dimensions_map.clear();
  m_voting.m_variance_map.clear();
  for (auto& it : class_boxes) {
    float med_sum = 0, med_sq = 0, rad_sum = 0, rad_sq = 0;
    for (auto& box : it.second) {
      float mx = std::max({box.size[0], box.size[1], box.size[2]}), mn = std::min({box.size[0], box.size[1], box.size[2]});
      float med = box.size[0];
      for (int i = 1; i < 3; i++) if (med == mx || med == mn) med = box.size[i];
      med_sum += med;
      med_sq += med * med;
    }
    for (float r : class_radii[it.first]) { rad_sum += r; rad_sq += r * r; }
    const float n = (float)it.second.size();
    med_sum /= n; med_sq /= n; rad_sum /= n; rad_sq /= n;
    m_voting.m_dimensions_map.insert({it.first, {rad_sum, med_sum}});
    m_voting.m_variance_map.insert({it.first, {rad_sq - rad_sum * rad_sum, med_sq - med_sum * med_sum}});
  }
  const int64_t F = foff.back();
  const int n_clouds = (int)cloud_class.size();
  if (F == 0) throw RuntimeException("no features were extracted from the training models");
  const int k = m_params.knn_k;
  unsigned max_class = 0;
  for (unsigned c : cloud_class) max_class = std::max(max_class, c);
  const int n_classes = (int)max_class + 1;
  // codewords = features (clustering None): activate every feature against all of them on the GPU
  log("INFO", "activating codewords");
  {
    std::vector<int64_t> voff((size_t)F + 1);
    std::iota(voff.begin(), voff.end(), 0);
    std::vector<float> zeros3((size_t)F * 3, 0.f), ones((size_t)F, 1.f), bbox((size_t)F * 7, 0.f), sig((size_t)n_classes, 1.f);
    std::vector<uint32_t> zu((size_t)F, 0u);
    check(pcdb_set_codebook(m_ctx, fdesc.data(), F, D, voff.data(), zeros3.data(), ones.data(), zu.data(), zu.data(),
                            bbox.data(), nullptr, fxyz.data(), nullptr, nullptr, sig.data(), n_classes, 0));
  }
  pcdb_params tp = m_params;
  tp.use_distance_ratio = 0;  // detection only (activation_strategy_knn.h:67)
  check(pcdb_set_params(m_ctx, &tp));
  std::vector<int32_t> idx((size_t)F * k), cnt((size_t)F);
  std::vector<float> dist((size_t)F * k);
  check(pcdb_knn(m_ctx, fdesc.data(), F, k, m_params.distance_type, PCDB_KNN_AUTO, idx.data(), dist.data(), cnt.data()));
  check(pcdb_set_params(m_ctx, &m_params));
  // Codebook::activate codebook.cpp:64-368
  struct Entry {
    std::vector<float> votes, bbox;
    std::vector<uint32_t> cls, inst;
    std::vector<int64_t> feat;
    std::vector<int> cloud;
  };
  std::map<int, Entry> distribution;
  std::map<unsigned, float> sigmas;
  int c0 = 0;
  while (c0 < n_clouds) {
    int c1 = c0;
    const unsigned cls = cloud_class[c0];
    while (c1 < n_clouds && cloud_class[c1] == cls) ++c1;
    const int64_t num_features = foff[c1] - foff[c0];
    const int max_elements = (int)std::sqrt((double)num_features);
    std::vector<int64_t> allModelFeatures;
    std::vector<int> allActivated;
    for (int c = c0; c < c1; ++c) {
      const Utils::BoundingBox& bb = boxes[c];
      for (int64_t f = foff[c]; f < foff[c + 1]; ++f) {
        const Quat rq = lrf_quat(&flrf[9 * f]);
        for (int j = 0; j < cnt[f]; ++j) {
          Entry& e = distribution[idx[f * k + j]];  // CodewordDistribution::addCodeword :37-71
          float vote[3] = {bb.position[0] - fxyz[3 * f], bb.position[1] - fxyz[3 * f + 1], bb.position[2] - fxyz[3 * f + 2]};
          float rot[3];
          quat_rotate(rq, vote, rot);
          e.votes.insert(e.votes.end(), rot, rot + 3);
          e.cls.push_back(cls);
          e.inst.push_back(cloud_inst[c]);
          Quat nq = qmul(Quat{bb.rotQuat[0], bb.rotQuat[1], bb.rotQuat[2], bb.rotQuat[3]}, qconj(rq));
          const float nb[7] = {nq.a, nq.b, nq.c, nq.d, bb.size[0], bb.size[1], bb.size[2]};
          e.bbox.insert(e.bbox.end(), nb, nb + 7);
          e.feat.push_back(f);
          e.cloud.push_back(c);
        }
        if ((int)allActivated.size() < max_elements)
          for (int j = 0; j < cnt[f]; ++j) allActivated.push_back(idx[f * k + j]);
      }
      if ((int)allModelFeatures.size() < max_elements)
        for (int64_t f = foff[c]; f < foff[c + 1]; ++f) allModelFeatures.push_back(f);
    }
    // class variance :166-193 — the functor values come from the GPU (pcdb_distance_pairs)
    const size_t np = allModelFeatures.size() * allActivated.size();
    if (np > 0) {
      std::vector<float> a(np * D), b(np * D), d(np);
      size_t t = 0;
      for (int64_t f : allModelFeatures)
        for (int wd : allActivated) {
          std::copy(fdesc.begin() + f * D, fdesc.begin() + (f + 1) * D, a.begin() + t * D);
          std::copy(fdesc.begin() + (int64_t)wd * D, fdesc.begin() + ((int64_t)wd + 1) * D, b.begin() + t * D);
          ++t;
        }
      check(pcdb_distance_pairs(m_ctx, a.data(), b.data(), (int64_t)np, D, m_params.distance_type, d.data()));
      float sum = 0;
      for (float v : d) sum += v;
      const int num = (int)np;
      const float mean = sum / num;
      float variance = 0;
      for (float v : d) { float diff = v - mean; variance += diff * diff; }
      variance /= num - 1;
      sigmas[cls] = variance;
    }
    c0 = c1;
  }
  if (k == 1)  // clean-up :201-224
    for (auto it = distribution.begin(); it != distribution.end();)
      it = it->second.cls.size() != 1 ? distribution.erase(it) : std::next(it);
  Codebook cb;
  cb.dim = D;
  cb.classSigmas = sigmas;
  cb.vote_off.push_back(0);
  cb.cw_off.push_back(0);
  for (auto& kv : distribution) {
    const Entry& e = kv.second;
    const int nv = (int)e.cls.size();
    const int64_t id = kv.first;
    cb.ids.push_back((int32_t)id);
    cb.numFeatures.push_back(1);
    cb.weights.push_back(1.0f);
    cb.classIds.push_back((int32_t)cloud_class[std::upper_bound(foff.begin(), foff.end(), id) - foff.begin() - 1]);
    cb.words.insert(cb.words.end(), fdesc.begin() + id * D, fdesc.begin() + (id + 1) * D);
    cb.keypoints.insert(cb.keypoints.end(), fxyz.begin() + 3 * id, fxyz.begin() + 3 * id + 3);
    for (int i = 0; i < nv; ++i) {  // computeWeights :171-243
      std::vector<float> lw;
      const float* centre = boxes[e.cloud[i]].position;
      for (int j = 0; j < nv; ++j) {
        const int64_t f = e.feat[j];
        float rot[3];
        quat_rotate_inv(lrf_quat(&flrf[9 * f]), &e.votes[3 * i], rot);
        float d0 = fxyz[3 * f] + rot[0] - centre[0], d1 = fxyz[3 * f + 1] + rot[1] - centre[1], d2 = fxyz[3 * f + 2] + rot[2] - centre[2];
        float dd = std::sqrt(d0 * d0 + d1 * d1 + d2 * d2);
        const float sigma = 0.5f;
        lw.push_back(std::exp((-1 * (dd * dd)) / (sigma * sigma)));
      }
      std::sort(lw.begin(), lw.end());
      cb.vote_weight.push_back(lw.size() % 2 == 0 ? (lw[lw.size() / 2 - 1] + lw[lw.size() / 2]) / 2 : lw[lw.size() / 2]);
    }
    cb.vote_xyz.insert(cb.vote_xyz.end(), e.votes.begin(), e.votes.end());
    cb.vote_bbox.insert(cb.vote_bbox.end(), e.bbox.begin(), e.bbox.end());
    cb.vote_class.insert(cb.vote_class.end(), e.cls.begin(), e.cls.end());
    cb.vote_instance.insert(cb.vote_instance.end(), e.inst.begin(), e.inst.end());
    cb.vote_class_weight.insert(cb.vote_class_weight.end(), (size_t)nv, 1.0f);
    // statistical class weights (codebook.cpp:226-366) are only read with UseClassWeight=true, which this path does
    // not train: store weight 1 per distinct class so the archive stays loadable by the reference
    std::vector<uint32_t> distinct(e.cls);
    std::sort(distinct.begin(), distinct.end());
    distinct.erase(std::unique(distinct.begin(), distinct.end()), distinct.end());
    for (uint32_t c : distinct) { cb.cw_class.push_back((int32_t)c); cb.cw_weight.push_back(1.0f); }
    cb.vote_off.push_back((int64_t)cb.vote_class.size());
    cb.cw_off.push_back((int64_t)cb.cw_class.size());
  }
  if (m_params.use_class_weight)
    throw BadParamException("UseClassWeight=true: training the statistical weights (codebook.cpp:226-366) is a 'next' row");
  m_codebook = std::move(cb);
  log("INFO", "Size of distribution at the end of training: " + std::to_string(m_codebook.getSize()));
  uploadCodebook();
}

// ---- detection -----------------------------------------------------------------------------------------------------------
std::vector<VotingMaximum> ImplicitShapeModel::toMaxima(const pcdb_maximum* mx, int64_t n, const std::vector<pcdb_vote>* votes,
                                                        const std::vector<int64_t>* member_idx,
                                                        const std::vector<float>* member_w) const {
  std::vector<VotingMaximum> out((size_t)n);
  for (int64_t i = 0; i < n; ++i) {
    VotingMaximum& m = out[i];
    const pcdb_maximum& s = mx[i];
    std::copy(s.position, s.position + 3, m.position);
    m.weight = s.weight;
    m.classId = s.class_id;
    m.instanceId = s.instance_id;
    m.instanceWeight = s.instance_weight;
    std::copy(s.position, s.position + 3, m.boundingBox.position);
    std::copy(s.bbox_quat, s.bbox_quat + 4, m.boundingBox.rotQuat);
    std::copy(s.bbox_size, s.bbox_size + 3, m.boundingBox.size);
    m.globalHypothesis.classId = s.class_id;  // voting.cpp:176-178
    m.globalHypothesis.instanceId = s.instance_id;
    m.globalHypothesis.classWeight = s.raw_weight;
    m.globalHypothesis.instanceWeight = s.instance_weight;
    if (votes && member_idx) {
      m.votes.resize((size_t)s.n_votes);
      for (int j = 0; j < s.n_votes; ++j) {
        const pcdb_vote& v = (*votes)[(size_t)(*member_idx)[(size_t)s.vote_begin + j]];
        Vote& o = m.votes[j];
        std::copy(v.position, v.position + 3, o.position);
        o.weight = (*member_w)[(size_t)s.vote_begin + j];
        o.classId = v.class_id;
        o.instanceId = v.instance_id;
        std::copy(v.keypoint, v.keypoint + 3, o.keypoint);
        std::copy(v.keypoint_training, v.keypoint_training + 3, o.keypoint_training);
        std::copy(v.bbox_quat, v.bbox_quat + 4, o.boundingBox.rotQuat);
        std::copy(v.bbox_size, v.bbox_size + 3, o.boundingBox.size);
        o.codewordId = v.codeword_id;
      }
    }
  }
  return out;
}

bool ImplicitShapeModel::detectBatch(const std::vector<std::string>& filenames,
                                     std::vector<std::vector<VotingMaximum>>& maxima,
                                     std::map<std::string, double>& times, int batch, bool with_votes) {
  if (!m_codebook_uploaded) uploadCodebook();
  maxima.assign(filenames.size(), {});
  for (size_t b0 = 0; b0 < filenames.size(); b0 += (size_t)batch) {
    const size_t b1 = std::min(filenames.size(), b0 + (size_t)batch);
    std::vector<float> xyz, nrm;
    std::vector<uint32_t> rgb;
    std::vector<int64_t> off{0};
    for (size_t i = b0; i < b1; ++i) {
      io::Cloud c;
      std::string err;
      if (!io::load_cloud(filenames[i], c, err)) { log("ERROR", err); return false; }
      if (c.size() == 0) { log("ERROR", "point cloud is empty"); return false; }
      // detect(): "first normal is zero/NaN => hasNormals=false" (implicit_shape_model.cpp:614-625)
      if (!c.has_normals || (c.normals[0] == 0 && c.normals[1] == 0 && c.normals[2] == 0) || std::isnan(c.normals[0])) {
        // a batch shares one normals array: estimate this cloud's normals now (the "normals" bucket of the timing table)
        const int64_t o1[2] = {0, (int64_t)c.size()};
        c.normals.assign(c.size() * 3, 0.f);
        const auto t0 = std::chrono::steady_clock::now();
        if (c.organized())  // points->isOrganized(): integral-image normals (implicit_shape_model.cpp:948-966)
          check(pcdb_compute_normals_organized(m_ctx, c.xyz.data(), (int32_t)c.width, (int32_t)c.height, c.normals.data()));
        else
          check(pcdb_compute_normals(m_ctx, c.xyz.data(), o1, 1, c.normals.data(), nullptr));
        m_processing_times["normals"] += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
      }
      xyz.insert(xyz.end(), c.xyz.begin(), c.xyz.end());
      nrm.insert(nrm.end(), c.normals.begin(), c.normals.end());
      rgb.insert(rgb.end(), c.rgb.begin(), c.rgb.end());
      off.push_back(off.back() + (int64_t)c.size());
    }
    const int B = (int)(b1 - b0);
    std::vector<int32_t> labels((size_t)B);
    // labels and per-cloud maxima counts first, then buffers of exactly the size the device produced (the maxima are
    // bounded by the seeds, the votes by F * k * votes per codeword: neither by the point count)
    std::vector<int64_t> moff((size_t)B + 1);
    double t[7];
    check(pcdb_classify_batch(m_ctx, xyz.data(), nrm.data(), rgb.data(), off.data(), B, labels.data(), nullptr,
                              moff.data(), 0, t));
    int64_t n_votes = 0, n_max = 0, n_mem = 0;
    check(pcdb_get_last_sizes(m_ctx, &n_votes, &n_max, &n_mem));
    std::vector<pcdb_maximum> mx((size_t)std::max<int64_t>(1, n_max));
    check(pcdb_get_maxima(m_ctx, mx.data(), moff.data(), (int64_t)mx.size()));
    static const char* keys[7] = {"complete", "features", "keypoints", "normals", "flann", "voting", "maxima"};
    for (int i = 0; i < 7; ++i) m_processing_times[keys[i]] += t[i];
    std::vector<pcdb_vote> votes;
    std::vector<int64_t> midx, voff((size_t)B + 1, 0);
    std::vector<float> mw;
    if (with_votes) {
      midx.resize((size_t)n_mem + 1);
      mw.resize((size_t)n_mem + 1);
      int64_t n = 0;
      check(pcdb_get_maximum_votes(m_ctx, midx.data(), mw.data(), n_mem + 1, &n));
      votes.resize((size_t)n_votes + 1);
      check(pcdb_get_votes(m_ctx, votes.data(), voff.data(), (int64_t)votes.size()));
    }
    for (int b = 0; b < B; ++b)
      maxima[b0 + b] = toMaxima(mx.data() + moff[b], moff[b + 1] - moff[b], with_votes ? &votes : nullptr,
                                with_votes ? &midx : nullptr, with_votes ? &mw : nullptr);
    if (with_votes && B == 1) {  // Voting::getVotes(): votes of the last detect() by class (training_gui.cpp:1012)
      m_voting.m_votes.clear();
      for (int64_t v = voff[0]; v < voff[1]; ++v) {
        const pcdb_vote& s = votes[(size_t)v];
        Vote o;
        std::copy(s.position, s.position + 3, o.position);
        o.weight = s.weight;
        o.classId = s.class_id;
        o.instanceId = s.instance_id;
        std::copy(s.keypoint, s.keypoint + 3, o.keypoint);
        std::copy(s.keypoint_training, s.keypoint_training + 3, o.keypoint_training);
        std::copy(s.bbox_quat, s.bbox_quat + 4, o.boundingBox.rotQuat);
        std::copy(s.bbox_size, s.bbox_size + 3, o.boundingBox.size);
        o.codewordId = s.codeword_id;
        m_voting.m_votes[s.class_id].push_back(o);
      }
    }
  }
  times = m_processing_times;
  return true;
}

bool ImplicitShapeModel::detect(const std::string& filename, std::vector<VotingMaximum>& maxima,
                                std::map<std::string, double>& times) {
  std::vector<std::vector<VotingMaximum>> all;
  if (!detectBatch({filename}, all, times, 1, true)) return false;
  maxima = std::move(all[0]);
  log("INFO", "detected " + std::to_string(maxima.size()) + " maxima");
  return true;
}

std::tuple<std::vector<VotingMaximum>, std::map<std::string, double>> ImplicitShapeModel::detect(const io::Cloud& points,
                                                                                                 bool hasNormals) {
  if (points.size() == 0) {
    log("WARN", "point cloud is empty");
    return std::make_tuple(std::vector<VotingMaximum>(), m_processing_times);
  }
  // implicit_shape_model.cpp:614-625: a zero / NaN first normal means "no normals"
  if (hasNormals && (!points.has_normals || (points.normals[0] == 0 && points.normals[1] == 0 && points.normals[2] == 0) ||
                     std::isnan(points.normals[0])))
    hasNormals = false;
  if (!m_codebook_uploaded) uploadCodebook();
  const int64_t off[2] = {0, (int64_t)points.size()};
  int32_t label;
  int64_t moff[2];
  double t[7];
  std::vector<float> org_normals;
  const float* normals_in = hasNormals ? points.normals.data() : nullptr;
  if (!hasNormals && points.organized()) {  // isOrganized(): integral-image normals (implicit_shape_model.cpp:948-966)
    org_normals.assign(points.size() * 3, 0.f);
    check(pcdb_compute_normals_organized(m_ctx, points.xyz.data(), (int32_t)points.width, (int32_t)points.height,
                                         org_normals.data()));
    normals_in = org_normals.data();
  }
  check(pcdb_classify_batch(m_ctx, points.xyz.data(), normals_in, points.has_rgb ? points.rgb.data() : nullptr, off, 1,
                            &label, nullptr, moff, 0, t));
  int64_t n_max = 0;
  check(pcdb_get_last_sizes(m_ctx, nullptr, &n_max, nullptr));
  std::vector<pcdb_maximum> mx((size_t)std::max<int64_t>(1, n_max));
  check(pcdb_get_maxima(m_ctx, mx.data(), moff, (int64_t)mx.size()));
  static const char* keys[7] = {"complete", "features", "keypoints", "normals", "flann", "voting", "maxima"};
  for (int i = 0; i < 7; ++i) m_processing_times[keys[i]] += t[i];
  return std::make_tuple(toMaxima(mx.data(), moff[1], nullptr, nullptr, nullptr), m_processing_times);
}

}  // namespace ism3d
