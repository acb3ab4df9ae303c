// config_params.cpp — prints what ImplicitShapeModel::readObject makes of a .ism configuration (every hot-path
// parameter of pcdb_params), host-only: no device context is created.  Exit code 2 + the exception text when the
// configuration asks for something outside the built path.  Driven by tests/test_host_formats.py.
#include <cstdio>
#include <iostream>

#include "ism3d_b200.h"

int main(int argc, char** argv) {
  if (argc < 2) {
    std::fprintf(stderr, "usage: config_params <file.ism>\n");
    return 1;
  }
  try {
    std::string bb;
    const pcdb_params p = ism3d::ImplicitShapeModel::paramsOfConfigFile(argv[1], &bb);
    std::printf("feature_type=%d feature_radius=%.9g lrf_radius=%.9g leaf_size=%.9g distance_type=%d knn_k=%d "
                "use_distance_ratio=%d distance_ratio_threshold=%.9g use_class_weight=%d use_vote_weight=%d "
                "use_matching_weight=%d use_codeword_weight=%d bandwidth=%.9g ms_threshold=%.9g ms_max_iter=%d ms_kernel=%d "
                "maxima_suppression=%d min_threshold=%.9g min_votes_threshold=%d best_k=%d average_rotation=%d "
                "single_object_mode=%d normal_radius=%.9g consistent_normals_method=%d max_filter_type=%d radius_type=%d "
                "radius_factor=%.9g single_object_max_type=%d bounding_box_type=%s\n",
                p.feature_type, p.feature_radius, p.lrf_radius, (double)p.leaf_size, p.distance_type, p.knn_k,
                p.use_distance_ratio, (double)p.distance_ratio_threshold, p.use_class_weight, p.use_vote_weight,
                p.use_matching_weight, p.use_codeword_weight, (double)p.bandwidth, (double)p.ms_threshold, p.ms_max_iter,
                p.ms_kernel, p.maxima_suppression, (double)p.min_threshold, p.min_votes_threshold, p.best_k,
                p.average_rotation, p.single_object_mode, (double)p.normal_radius, p.consistent_normals_method,
                p.max_filter_type, p.radius_type, (double)p.radius_factor, p.single_object_max_type, bb.c_str());
  } catch (const std::exception& e) {
    std::cerr << e.what() << std::endl;
    return 2;
  }
  return 0;
}
