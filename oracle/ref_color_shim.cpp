// C entry points over the reference's own ism3d::ColorConversion
// (/root/reference/src/implicit_shape_model/third_party/pcl_color_conversion/color_conversion.{h,cpp}),
// compiled from where it lies by oracle/Makefile into oracle/_ref/libref_color.so.
// Test infrastructure: used only to pin oracle/pcd_oracle.cpp's Lab restatement.
#include "color_conversion.h"
#include <cstdint>
extern "C" {
int ref_rgb_to_lab_normalized(const uint32_t* rgb, int64_t n, float* lab_out) {
  const ism3d::ColorConversion& cc = ism3d::ColorConversionStatic::getColorConversion();
  for (int64_t i = 0; i < n; ++i)
    cc.RgbToCieLabNormalized((rgb[i] >> 16) & 0xFF, (rgb[i] >> 8) & 0xFF, rgb[i] & 0xFF, lab_out[3 * i],
                             lab_out[3 * i + 1], lab_out[3 * i + 2]);
  return 0;
}
int ref_color_distance(const float* lab, const float* ref, int64_t n, float* out) {
  const ism3d::ColorConversion& cc = ism3d::ColorConversionStatic::getColorConversion();
  for (int64_t i = 0; i < n; ++i)
    out[i] = cc.getColorDistance(lab[3 * i], lab[3 * i + 1], lab[3 * i + 2], ref[3 * i], ref[3 * i + 1], ref[3 * i + 2]);
  return 0;
}
int ref_lab_luts(float* srgb256, float* sxyz4000) {
  const ism3d::ColorConversion& cc = ism3d::ColorConversionStatic::getColorConversion();
  for (int i = 0; i < 256; ++i) srgb256[i] = cc.sRGB_LUT[i];
  for (int i = 0; i < 4000; ++i) sxyz4000[i] = cc.sXYZ_LUT[i];
  return 0;
}
}
