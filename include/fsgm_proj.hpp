// fsgm_proj.hpp — C++ facade over the C ABI (include/fsgm.h) with the shape of the reference's proj/ library:
//   EpiSGM::compute / PydSGM::compute     proj/include/epi_sgm.h:6-12, proj/include/pyd_sgm.h:7-13 (stubs returning zeros in the
//                                         reference, proj/src/epi_sgm.cpp:3-6, proj/src/pyd_sgm.cpp:3-6)
//   FlowField read/write                  KITTI 16-bit flow PNG, proj/src/utils.cpp:3-73  ((u*64 + 32768), valid flag)
//   read_calib_file                       proj/src/utils.cpp:129-169
// The reference builds on OpenCV (cv::Mat, imread/imwrite); OpenCV's C++ headers are not available, so the facade carries its
// own minimal image type and PNG codec (zlib).  Host code only: every pixel operation goes through libfsgm.so (CUDA, sm_100a).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

struct fsgm_ctx;

namespace fsgm_proj {

// 8-bit image, row-major, `channels` interleaved samples per pixel (1 = gray, 3 = RGB)
struct Image {
    int rows = 0, cols = 0, channels = 1;
    std::vector<uint8_t> data;
    bool empty() const { return data.empty(); }
};
// H x W x 2 float flow (the reference's CV_32FC2 return type) plus a validity mask (KITTI's third channel)
struct FlowField {
    int rows = 0, cols = 0;
    std::vector<float> uv;          // [rows][cols][2] = (u, v)
    std::vector<uint8_t> valid;     // [rows][cols]
    std::vector<uint32_t> cost;     // [rows][cols] minimum aggregated cost (minC), empty when read from a file
};

// ---- PNG (non-interlaced, 8/16 bit, gray / RGB / with alpha) ---------------------------------------------------------------
// decoded samples: 8-bit files -> u8, 16-bit files -> u16 (host order); channels as in the file
struct PngData { int rows = 0, cols = 0, channels = 0, bit_depth = 0; std::vector<uint16_t> samples; };
bool png_read(const std::string& path, PngData* out, std::string* err);
bool png_write(const std::string& path, int rows, int cols, int channels, int bit_depth, const uint16_t* samples, std::string* err);

Image imread(const std::string& path, std::string* err);                 // 8-bit gray or RGB (alpha dropped, 16-bit >> 8)
// rgb2gray as the MATLAB drivers apply it before the gateways (epipolar_sgm_of.m:36-39, pyramidal_sgm.m:44-45):
// 0.298936021293775 R + 0.587043074451121 G + 0.114020904255103 B, rounded to nearest
Image to_gray(const Image& img);

// KITTI flow PNG (proj/src/utils.cpp:3-73): 16-bit RGB, R = u*64 + 32768, G = v*64 + 32768, B = valid; values clamped to 0..65535
bool flow_write_kitti(const std::string& path, const FlowField& f, std::string* err);
bool flow_read_kitti(const std::string& path, FlowField* f, std::string* err);
// KITTI outlier statistic on the valid ground-truth pixels: endpoint error > 3 px AND > 5 % of the magnitude
double flow_outlier_rate(const FlowField& estimate, const FlowField& ground_truth, double* mean_epe);

// 3x4 projection matrix of camera 0 from a KITTI calibration file (proj/src/utils.cpp:129-169): 2012 files start with "P0: ...";
// 2015 files (isKITTI2015) carry it on the tenth line.  Row-major float[12].
bool read_calib_file(const std::string& path, bool isKITTI2015, float P[12], std::string* err);

// ---- the two classes of proj/include ------------------------------------------------------------------------------------
class PydSGM {
public:
    int numPyd = 5;                       // -N (proj/src/sgmof_main.cpp:22); the rest are the constants of pyramidal_sgm.m:14-22
    int passNum = 2;                      // -p
    bool enableDiagonal = true;           // -d
    int P1 = 6, P2 = 32, aggHalfWinSize = 2, verSearchHalfWinSize = 5, horSearchHalfWinSize = 5;
    explicit PydSGM(int device = 0);
    ~PydSGM();
    PydSGM(const PydSGM&) = delete;
    PydSGM& operator=(const PydSGM&) = delete;
    // flow from I1 to I2 (pyramidal_sgm.m); throws std::runtime_error on failure (size mismatch, CUDA error, no sm_100 device)
    FlowField compute(const Image& I1, const Image& I2);
private:
    fsgm_ctx* ctx_ = nullptr;
};

class EpiSGM {
public:
    int dMax = 64;                        // epipolar_sgm_of.m:13-15
    double vMax = 0.3;                    // :16-18
    int P1 = 6, P2 = 64;                  // :20
    bool enableDiagonal = false;          // -d; the reference ships 4 paths (calc_cost_sgm.cpp:104)
    bool vzIndex = true;                  // -V (USE_VZIND, calc_cost_sgm.cpp:4)
    explicit EpiSGM(int device = 0);
    ~EpiSGM();
    EpiSGM(const EpiSGM&) = delete;
    EpiSGM& operator=(const EpiSGM&) = delete;
    // Two-view geometry, the output of the reference's epipolar_geometry.m:31-98 (SURF matching + LMedS + SVD stay with the
    // caller): fundamental matrix F and rotation homography H (row-major 3x3), epipole in image 2 (1-based pixel coordinates),
    // expansion flag.  compute() throws if it has not been set.
    void setGeometry(const double F[9], const double H[9], const double epipole[2], bool direction);
    FlowField compute(const Image& I1, const Image& I2);
private:
    fsgm_ctx* ctx_ = nullptr;
    bool have_geo_ = false;
    double F_[9], H_[9], epi_[2];
    int direction_ = 0;
};

}  // namespace fsgm_proj

// ---- flat C hooks over the I/O helpers (ctypes test harness; libfsgm_proj.so) ---------------------------------------------
extern "C" {
int fsgm_proj_png_info(const char* path, int* rows, int* cols, int* channels, int* bit_depth);
int fsgm_proj_png_read(const char* path, uint16_t* samples /* rows*cols*channels */);
int fsgm_proj_png_write(const char* path, int rows, int cols, int channels, int bit_depth, const uint16_t* samples);
int fsgm_proj_imread_gray(const char* path, uint8_t* gray /* rows*cols */);
int fsgm_proj_flow_write(const char* path, int rows, int cols, const float* uv, const uint8_t* valid);
int fsgm_proj_flow_read(const char* path, float* uv, uint8_t* valid);
int fsgm_proj_read_calib(const char* path, int isKITTI2015, float* P12);
}
