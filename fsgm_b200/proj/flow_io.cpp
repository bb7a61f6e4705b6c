// KITTI flow PNG and calibration files for the proj/ facade (proj/src/utils.cpp:3-73 and :129-169).
#include "../../include/fsgm_proj.hpp"
#include <algorithm>
#include <cmath>
#include <fstream>
#include <sstream>

namespace fsgm_proj {

// proj/src/utils.cpp:46-73: channel order there is OpenCV's BGR, so file R = u, G = v, B = valid; float arithmetic, truncation
bool flow_write_kitti(const std::string& path, const FlowField& f, std::string* err)
{
    const size_t n = (size_t)f.rows * f.cols;
    if (f.uv.size() != 2 * n || f.valid.size() != n) { if (err) *err = "flow_write_kitti: inconsistent flow field"; return false; }
    std::vector<uint16_t> px(3 * n, 0);
    for (size_t i = 0; i < n; ++i) {
        if (!f.valid[i]) continue;
        px[3 * i + 0] = (uint16_t)std::max(std::min(f.uv[2 * i] * 64.0f + 32768.0f, 65535.0f), 0.0f);
        px[3 * i + 1] = (uint16_t)std::max(std::min(f.uv[2 * i + 1] * 64.0f + 32768.0f, 65535.0f), 0.0f);
        px[3 * i + 2] = 1;
    }
    return png_write(path, f.rows, f.cols, 3, 16, px.data(), err);
}

// proj/src/utils.cpp:3-44
bool flow_read_kitti(const std::string& path, FlowField* f, std::string* err)
{
    PngData p;
    if (!png_read(path, &p, err)) return false;
    if (p.channels < 3 || p.bit_depth != 16) { if (err) *err = path + ": not a valid KITTI format flow file"; return false; }
    const size_t n = (size_t)p.rows * p.cols;
    f->rows = p.rows; f->cols = p.cols;
    f->uv.assign(2 * n, 0.0f); f->valid.assign(n, 0); f->cost.clear();
    for (size_t i = 0; i < n; ++i) {
        const uint16_t* s = &p.samples[i * p.channels];
        if (!s[2]) continue;
        f->uv[2 * i] = (s[0] - 32768.0f) / 64.0f;
        f->uv[2 * i + 1] = (s[1] - 32768.0f) / 64.0f;
        f->valid[i] = 1;
    }
    return true;
}

double flow_outlier_rate(const FlowField& est, const FlowField& gt, double* mean_epe)
{
    size_t cnt = 0, bad = 0;
    double sum = 0;
    if (est.rows != gt.rows || est.cols != gt.cols) return -1.0;
    for (size_t i = 0; i < (size_t)gt.rows * gt.cols; ++i) {
        if (!gt.valid[i]) continue;
        const double du = est.uv[2 * i] - gt.uv[2 * i], dv = est.uv[2 * i + 1] - gt.uv[2 * i + 1];
        const double e = std::sqrt(du * du + dv * dv), mag = std::sqrt((double)gt.uv[2 * i] * gt.uv[2 * i] + (double)gt.uv[2 * i + 1] * gt.uv[2 * i + 1]);
        ++cnt; sum += e;
        if (e > 3.0 && e > 0.05 * mag) ++bad;
    }
    if (mean_epe) *mean_epe = cnt ? sum / cnt : 0.0;
    return cnt ? (double)bad / cnt : 0.0;
}

// proj/src/utils.cpp:129-169: KITTI 2012 files start with the "P0:" row; 2015 files have it after nine other lines
bool read_calib_file(const std::string& path, bool isKITTI2015, float P[12], std::string* err)
{
    std::ifstream f(path);
    if (!f.is_open()) { if (err) *err = "can't open calibration file " + path; return false; }
    std::string line, tag;
    if (isKITTI2015)
        for (int i = 0; i < 9; ++i) std::getline(f, line);
    f >> tag;
    for (int i = 0; i < 12; ++i)
        if (!(f >> P[i])) { if (err) *err = path + ": expected a tag followed by 12 numbers"; return false; }
    return true;
}

}  // namespace fsgm_proj

extern "C" {
using namespace fsgm_proj;
int fsgm_proj_png_info(const char* path, int* rows, int* cols, int* channels, int* bit_depth)
{
    PngData p; std::string e;
    if (!png_read(path, &p, &e)) return -1;
    *rows = p.rows; *cols = p.cols; *channels = p.channels; *bit_depth = p.bit_depth;
    return 0;
}
int fsgm_proj_png_read(const char* path, uint16_t* samples)
{
    PngData p; std::string e;
    if (!png_read(path, &p, &e)) return -1;
    std::copy(p.samples.begin(), p.samples.end(), samples);
    return 0;
}
int fsgm_proj_png_write(const char* path, int rows, int cols, int channels, int bit_depth, const uint16_t* samples)
{
    std::string e;
    return png_write(path, rows, cols, channels, bit_depth, samples, &e) ? 0 : -1;
}
int fsgm_proj_imread_gray(const char* path, uint8_t* gray)
{
    std::string e;
    const Image g = to_gray(imread(path, &e));
    if (g.empty()) return -1;
    std::copy(g.data.begin(), g.data.end(), gray);
    return 0;
}
int fsgm_proj_flow_write(const char* path, int rows, int cols, const float* uv, const uint8_t* valid)
{
    FlowField f; std::string e;
    f.rows = rows; f.cols = cols;
    f.uv.assign(uv, uv + (size_t)2 * rows * cols); f.valid.assign(valid, valid + (size_t)rows * cols);
    return flow_write_kitti(path, f, &e) ? 0 : -1;
}
int fsgm_proj_flow_read(const char* path, float* uv, uint8_t* valid)
{
    FlowField f; std::string e;
    if (!flow_read_kitti(path, &f, &e)) return -1;
    std::copy(f.uv.begin(), f.uv.end(), uv); std::copy(f.valid.begin(), f.valid.end(), valid);
    return 0;
}
int fsgm_proj_read_calib(const char* path, int isKITTI2015, float* P12)
{
    std::string e;
    return read_calib_file(path, isKITTI2015 != 0, P12, &e) ? 0 : -1;
}
}
