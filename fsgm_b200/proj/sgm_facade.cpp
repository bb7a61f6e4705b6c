// EpiSGM / PydSGM: the two classes of the reference's proj/ library (proj/include/epi_sgm.h, proj/include/pyd_sgm.h), whose
// compute() methods are stubs there (proj/src/epi_sgm.cpp:3-6, proj/src/pyd_sgm.cpp:3-6).  Here they run the MATLAB drivers'
// per-pixel work on the GPU through the C ABI: pyramidal_sgm.m -> fsgm_pyramidal_sgm, epipolar_sgm_of.m:23-51 (after the
// geometry fit) -> fsgm_epipolar_sgm_of.  Host C++ only; there is no CPU path.
#include "../../include/fsgm_proj.hpp"
#include "../../include/fsgm.h"
#include <cstring>
#include <stdexcept>

namespace fsgm_proj {
namespace {

fsgm_ctx* make_ctx(int device)
{
    fsgm_ctx* c = nullptr;
    const int rc = fsgm_create(device, &c);
    if (rc != FSGM_OK || !c) throw std::runtime_error("fsgm_create failed (a CUDA device of compute capability 10.x is required): code " + std::to_string(rc));
    return c;
}
void check_pair(const Image& a, const Image& b)
{
    if (a.empty() || b.empty()) throw std::runtime_error("empty image");
    if (a.rows != b.rows || a.cols != b.cols) throw std::runtime_error("Size of image1/2 must match");
}
FlowField to_field(int H, int W, const std::vector<double>& mv, std::vector<uint32_t>&& minC)
{
    FlowField f;
    const size_t n = (size_t)W * H;
    f.rows = H; f.cols = W;
    f.uv.resize(2 * n); f.valid.assign(n, 1); f.cost = std::move(minC);
    for (size_t i = 0; i < n; ++i) { f.uv[2 * i] = (float)mv[i]; f.uv[2 * i + 1] = (float)mv[n + i]; }
    return f;
}
[[noreturn]] void raise(fsgm_ctx* c, const char* what, int rc)
{
    throw std::runtime_error(std::string(what) + " failed (" + std::to_string(rc) + "): " + fsgm_last_error(c));
}

}  // namespace

PydSGM::PydSGM(int device) : ctx_(make_ctx(device)) {}
PydSGM::~PydSGM() { fsgm_destroy(ctx_); }

FlowField PydSGM::compute(const Image& I1, const Image& I2)
{
    check_pair(I1, I2);
    const Image g1 = to_gray(I1), g2 = to_gray(I2);
    const int W = g1.cols, H = g1.rows;
    fsgm_pyd_opts o;
    fsgm_pyd_opts_default(&o);
    o.numPyd = numPyd; o.P1 = P1; o.P2 = P2; o.aggHalfWinSize = aggHalfWinSize; o.verSearchHalfWinSize = verSearchHalfWinSize;
    o.horSearchHalfWinSize = horSearchHalfWinSize; o.enableDiagonal = enableDiagonal ? 1 : 0; o.totalPass = passNum;
    std::vector<double> mv((size_t)2 * W * H);
    std::vector<uint32_t> minC((size_t)W * H);
    const int rc = fsgm_pyramidal_sgm(ctx_, g1.data.data(), g2.data.data(), W, H, &o, mv.data(), minC.data(), nullptr);
    if (rc != FSGM_OK) raise(ctx_, "fsgm_pyramidal_sgm", rc);
    return to_field(H, W, mv, std::move(minC));
}

EpiSGM::EpiSGM(int device) : ctx_(make_ctx(device)) {}
EpiSGM::~EpiSGM() { fsgm_destroy(ctx_); }

void EpiSGM::setGeometry(const double F[9], const double H[9], const double epipole[2], bool direction)
{
    std::memcpy(F_, F, sizeof F_); std::memcpy(H_, H, sizeof H_); std::memcpy(epi_, epipole, sizeof epi_);
    direction_ = direction ? 1 : 0;
    have_geo_ = true;
}

FlowField EpiSGM::compute(const Image& I1, const Image& I2)
{
    check_pair(I1, I2);
    if (!have_geo_) throw std::runtime_error("EpiSGM::compute needs the two-view geometry (setGeometry): feature matching and the "
                                             "fundamental-matrix fit of epipolar_geometry.m are outside this library");
    const Image g1 = to_gray(I1), g2 = to_gray(I2);
    const int W = g1.cols, H = g1.rows;
    fsgm_epi_opts o;
    fsgm_epi_opts_default(&o);
    o.paths = enableDiagonal ? 8 : 4;
    (void)vzIndex;                                // the reference has no non-vz-index cost (USE_VZIND is always defined, calc_cost_sgm.cpp:4)
    // the float form of the fused call returns the field in FlowField's own layout (CV_32FC2: u, v interleaved)
    FlowField f;
    const size_t n = (size_t)W * H;
    f.rows = H; f.cols = W;
    f.uv.resize(2 * n); f.valid.assign(n, 1); f.cost.resize(n);
    int rc = fsgm_epipolar_sgm_of_f32_batch_async(ctx_, 1, g1.data.data(), g2.data.data(), W, H, F_, H_, epi_, &direction_, dMax, vMax,
                                                  P1, P2, &o, f.uv.data(), f.cost.data());
    if (rc == FSGM_OK) rc = fsgm_synchronize(ctx_);
    if (rc != FSGM_OK) raise(ctx_, "fsgm_epipolar_sgm_of_f32_batch_async", rc);
    return f;
}

}  // namespace fsgm_proj
