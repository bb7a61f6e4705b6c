// Minimal PNG codec for the proj/ facade (replaces the cv::imread / cv::imwrite calls of proj/src/sgmof_main.cpp:50-51 and
// proj/src/utils.cpp:5,72).  Non-interlaced files, bit depth 8 or 16, colour types 0/2/4/6.  zlib does the (de)compression.
#include "../../include/fsgm_proj.hpp"
#include <zlib.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace fsgm_proj {
namespace {

uint32_t be32(const uint8_t* p) { return (uint32_t)p[0] << 24 | (uint32_t)p[1] << 16 | (uint32_t)p[2] << 8 | p[3]; }
void put32(std::vector<uint8_t>& v, uint32_t x) { for (int s = 24; s >= 0; s -= 8) v.push_back((uint8_t)(x >> s)); }
bool fail(std::string* err, const std::string& what) { if (err) *err = what; return false; }
int paeth(int a, int b, int c)
{
    const int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
    return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}
void chunk(std::vector<uint8_t>& out, const char type[4], const std::vector<uint8_t>& body)
{
    put32(out, (uint32_t)body.size());
    const size_t start = out.size();
    out.insert(out.end(), type, type + 4);
    out.insert(out.end(), body.begin(), body.end());
    put32(out, (uint32_t)crc32(0L, out.data() + start, (uInt)(out.size() - start)));
}

}  // namespace

bool png_read(const std::string& path, PngData* out, std::string* err)
{
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) return fail(err, "cannot open " + path);
    std::vector<uint8_t> file;
    uint8_t buf[65536];
    size_t n;
    while ((n = std::fread(buf, 1, sizeof buf, f)) > 0) file.insert(file.end(), buf, buf + n);
    std::fclose(f);
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    if (file.size() < 8 || std::memcmp(file.data(), sig, 8) != 0) return fail(err, path + ": not a PNG file");
    size_t pos = 8;
    int W = 0, H = 0, depth = 0, ctype = -1, interlace = 0;
    std::vector<uint8_t> idat;
    while (pos + 12 <= file.size()) {
        const uint32_t len = be32(&file[pos]);
        const char* type = reinterpret_cast<const char*>(&file[pos + 4]);
        if (pos + 12 + (size_t)len > file.size()) return fail(err, path + ": truncated chunk");
        const uint8_t* body = &file[pos + 8];
        if (!std::memcmp(type, "IHDR", 4) && len >= 13) {
            W = (int)be32(body); H = (int)be32(body + 4); depth = body[8]; ctype = body[9]; interlace = body[12];
        } else if (!std::memcmp(type, "IDAT", 4)) idat.insert(idat.end(), body, body + len);
        else if (!std::memcmp(type, "IEND", 4)) break;
        pos += 12 + (size_t)len;
    }
    const int ch = ctype == 0 ? 1 : ctype == 2 ? 3 : ctype == 4 ? 2 : ctype == 6 ? 4 : 0;
    if (W < 1 || H < 1 || !ch || (depth != 8 && depth != 16) || interlace) return fail(err, path + ": unsupported PNG variant");
    const size_t bpp = (size_t)ch * depth / 8, stride = (size_t)W * bpp;
    std::vector<uint8_t> raw((stride + 1) * H);
    uLongf raw_len = (uLongf)raw.size();
    if (uncompress(raw.data(), &raw_len, idat.data(), (uLong)idat.size()) != Z_OK || raw_len != raw.size())
        return fail(err, path + ": inflate failed");
    std::vector<uint8_t> prev(stride, 0), cur(stride);
    out->rows = H; out->cols = W; out->channels = ch; out->bit_depth = depth;
    out->samples.resize((size_t)W * H * ch);
    for (int y = 0; y < H; ++y) {
        const uint8_t ft = raw[(stride + 1) * y];
        const uint8_t* src = &raw[(stride + 1) * y + 1];
        for (size_t i = 0; i < stride; ++i) {
            const int a = i >= bpp ? cur[i - bpp] : 0, b = prev[i], c = i >= bpp ? prev[i - bpp] : 0;
            int v = src[i];
            switch (ft) {
                case 0: break;
                case 1: v += a; break;
                case 2: v += b; break;
                case 3: v += (a + b) >> 1; break;
                case 4: v += paeth(a, b, c); break;
                default: return fail(err, path + ": bad filter type");
            }
            cur[i] = (uint8_t)v;
        }
        uint16_t* dst = &out->samples[(size_t)y * W * ch];
        if (depth == 8) for (size_t i = 0; i < stride; ++i) dst[i] = cur[i];
        else for (size_t i = 0; i < stride / 2; ++i) dst[i] = (uint16_t)(cur[2 * i] << 8 | cur[2 * i + 1]);
        prev.swap(cur);
    }
    return true;
}

bool png_write(const std::string& path, int rows, int cols, int channels, int bit_depth, const uint16_t* samples, std::string* err)
{
    const int ctype = channels == 1 ? 0 : channels == 3 ? 2 : channels == 2 ? 4 : channels == 4 ? 6 : -1;
    if (rows < 1 || cols < 1 || ctype < 0 || (bit_depth != 8 && bit_depth != 16)) return fail(err, "png_write: unsupported format");
    const size_t stride = (size_t)cols * channels * bit_depth / 8;
    std::vector<uint8_t> raw((stride + 1) * rows);
    for (int y = 0; y < rows; ++y) {
        uint8_t* dst = &raw[(stride + 1) * y];
        *dst++ = 0;                                                     // filter type None
        const uint16_t* src = samples + (size_t)y * cols * channels;
        for (size_t i = 0; i < (size_t)cols * channels; ++i) {
            if (bit_depth == 8) *dst++ = (uint8_t)src[i];
            else { *dst++ = (uint8_t)(src[i] >> 8); *dst++ = (uint8_t)src[i]; }
        }
    }
    uLongf zlen = compressBound((uLong)raw.size());
    std::vector<uint8_t> z(zlen);
    if (compress2(z.data(), &zlen, raw.data(), (uLong)raw.size(), 6) != Z_OK) return fail(err, "png_write: deflate failed");
    z.resize(zlen);
    std::vector<uint8_t> out = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A}, ihdr;
    put32(ihdr, (uint32_t)cols); put32(ihdr, (uint32_t)rows);
    ihdr.push_back((uint8_t)bit_depth); ihdr.push_back((uint8_t)ctype); ihdr.push_back(0); ihdr.push_back(0); ihdr.push_back(0);
    chunk(out, "IHDR", ihdr);
    chunk(out, "IDAT", z);
    chunk(out, "IEND", {});
    FILE* f = std::fopen(path.c_str(), "wb");
    if (!f) return fail(err, "cannot create " + path);
    const bool ok = std::fwrite(out.data(), 1, out.size(), f) == out.size();
    std::fclose(f);
    return ok ? true : fail(err, "short write to " + path);
}

Image imread(const std::string& path, std::string* err)
{
    PngData p;
    Image img;
    if (!png_read(path, &p, err)) return img;
    const int ch = p.channels >= 3 ? 3 : 1, shift = p.bit_depth == 16 ? 8 : 0;
    img.rows = p.rows; img.cols = p.cols; img.channels = ch;
    img.data.resize((size_t)p.rows * p.cols * ch);
    for (size_t i = 0; i < (size_t)p.rows * p.cols; ++i)
        for (int c = 0; c < ch; ++c) img.data[i * ch + c] = (uint8_t)(p.samples[i * p.channels + c] >> shift);
    return img;
}

Image to_gray(const Image& img)
{
    if (img.channels == 1) return img;
    Image g;
    g.rows = img.rows; g.cols = img.cols; g.channels = 1;
    g.data.resize((size_t)img.rows * img.cols);
    for (size_t i = 0; i < g.data.size(); ++i) {
        const uint8_t* p = &img.data[i * img.channels];
        const double v = 0.298936021293775 * p[0] + 0.587043074451121 * p[1] + 0.114020904255103 * p[2];
        g.data[i] = (uint8_t)(v + 0.5);
    }
    return g;
}

}  // namespace fsgm_proj
