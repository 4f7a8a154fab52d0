// pt_png.cpp -- minimal PNG reader for the texture loader of the facade / C ABI.
// Replaces `cv::imread(path)` + `cv::cvtColor(src, dst, CV_BGR2RGBA)` of the reference's
// PocaTextureUtils::AddTexByFile (cuSrc/textures.cu:15-17): OpenCV is a Windows binary dependency there and does not
// exist here.  Supports what imread hands back for ordinary files: 8-bit greyscale, grey+alpha, RGB, RGBA and
// palette images, non-interlaced; everything is expanded to RGBA8 rows (alpha 255 when the file has none, exactly
// what CV_BGR2RGBA produces from imread's 3-channel result).  zlib does the inflate; chunk CRCs are verified.
#include <zlib.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/pt_api.h"

namespace {
uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }
int paeth(int a, int b, int c) {
    int p = a + b - c, pa = abs(p - a), pb = abs(p - b), pc = abs(p - c);
    return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}
}  // namespace

extern "C" int pt_set_error_text(const char* msg);   // pt_api.cu

extern "C" int pt_image_load_png(const char* path, uint8_t** rgba_out, int* width, int* height) {
    if (!path || !rgba_out || !width || !height) { pt_set_error_text("pt_image_load_png: null argument"); return PT_ERR_INVALID; }
    FILE* f = fopen(path, "rb");
    if (!f) { pt_set_error_text((std::string("pt_image_load_png: cannot open ") + path).c_str()); return PT_ERR_INVALID; }
    std::vector<uint8_t> file;
    uint8_t buf[65536];
    size_t n;
    while ((n = fread(buf, 1, sizeof(buf), f)) > 0) file.insert(file.end(), buf, buf + n);
    fclose(f);
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    if (file.size() < 33 || memcmp(file.data(), sig, 8) != 0) { pt_set_error_text("pt_image_load_png: not a PNG file"); return PT_ERR_INVALID; }
    uint32_t w = 0, h = 0;
    int depth = 0, ctype = -1, interlace = 0;
    std::vector<uint8_t> idat, plte, trns;
    size_t pos = 8;
    bool end = false;
    while (!end && pos + 12 <= file.size()) {
        const uint32_t len = be32(&file[pos]);
        const uint8_t* type = &file[pos + 4];
        if (pos + 12 + (size_t)len > file.size()) { pt_set_error_text("pt_image_load_png: truncated chunk"); return PT_ERR_INVALID; }
        const uint8_t* data = &file[pos + 8];
        if ((uint32_t)crc32(crc32(0L, Z_NULL, 0), type, len + 4) != be32(data + len)) { pt_set_error_text("pt_image_load_png: chunk CRC mismatch"); return PT_ERR_INVALID; }
        if (!memcmp(type, "IHDR", 4) && len == 13) {
            w = be32(data); h = be32(data + 4); depth = data[8]; ctype = data[9]; interlace = data[12];
        } else if (!memcmp(type, "PLTE", 4)) plte.assign(data, data + len);
        else if (!memcmp(type, "tRNS", 4)) trns.assign(data, data + len);
        else if (!memcmp(type, "IDAT", 4)) idat.insert(idat.end(), data, data + len);
        else if (!memcmp(type, "IEND", 4)) end = true;
        pos += 12 + (size_t)len;
    }
    int channels = ctype == 0 ? 1 : ctype == 2 ? 3 : ctype == 3 ? 1 : ctype == 4 ? 2 : ctype == 6 ? 4 : 0;
    if (!w || !h || w > 65535 || h > 65535 || depth != 8 || !channels || interlace != 0 || (ctype == 3 && plte.empty())) {
        pt_set_error_text("pt_image_load_png: unsupported PNG (need 8-bit, non-interlaced grey / grey+alpha / RGB / RGBA / palette)");
        return PT_ERR_INVALID;
    }
    const size_t stride = (size_t)w * channels;
    std::vector<uint8_t> raw((stride + 1) * h);
    uLongf out_len = (uLongf)raw.size();
    if (uncompress(raw.data(), &out_len, idat.data(), (uLong)idat.size()) != Z_OK || out_len != raw.size()) {
        pt_set_error_text("pt_image_load_png: inflate failed");
        return PT_ERR_INVALID;
    }
    // undo the per-row filters (PNG specification, section 9)
    std::vector<uint8_t> img(stride * h);
    for (uint32_t y = 0; y < h; ++y) {
        const uint8_t ft = raw[(stride + 1) * y];
        const uint8_t* src = &raw[(stride + 1) * y + 1];
        uint8_t* dst = &img[stride * y];
        const uint8_t* up = y ? &img[stride * (y - 1)] : nullptr;
        for (size_t i = 0; i < stride; ++i) {
            const int a = i >= (size_t)channels ? dst[i - channels] : 0;
            const int b = up ? up[i] : 0;
            const int c = (up && i >= (size_t)channels) ? up[i - channels] : 0;
            int v = src[i];
            switch (ft) {
                case 0: break;
                case 1: v += a; break;
                case 2: v += b; break;
                case 3: v += (a + b) >> 1; break;
                case 4: v += paeth(a, b, c); break;
                default: pt_set_error_text("pt_image_load_png: bad filter type"); return PT_ERR_INVALID;
            }
            dst[i] = (uint8_t)v;
        }
    }
    uint8_t* out = (uint8_t*)malloc((size_t)w * h * 4);
    if (!out) { pt_set_error_text("pt_image_load_png: out of memory"); return PT_ERR_INVALID; }
    for (size_t p = 0; p < (size_t)w * h; ++p) {
        const uint8_t* s = &img[p * channels];
        uint8_t r, g, b, a = 255;
        switch (ctype) {
            case 0: r = g = b = s[0]; break;
            case 4: r = g = b = s[0]; a = s[1]; break;
            case 2: r = s[0]; g = s[1]; b = s[2]; break;
            case 6: r = s[0]; g = s[1]; b = s[2]; a = s[3]; break;
            default: {   // palette
                const size_t k = s[0];
                if (3 * k + 2 >= plte.size() + 0 && 3 * k + 2 > plte.size() - 1) { r = g = b = 0; }
                else { r = plte[3 * k]; g = plte[3 * k + 1]; b = plte[3 * k + 2]; }
                if (k < trns.size()) a = trns[k];
            }
        }
        out[4 * p] = r; out[4 * p + 1] = g; out[4 * p + 2] = b; out[4 * p + 3] = a;
    }
    *rgba_out = out; *width = (int)w; *height = (int)h;
    return PT_OK;
}

extern "C" void pt_image_free(uint8_t* rgba) { free(rgba); }
