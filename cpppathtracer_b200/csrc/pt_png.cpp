// pt_png.cpp -- minimal PNG reader for the texture loader of the facade / C ABI.
// Replaces `cv::imread(path)` + `cv::cvtColor(src, dst, CV_BGR2RGBA)` of the reference's
// PocaTextureUtils::AddTexByFile (cuSrc/textures.cu:15-17): OpenCV is a Windows binary dependency there and does not
// exist here.  Supports every colour type and bit depth of the PNG specification (grey 1/2/4/8/16, grey+alpha, RGB and
// RGBA 8/16, palette 1/2/4/8), sequential or Adam7-interlaced; everything is expanded to RGBA8 rows (alpha 255 when the file has none, exactly
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
    const int channels = ctype == 0 ? 1 : ctype == 2 ? 3 : ctype == 3 ? 1 : ctype == 4 ? 2 : ctype == 6 ? 4 : 0;
    const bool depth_ok = (ctype == 0 && (depth == 1 || depth == 2 || depth == 4 || depth == 8 || depth == 16)) ||
                          (ctype == 3 && (depth == 1 || depth == 2 || depth == 4 || depth == 8)) ||
                          ((ctype == 2 || ctype == 4 || ctype == 6) && (depth == 8 || depth == 16));
    if (!w || !h || w > 65535 || h > 65535 || !channels || !depth_ok || interlace > 1 || (ctype == 3 && plte.empty())) {
        pt_set_error_text("pt_image_load_png: unsupported PNG (colour type / bit depth / interlace method)");
        return PT_ERR_INVALID;
    }
    // Adam7 (PNG specification, section 8.2): seven reduced images, each filtered on its own; a non-interlaced file
    // is the single pass (0, 0, 1, 1)
    static const int P7[7][4] = {{0, 0, 8, 8}, {4, 0, 8, 8}, {0, 4, 4, 8}, {2, 0, 4, 4}, {0, 2, 2, 4}, {1, 0, 2, 2}, {0, 1, 1, 2}};
    static const int P1[1][4] = {{0, 0, 1, 1}};
    const int (*passes)[4] = interlace ? P7 : P1;
    const int n_pass = interlace ? 7 : 1;
    const int bits_pp = channels * depth;
    const size_t bpp = (size_t)(bits_pp + 7) / 8;                       // filter distance in bytes
    size_t raw_size = 0;
    for (int k = 0; k < n_pass; ++k) {
        const uint32_t pw = (w - passes[k][0] + passes[k][2] - 1) / passes[k][2], ph = (h - passes[k][1] + passes[k][3] - 1) / passes[k][3];
        if (w > (uint32_t)passes[k][0] && h > (uint32_t)passes[k][1] && pw && ph) raw_size += (((size_t)pw * bits_pp + 7) / 8 + 1) * ph;
    }
    std::vector<uint8_t> raw(raw_size);
    uLongf out_len = (uLongf)raw.size();
    if (uncompress(raw.data(), &out_len, idat.data(), (uLong)idat.size()) != Z_OK || out_len != raw.size()) {
        pt_set_error_text("pt_image_load_png: inflate failed");
        return PT_ERR_INVALID;
    }
    uint8_t* out = (uint8_t*)malloc((size_t)w * h * 4);
    if (!out) { pt_set_error_text("pt_image_load_png: out of memory"); return PT_ERR_INVALID; }
    const int maxv = (1 << (depth > 8 ? 8 : depth)) - 1;
    size_t rp = 0;
    std::vector<uint8_t> cur, prev;
    for (int k = 0; k < n_pass; ++k) {
        if (w <= (uint32_t)passes[k][0] || h <= (uint32_t)passes[k][1]) continue;
        const uint32_t pw = (w - passes[k][0] + passes[k][2] - 1) / passes[k][2], ph = (h - passes[k][1] + passes[k][3] - 1) / passes[k][3];
        if (!pw || !ph) continue;
        const size_t stride = ((size_t)pw * bits_pp + 7) / 8;
        cur.assign(stride, 0); prev.assign(stride, 0);
        for (uint32_t y = 0; y < ph; ++y) {
            // undo the row filter (PNG specification, section 9), byte-wise with the pixel size rounded up to bytes
            const uint8_t ft = raw[rp];
            const uint8_t* src = &raw[rp + 1];
            rp += stride + 1;
            for (size_t i = 0; i < stride; ++i) {
                const int a = i >= bpp ? cur[i - bpp] : 0;
                const int b = prev[i];
                const int c = i >= bpp ? prev[i - bpp] : 0;
                int v = src[i];
                switch (ft) {
                    case 0: break;
                    case 1: v += a; break;
                    case 2: v += b; break;
                    case 3: v += (a + b) >> 1; break;
                    case 4: v += paeth(a, b, c); break;
                    default: free(out); pt_set_error_text("pt_image_load_png: bad filter type"); return PT_ERR_INVALID;
                }
                cur[i] = (uint8_t)v;
            }
            // samples -> RGBA8.  16-bit samples keep their high byte and 1/2/4-bit grey is scaled to 0..255, which is
            // what cv::imread's 8-bit conversion (libpng strip_16 / expand_gray) hands the reference
            const size_t oy = (size_t)passes[k][1] + (size_t)y * passes[k][3];
            for (uint32_t x = 0; x < pw; ++x) {
                int sm[4] = {0, 0, 0, 0};
                for (int ch = 0; ch < channels; ++ch) {
                    const size_t bit = ((size_t)x * channels + ch) * depth;
                    if (depth == 16) sm[ch] = cur[bit >> 3];
                    else if (depth == 8) sm[ch] = cur[bit >> 3];
                    else sm[ch] = (cur[bit >> 3] >> (8 - depth - (bit & 7))) & maxv;
                }
                uint8_t r, g, b, a = 255;
                switch (ctype) {
                    case 0: r = g = b = (uint8_t)(depth < 8 ? sm[0] * 255 / maxv : sm[0]); break;
                    case 4: r = g = b = (uint8_t)sm[0]; a = (uint8_t)sm[1]; break;
                    case 2: r = (uint8_t)sm[0]; g = (uint8_t)sm[1]; b = (uint8_t)sm[2]; break;
                    case 6: r = (uint8_t)sm[0]; g = (uint8_t)sm[1]; b = (uint8_t)sm[2]; a = (uint8_t)sm[3]; break;
                    default: {   // palette
                        const size_t q = (size_t)sm[0];
                        if (3 * q + 2 < plte.size()) { r = plte[3 * q]; g = plte[3 * q + 1]; b = plte[3 * q + 2]; }
                        else { r = g = b = 0; }
                        if (q < trns.size()) a = trns[q];
                    }
                }
                const size_t ox = (size_t)passes[k][0] + (size_t)x * passes[k][2];
                uint8_t* o4 = &out[4 * (oy * w + ox)];
                o4[0] = r; o4[1] = g; o4[2] = b; o4[3] = a;
            }
            cur.swap(prev);
        }
    }
    *rgba_out = out; *width = (int)w; *height = (int)h;
    return PT_OK;
}

extern "C" void pt_image_free(uint8_t* rgba) { free(rgba); }
