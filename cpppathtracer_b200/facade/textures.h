// Facade of include/textures.h:7-16.  The reference decodes the file with OpenCV (cuSrc/textures.cu:15-17);
// there is no image decoder here, so the facade accepts binary PPM (P6) files and raw RGBA8 buffers.
#pragma once
#include <cstdio>
#include <string>
#include <vector>

#include "ray_tracing_common.h"

class PocaTextureUtils {
public:
    // texture for Material::tex_; bound to the scene given (the reference's textures are process-global)
    static unsigned long long AddTexByRgba8(pt_scene* scene, const uint8_t* rgba, int width, int height) {
        uint64_t h = 0;
        return pt_texture_create_rgba8(scene, rgba, width, height, &h) == PT_OK ? (unsigned long long)h : 0ull;
    }
    static bool ReadPpm(const std::string& path, std::vector<uint8_t>& rgba, int& w, int& h) {
        FILE* f = fopen(path.c_str(), "rb");
        if (!f) return false;
        int maxv = 0;
        bool ok = fscanf(f, "P6 %d %d %d", &w, &h, &maxv) == 3 && maxv == 255 && fgetc(f) != EOF;
        if (ok) {
            std::vector<uint8_t> rgb((size_t)w * h * 3);
            ok = fread(rgb.data(), 1, rgb.size(), f) == rgb.size();
            rgba.resize((size_t)w * h * 4);
            for (size_t i = 0; ok && i < (size_t)w * h; ++i) {
                rgba[4 * i] = rgb[3 * i]; rgba[4 * i + 1] = rgb[3 * i + 1]; rgba[4 * i + 2] = rgb[3 * i + 2]; rgba[4 * i + 3] = 255;
            }
        }
        fclose(f);
        return ok;
    }
    static unsigned long long AddTexByFile(pt_scene* scene, const std::string& ppm_path) {
        std::vector<uint8_t> rgba; int w = 0, h = 0;
        if (!ReadPpm(ppm_path, rgba, w, h)) return 0ull;
        return AddTexByRgba8(scene, rgba.data(), w, h);
    }
    static void DestroyTexture(pt_scene* scene, unsigned long long tex) { pt_texture_destroy(scene, (uint64_t)tex); }
};
