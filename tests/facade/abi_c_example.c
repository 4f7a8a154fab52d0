/* The C snippet of INTEGRATION.md section 2, compiled as C99 against include/pt_api.h: the boundary is a C ABI
 * (plain pointers and sizes), not a C++ one.  Returns 0 when the scene calls succeed, 3 when there is no device. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/pt_api.h"

int main(void) {
    pt_scene* s = NULL;
    int rc = pt_scene_create(&s);
    if (rc == PT_ERR_NO_DEVICE) { printf("NO_DEVICE %s\n", pt_last_error()); return 3; }
    if (rc != PT_OK) { printf("ERR %s\n", pt_last_error()); return 1; }
    float c[3] = {0.f, 1.f, 0.f};
    int id = pt_add_sphere(s, c, 1.0f);
    pt_material m;
    memset(&m, 0, sizeof(m));
    m.type = PT_MAT_GLASS; m.kd[0] = m.kd[1] = m.kd[2] = 1.f; m.refractive_index = 1.5f; m.smoothness = 5.f;
    pt_set_material(s, id, &m);
    pt_add_platform(s, 0.f);
    unsigned char sky[4 * 4 * 4];
    memset(sky, 200, sizeof(sky));
    pt_set_sky_rgba8(s, sky, 4, 4);
    if (pt_build(s) != PT_OK) { printf("ERR %s\n", pt_last_error()); return 1; }
    pt_camera cam = {64, 36, {13.f, 10.f, 13.f}, {0.f, 0.f, 0.f}, {0.f, 1.f, 0.f}, 30.f, 0.0005f};
    if (pt_render_frame(s, &cam, 4, 8, 1u, 0u) != PT_OK) { printf("ERR %s\n", pt_last_error()); return 1; }
    float* img = (float*)malloc(sizeof(float) * 4 * 64 * 36);
    if (pt_read_accum(s, img, (size_t)4 * 64 * 36) != PT_OK) { printf("ERR %s\n", pt_last_error()); return 1; }
    double sum = 0.0;
    for (int i = 0; i < 64 * 36; ++i) sum += img[4 * i] + img[4 * i + 1] + img[4 * i + 2];
    /* an object moves (SceneBVH::UpdateObject), the frame is rendered again and read back in two halves: the copy of a
     * frame may run under the next frame's build and render (pt_read_accum_begin / _end) */
    float c2[3] = {0.5f, 1.f, -0.5f};
    float* img2 = (float*)malloc(sizeof(float) * 4 * 64 * 36);
    if (pt_update_prim(s, id, c2, 1.0f, 0.f, 0.f) != PT_OK || pt_build(s) != PT_OK || pt_reset_accum(s) != PT_OK ||
        pt_render_frame_async(s, &cam, 4, 8, 1u, 0u) != PT_OK || pt_read_accum_begin(s, img2, (size_t)4 * 64 * 36) != PT_OK ||
        pt_read_accum_end(s) != PT_OK) { printf("ERR %s\n", pt_last_error()); return 1; }
    int moved = memcmp(img, img2, sizeof(float) * 4 * 64 * 36) != 0;
    pt_stats st;
    pt_get_stats(s, &st);
    printf("OK samples %.0f mean %.4f version %s moved %d build_ms %.3f\n", img[3], sum / (3.0 * 64 * 36), pt_version(), moved, st.build_ms);
    free(img); free(img2);
    pt_scene_destroy(s);
    return 0;
}
