// Applies a list of MotionalCamera controls (op arg per line on stdin) through the C++ facade and prints the pose
// after each one as raw hex floats; tests/test_facade.py compares them with the poses the reference's own
// MotionalCamera produced (tests/golden/sky_camera.npz).
#include <cstdio>
#include <cstring>

#include "../../cpppathtracer_b200/facade/motional_camera.h"

static unsigned bits(float f) { unsigned u; memcpy(&u, &f, 4); return u; }

int main() {
    MotionalCamera cam(16, 16, make_float3(130.f, 103.f, 130.f), make_float3(0.f, 0.f, 0.f));
    cam.SetViewFov(30.f);
    int op; float arg;
    while (scanf("%d %f", &op, &arg) == 2) {
        switch (op) {
            case 0: cam.MoveEyeLeft(arg); break;      case 1: cam.MoveEyeRight(arg); break;
            case 2: cam.MoveEyeForward(arg); break;   case 3: cam.MoveEyeBackward(arg); break;
            case 4: cam.MoveEyeUp(arg); break;        case 5: cam.MoveEyeDown(arg); break;
            case 6: cam.RotateAroundUp(arg); break;   case 7: cam.RotateAroundDown(arg); break;
            case 8: cam.RotateAroundLeft(arg); break; case 9: cam.RotateAroundRight(arg); break;
            case 10: cam.ScaleFov(arg); break;
        }
        printf("%08x %08x %08x %08x %08x %08x %08x\n", bits(cam.origin_.x), bits(cam.origin_.y), bits(cam.origin_.z),
               bits(cam.look_at_.x), bits(cam.look_at_.y), bits(cam.look_at_.z), bits(cam.view_fov_));
    }
    MotionalCamera snap = cam.GetCopy();
    printf("basis %08x %08x %08x idx %u\n", bits(snap.u_.x), bits(snap.top_left_corner_.y), bits(snap.dist_to_focus_), snap.cur_sample_idx_);
    return 0;
}
