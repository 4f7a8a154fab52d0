// Drives the C++ facade the way the reference's VideoRenderer::Create / OnRender do
// (cppSrc/video_renderer.cpp:32-120,147-164): scene by hand-filled Objects, InitPipeline, DispatchRay with a
// callback, camera controls under Lock/Unlock, UpdateObject.  Prints a few numbers the Python test checks.
#include <cstdio>
#include <cstring>
#include <thread>

#include "../../cpppathtracer_b200/facade/path_tracer.h"

struct Frame { int calls = 0, w = 0, h = 0; unsigned long long sum = 0; uint8_t first[4] = {0, 0, 0, 0}; };
static void OnFrame(uint8_t* data, int width, int height, void* cb) {
    Frame* f = (Frame*)cb;
    f->calls++; f->w = width; f->h = height; f->sum = 0;
    for (size_t i = 0; i < (size_t)width * height * 4; ++i) f->sum += data[i];
    memcpy(f->first, data, 4);
}

int main() {
    std::unique_ptr<PathTracer> tracer(new PathTracer);
    if (!tracer->Handle()) { printf("NO_DEVICE %s\n", tracer->LastError().c_str()); return 3; }
    tracer->SetCamera(std::shared_ptr<MotionalCamera>(new MotionalCamera(320, 176, make_float3(130.f, 103.f, 130.f), make_float3(0.f, 0.f, 0.f))));
    Material* white = new Material; memset(white, 0, sizeof(Material));
    white->type_ = MaterialType::Diffuse; white->kd_ = make_float3(0.95f, 0.95f, 0.95f);
    Material* glass = new Material; memset(glass, 0, sizeof(Material));
    glass->type_ = MaterialType::Glass; glass->kd_ = make_float3(1.f); glass->smoothness_ = 5.f; glass->refractive_index_ = 1.5f;
    Object* floor = new Object(); memset(floor, 0, sizeof(Object));
    floor->material_ = *white; floor->type_ = PrimitiveType::Platform; floor->y_pos_ = 0.f;
    tracer->AddObject(floor);
    tracer->AddObject(floor);                       // duplicates are ignored (bvh.cu:23-25)
    Object* ball = new Object(); memset(ball, 0, sizeof(Object));
    ball->type_ = PrimitiveType::Sphere; ball->material_ = *glass; ball->radius_ = 20.f; ball->center_ = make_float3(0.f, 20.f, 0.f);
    tracer->AddObject(ball);
    Object* inball = new Object();                  // value-initialised: type_ stays Sphere (video_renderer.cpp:94-98)
    inball->material_ = ball->material_; inball->center_ = ball->center_; inball->radius_ = 0.01f - ball->radius_;
    tracer->AddObject(inball);
    Object* can = new Object(); memset(can, 0, sizeof(Object));
    can->type_ = PrimitiveType::Cylinder; can->material_ = *white; can->radius_ = 8.f; can->height_ = 12.f; can->center_ = make_float3(40.f, 6.f, -10.f);
    tracer->AddObject(can);
    std::vector<uint8_t> sky(64 * 64 * 4, 200);
    tracer->SetSkyRgba8(sky.data(), 64, 64);
    tracer->InitPipeline();
    printf("ids %d %d %d %d\n", SceneBVH::PrimitiveId(floor), SceneBVH::PrimitiveId(ball), SceneBVH::PrimitiveId(inball), SceneBVH::PrimitiveId(can));

    Frame fr; DispatchRayArgs args; args.cbParam = &fr; args.Callback = OnFrame;
    for (int i = 0; i < 4; ++i) tracer->DispatchRay(args);
    printf("frames %d %dx%d idx %u sum %llu err '%s'\n", fr.calls, fr.w, fr.h, tracer->GetCamera()->cur_sample_idx_, fr.sum, tracer->LastError().c_str());
    unsigned long long before = fr.sum;

    // camera controls under the lock, then Refresh restarts the running mean (video_renderer.cpp:147-164,192-226)
    auto cam = tracer->GetCamera();
    cam->Lock(); cam->RotateAroundRight(0.2f); cam->MoveEyeLeft(0.1f); cam->ScaleFov(5.f); cam->Refresh(); cam->Unlock();
    tracer->DispatchRay(args);
    printf("after_move idx %u changed %d origin %.4f %.4f %.4f fov %.5f\n", cam->cur_sample_idx_, fr.sum != before, cam->origin_.x, cam->origin_.y, cam->origin_.z, cam->view_fov_);

    // moving an object: edit the Object, UpdateObject (bvh.cu:144-157), restart the mean
    before = fr.sum;
    ball->center_.y = 60.f; SceneBVH::UpdateObject(ball);
    cam->Refresh();
    tracer->DispatchRay(args);
    printf("after_update changed %d\n", fr.sum != before);

    // the reference runs DispatchRay's work on its own thread; a caller can keep that shape
    std::thread t([&] { tracer->DispatchRay(args); });
    t.join();
    printf("threaded frames %d\n", fr.calls);

    // pipelined mode: the callback of dispatch i runs inside dispatch i+1; Flush delivers the last one
    const int base = fr.calls;
    tracer->SetPipelined(true);
    tracer->DispatchRay(args); const int after1 = fr.calls - base;
    tracer->DispatchRay(args); const int after2 = fr.calls - base;
    tracer->DispatchRay(args); const int after3 = fr.calls - base;
    tracer->Flush();
    printf("pipelined %d %d %d %d err '%s'\n", after1, after2, after3, fr.calls - base, tracer->LastError().c_str());
    tracer->SetPipelined(false);
    return 0;
}
