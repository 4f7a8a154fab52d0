// Facade of include/path_tracer.h:15-26: the class VideoRenderer drives (cppSrc/video_renderer.cpp:33-38,81,119,163).
// Same five methods.  Differences, all documented in INTEGRATION.md:
//   * DispatchRay is synchronous: it renders one sample, filters, mixes, copies the BGRA frame to the host and
//     calls args.Callback before returning (the reference does this on a detached thread, path_tracer.cu:256-306);
//   * errors are kept in LastError() instead of a log file;
//   * the sky comes from SetSkyRgba8() instead of a hard-wired "textures\\sky.png" (path_tracer.cu:47).
#pragma once
#include <memory>
#include <string>
#include <vector>

#include "bvh.h"
#include "motional_camera.h"
#include "object.h"
#include "textures.h"

#define MAX_RECURSION_DEPTH_SET 32

class PathTracer {
public:
    PathTracer() { if (pt_scene_create(&scene_) != PT_OK) { scene_ = nullptr; err_ = pt_last_error(); } SceneBVH::Bind(scene_); }
    ~PathTracer() { if (scene_) pt_scene_destroy(scene_); SceneBVH::ReleaseBVH(); SceneBVH::Bind(nullptr); }
    PathTracer(const PathTracer&) = delete;
    PathTracer& operator=(const PathTracer&) = delete;

    void AddObject(Object* obj) { SceneBVH::AddObject(obj); }                    // path_tracer.cu:29-34
    void InitPipeline() { if (!SceneBVH::BuildBVH()) err_ = pt_last_error(); }    // path_tracer.cu:308-314
    void SetCamera(std::shared_ptr<MotionalCamera>& camera) { camera_ = camera; }
    void SetCamera(std::shared_ptr<MotionalCamera>&& camera) { camera_ = camera; }
    std::shared_ptr<MotionalCamera> GetCamera() { return camera_; }

    // path_tracer.cu:316-319 + one PipelineLoop iteration (:260-305)
    void DispatchRay(DispatchRayArgs args) {
        if (!scene_ || !camera_) return;
        MotionalCamera cam = camera_->GetCopy();          // increments cur_sample_idx_ like the reference
        pt_camera c = cam.ToAbi();
        output_.resize((size_t)c.width * c.height * 4);
        if (pt_dispatch_ray(scene_, &c, max_recursion_depth_, seed_, cam.cur_sample_idx_, output_.data(), output_.size()) != PT_OK) {
            err_ = pt_last_error();
            return;
        }
        if (args.Callback) args.Callback(output_.data(), c.width, c.height, args.cbParam);
    }

    // additions (no counterpart in the reference)
    bool SetSkyRgba8(const uint8_t* rgba, int w, int h) { return scene_ && pt_set_sky_rgba8(scene_, rgba, w, h) == PT_OK; }
    void SetSeed(uint64_t seed) { seed_ = seed; }
    pt_scene* Handle() { return scene_; }
    const std::string& LastError() const { return err_; }

private:
    pt_scene* scene_ = nullptr;
    std::shared_ptr<MotionalCamera> camera_;
    std::vector<uint8_t> output_;
    uint8_t max_recursion_depth_ = 8;     // path_tracer.h:43
    uint64_t seed_ = 0;
    std::string err_;
};
