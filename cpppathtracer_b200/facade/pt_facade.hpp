// pt_facade.hpp -- header-only C++ facade over the C ABI (include/pt_api.h): the class names, public fields and
// methods the reference's caller uses (PathTracer, Object, Material, MotionalCamera, SceneBVH, PocaTextureUtils,
// DispatchRayArgs; reference: include/{path_tracer,object,material,motional_camera,bvh,textures,ray_tracing_common}.h),
// implemented from scratch on top of libpt_b200.so.  Host-only: no CUDA toolkit is needed to compile a caller.
// The forwarding headers next to this file carry the reference's header names so its caller's #include lines
// keep working.  See INTEGRATION.md.
#pragma once

// ---------------------------------------------------------------------------------------------
// section: ray_tracing_common.h
// Facade of the reference's include/ray_tracing_common.h + the float3 helpers its caller uses
// (make_float3 and friends come from Common/helper_math.h / cuda_runtime.h in the reference).
// Header-only, host-only: no CUDA toolkit is needed to compile a caller against libpt_b200.so.
#include <stdint.h>
#include <cmath>

#include "../../include/pt_api.h"

#ifndef __VECTOR_TYPES_H__   // do not clash with cuda_runtime.h when the caller also includes it
struct float3 { float x, y, z; };
inline float3 make_float3(float x, float y, float z) { float3 r = {x, y, z}; return r; }
#endif
inline float3 make_float3(float s) { return make_float3(s, s, s); }

typedef uint32_t uint;

#define DEFAULT_RAY_TMAX 1e30f   // include/ray_tracing_common.h:11
#define BOUNCE_RAY_TMIN 2e-5f    // include/ray_tracing_common.h:12

// include/ray_tracing_common.h:37-40
struct DispatchRayArgs {
    void* cbParam;
    void (*Callback)(uint8_t* data, int width, int height, void* cbParam);
};

namespace pt_facade {
// fp32 vector helpers with the operand order of Common/helper_math.h (host fallbacks, :78-81,1264,1325,1436)
inline float3 add(float3 a, float3 b) { return make_float3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline float3 sub(float3 a, float3 b) { return make_float3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline float3 scale(float3 a, float s) { return make_float3(a.x * s, a.y * s, a.z * s); }
inline float3 neg(float3 a) { return make_float3(-a.x, -a.y, -a.z); }
inline float dot(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline float3 cross(float3 a, float3 b) { return make_float3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
inline float3 normalize(float3 v) { float inv = 1.0f / sqrtf(dot(v, v)); return scale(v, inv); }
}  // namespace pt_facade

// ---------------------------------------------------------------------------------------------
// section: material.h
// Facade of include/material.h:5-35 -- same enum, same public fields, plain data filled by the caller.

namespace MaterialType {
enum Enum { Diffuse = 0, Metal, Mirror, Glass, Test, Count };
}

typedef unsigned long long cudaTextureObject_facade_t;

class Material {
public:
    MaterialType::Enum type_;
    bool have_tex_;
    union {
        float3 kd_;
        cudaTextureObject_facade_t tex_;   // handle from PocaTextureUtils::AddTexByFile / pt_texture_create_rgba8
    };
    float refractive_index_;
    float emit_intensity_;
    float smoothness_;
    float reflectivity_;

    pt_material ToAbi() const {
        pt_material m;
        m.type = (int32_t)type_;
        m.have_tex = have_tex_ ? 1 : 0;
        m.kd[0] = have_tex_ ? 0.f : kd_.x; m.kd[1] = have_tex_ ? 0.f : kd_.y; m.kd[2] = have_tex_ ? 0.f : kd_.z;
        m.refractive_index = refractive_index_; m.emit_intensity = emit_intensity_;
        m.smoothness = smoothness_; m.reflectivity = reflectivity_;
        m.tex = have_tex_ ? (uint64_t)tex_ : 0;
        return m;
    }
};

// ---------------------------------------------------------------------------------------------
// section: object.h
// Facade of include/object.h:7-32 -- the tagged primitive the caller fills by hand after memset(0).

namespace PrimitiveType {
enum Enum { Sphere, Platform, Cylinder, Count };
}

class Object {
public:
    // Object::GetAABBMin/Max, cuSrc/object.cu:134-170 (host functions in the reference too)
    float3 GetAABBMin() const {
        const float tol = BOUNCE_RAY_TMIN * 5.f;
        const float ar = radius_ >= 0 ? radius_ : -radius_;
        switch (type_) {
            case PrimitiveType::Sphere: return make_float3(center_.x - ar, center_.y - ar, center_.z - ar);
            case PrimitiveType::Platform: return make_float3(-DEFAULT_RAY_TMAX * 5, y_pos_ - tol, -DEFAULT_RAY_TMAX * 5);
            case PrimitiveType::Cylinder: return make_float3(center_.x - ar, center_.y - height_ / 2 - tol, center_.z - ar);
            default: return make_float3(0.f);
        }
    }
    float3 GetAABBMax() const {
        const float tol = BOUNCE_RAY_TMIN * 5.f;
        const float ar = radius_ >= 0 ? radius_ : -radius_;
        switch (type_) {
            case PrimitiveType::Sphere: return make_float3(center_.x + ar, center_.y + ar, center_.z + ar);
            case PrimitiveType::Platform: return make_float3(DEFAULT_RAY_TMAX * 5, y_pos_ + tol, DEFAULT_RAY_TMAX * 5);
            case PrimitiveType::Cylinder: return make_float3(center_.x + ar, center_.y + height_ / 2 + tol, center_.z + ar);
            default: return make_float3(0.f);
        }
    }

    PrimitiveType::Enum type_;
    Material material_;
    float3 center_;
    float radius_;   // Sphere and Cylinder
    float y_pos_;    // Platform
    float height_;   // Cylinder
};

// ---------------------------------------------------------------------------------------------
// section: motional_camera.h
// Facade of include/motional_camera.h:6-61 + the host half of cuSrc/motional_camera.cu:14-200.
// Same public fields, constructors and controls; the device half (RayGen) lives in libpt_b200.so.
#include <mutex>


class MotionalCamera {
public:
    float3 vup = make_float3(0.0f, 1.0f, 0.0f);
    int width_;
    int height_;
    uint cur_sample_idx_;
    float3 origin_;
    float3 look_at_;
    float view_fov_ = 30;
    float dist_to_focus_ = 100;
    float lens_radius_ = 0.0005f;
    float move_speed_ = 50.f;
    float3 u_, v_, w_;
    float3 top_left_corner_;
    float3 horizontal_;
    float3 vertical_;

    MotionalCamera() : width_(1920), height_(1080), cur_sample_idx_(0), origin_(make_float3(0.f)), look_at_(make_float3(0.f, 0.f, 1.f)) {}
    MotionalCamera(int width, int height) : width_(width), height_(height), cur_sample_idx_(0), origin_(make_float3(0.f)), look_at_(make_float3(0.f, 0.f, 1.f)) {}
    MotionalCamera(int width, int height, float3 ori, float3 at) : width_(width), height_(height), cur_sample_idx_(0), origin_(ori), look_at_(at) {}
    ~MotionalCamera() {}

    void Resize(int width, int height) { width_ = width; height_ = height; }
    void SetOrigin(float3 ori) { origin_ = ori; }
    void SetOrigin(float x, float y, float z) { origin_ = make_float3(x, y, z); }
    void SetLookAt(float3 at) { look_at_ = at; }
    void SetLookAt(float x, float y, float z) { look_at_ = make_float3(x, y, z); }
    void SetViewFov(float fov) { view_fov_ = fov; }

    // cuSrc/motional_camera.cu:76-118
    void MoveEyeLeft(float c = 1.f) { Shift(Left(), c * move_speed_); }
    void MoveEyeRight(float c = 1.f) { Shift(Left(), -(c * move_speed_)); }
    void MoveEyeForward(float c = 1.f) { Shift(Back(), -(c * move_speed_)); }
    void MoveEyeBackward(float c = 1.f) { Shift(Back(), c * move_speed_); }
    void MoveEyeUp(float c = 1.f) { Shift(vup, c * move_speed_); }
    void MoveEyeDown(float c = 1.f) { Shift(vup, -(c * move_speed_)); }
    // cuSrc/motional_camera.cu:120-164
    void RotateAroundUp(float dy) { Rotate(true, dy); }
    void RotateAroundDown(float dy) { Rotate(true, -dy); }
    void RotateAroundLeft(float dx) { Rotate(false, dx); }
    void RotateAroundRight(float dx) { Rotate(false, -dx); }
    // cuSrc/motional_camera.cu:166-168 (adds a radians-scaled delta to a degrees field; reproduced)
    void ScaleFov(float d) { view_fov_ = (float)((double)view_fov_ + (double)d * 3.14159265358979323846 / 180.0f); }

    // cuSrc/motional_camera.cu:177-200: derive the frame basis, bump cur_sample_idx_, return a snapshot
    MotionalCamera GetCopy() {
        std::lock_guard<std::mutex> lock(Mutex());
        pt_camera c = ToAbi();
        float b[19];
        pt_camera_basis(&c, b);
        u_ = make_float3(b[0], b[1], b[2]); v_ = make_float3(b[3], b[4], b[5]); w_ = make_float3(b[6], b[7], b[8]);
        top_left_corner_ = make_float3(b[9], b[10], b[11]);
        horizontal_ = make_float3(b[12], b[13], b[14]); vertical_ = make_float3(b[15], b[16], b[17]);
        dist_to_focus_ = b[18];
        cur_sample_idx_++;
        return *this;
    }
    void Refresh() { cur_sample_idx_ = 0; }     // :43-45
    void Lock() { Mutex().lock(); }             // :170-175 (one process-wide mutex, like the reference)
    void Unlock() { Mutex().unlock(); }

    pt_camera ToAbi() const {
        pt_camera c;
        c.width = width_; c.height = height_;
        c.origin[0] = origin_.x; c.origin[1] = origin_.y; c.origin[2] = origin_.z;
        c.look_at[0] = look_at_.x; c.look_at[1] = look_at_.y; c.look_at[2] = look_at_.z;
        c.vup[0] = vup.x; c.vup[1] = vup.y; c.vup[2] = vup.z;
        c.view_fov_deg = view_fov_; c.lens_radius = lens_radius_;
        return c;
    }

private:
    static std::mutex& Mutex() { static std::mutex m; return m; }
    float3 Left() const {
        using namespace pt_facade;
        return neg(normalize(cross(vup, normalize(sub(origin_, look_at_)))));
    }
    float3 Back() const {
        using namespace pt_facade;
        return neg(normalize(cross(Left(), vup)));
    }
    void Shift(float3 dir, float amount) {
        using namespace pt_facade;
        // `origin_ += c*speed*dir` / `origin_ -= c*speed*dir`: subtracting is adding the exactly negated product
        origin_ = add(origin_, scale(dir, amount));
        look_at_ = add(look_at_, scale(dir, amount));
    }
    void Rotate(bool vertical, float delta) {
        using namespace pt_facade;
        look_at_ = add(origin_, normalize(sub(look_at_, origin_)));
        float3 w = normalize(sub(look_at_, origin_));
        float3 left = normalize(cross(vup, w));
        float3 axis = vertical ? normalize(cross(w, left)) : left;
        look_at_ = add(look_at_, scale(axis, delta));
        look_at_ = add(origin_, normalize(sub(look_at_, origin_)));
    }
};

// ---------------------------------------------------------------------------------------------
// section: textures.h
// Facade of include/textures.h:7-16.  The reference decodes the file with OpenCV (cuSrc/textures.cu:15-17);
// there is no image decoder here, so the facade accepts binary PPM (P6) files and raw RGBA8 buffers.
#include <cstdio>
#include <string>
#include <vector>


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
    // PocaTextureUtils::AddTexByFile (cuSrc/textures.cu:14-62): PNG through the library's reader, PPM as a fallback
    static unsigned long long AddTexByFile(pt_scene* scene, const std::string& path) {
        uint8_t* px = nullptr; int w = 0, h = 0;
        if (pt_image_load_png(path.c_str(), &px, &w, &h) == PT_OK) {
            unsigned long long t = AddTexByRgba8(scene, px, w, h);
            pt_image_free(px);
            return t;
        }
        std::vector<uint8_t> rgba;
        if (!ReadPpm(path, rgba, w, h)) return 0ull;
        return AddTexByRgba8(scene, rgba.data(), w, h);
    }
    static void DestroyTexture(pt_scene* scene, unsigned long long tex) { pt_texture_destroy(scene, (uint64_t)tex); }
};

// ---------------------------------------------------------------------------------------------
// section: bvh.h
// Facade of include/bvh.h:8-39.  The reference keeps ONE scene in file-static state (cuSrc/bvh.cu:16-20) behind
// static member functions; the facade keeps that shape (a process-wide current scene) so callers compile
// unchanged, while PathTracer owns the pt_scene handle.
#include <unordered_map>
#include <vector>


class SceneBVH;
typedef pt_scene* SceneBVHGPUHandle;

class SceneBVH {
public:
    static void Bind(pt_scene* scene) { State().scene = scene; }
    // cuSrc/bvh.cu:22-29: null and duplicate pointers are ignored; the pointer is kept for UpdateObject
    static void AddObject(Object* obj) {
        S& s = State();
        if (obj == nullptr || s.ids.count(obj)) return;
        int id = -1;
        const float c[3] = {obj->center_.x, obj->center_.y, obj->center_.z};
        switch (obj->type_) {
            case PrimitiveType::Platform: id = pt_add_platform(s.scene, obj->y_pos_); break;
            case PrimitiveType::Cylinder: id = pt_add_cylinder(s.scene, c, obj->radius_, obj->height_); break;
            default: id = pt_add_sphere(s.scene, c, obj->radius_); break;   // value-initialised type_ is Sphere
        }
        if (id < 0) return;
        pt_material m = obj->material_.ToAbi();
        pt_set_material(s.scene, id, &m);
        s.ids[obj] = id;
    }
    // cuSrc/bvh.cu:116-120
    static SceneBVHGPUHandle BuildBVH() { return pt_build(State().scene) == PT_OK ? State().scene : nullptr; }
    // cuSrc/bvh.cu:144-157: copy the object's current fields and refit; here: update + device rebuild
    static void UpdateObject(Object* obj) {
        S& s = State();
        auto it = s.ids.find(obj);
        if (it == s.ids.end()) return;
        const float c[3] = {obj->center_.x, obj->center_.y, obj->center_.z};
        pt_update_prim(s.scene, it->second, c, obj->radius_, obj->y_pos_, obj->height_);
        pt_material m = obj->material_.ToAbi();
        pt_set_material(s.scene, it->second, &m);
        pt_build(s.scene);
    }
    static void ReleaseBVH() { State().ids.clear(); }   // cuSrc/bvh.cu:159-165
    static int PrimitiveId(Object* obj) { auto it = State().ids.find(obj); return it == State().ids.end() ? -1 : it->second; }

private:
    struct S { pt_scene* scene = nullptr; std::unordered_map<Object*, int> ids; };
    static S& State() { static S s; return s; }
};

// ---------------------------------------------------------------------------------------------
// section: path_tracer.h
// Facade of include/path_tracer.h:15-26: the class VideoRenderer drives (cppSrc/video_renderer.cpp:33-38,81,119,163).
// Same five methods.  Differences, all documented in INTEGRATION.md:
//   * DispatchRay is synchronous by default: it renders one sample, filters, mixes, copies the BGRA frame to the host
//     and calls args.Callback before returning (the reference does this on a detached thread, path_tracer.cu:256-306);
//     SetPipelined(true) makes it asynchronous -- the callback of a dispatch runs inside the next one (or Flush);
//   * errors are kept in LastError() instead of a log file;
//   * the sky comes from SetSkyRgba8() instead of a hard-wired "textures\\sky.png" (path_tracer.cu:47).
#include <memory>
#include <string>
#include <vector>


#define MAX_RECURSION_DEPTH_SET 32

class PathTracer {
public:
    PathTracer() { if (pt_scene_create(&scene_) != PT_OK) { scene_ = nullptr; err_ = pt_last_error(); } SceneBVH::Bind(scene_); }
    ~PathTracer() { if (scene_) { Flush(); pt_scene_destroy(scene_); } SceneBVH::ReleaseBVH(); SceneBVH::Bind(nullptr); }
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
        if (pipelined_) {
            // asynchronous like the reference (task queue + render thread, path_tracer.cu:256-319), without a thread: the
            // sample is enqueued and the call returns; its callback runs inside the NEXT DispatchRay (or Flush), by when
            // its frame has been copied under that next sample's trace
            if (pt_dispatch_ray_begin(scene_, &c, max_recursion_depth_, seed_, cam.cur_sample_idx_) != PT_OK) { err_ = pt_last_error(); return; }
            pending_.push_back(Pending{args, c.width, c.height});
            if (pending_.size() == 2) DeliverOldest();
            return;
        }
        // the frame lives in the scene's page-locked buffer, valid during the callback like the reference's
        // output_buffer_ (path_tracer.cu:303-304)
        const uint8_t* frame = nullptr; size_t nbytes = 0;
        if (pt_dispatch_ray_frame(scene_, &c, max_recursion_depth_, seed_, cam.cur_sample_idx_, &frame, &nbytes) != PT_OK) {
            err_ = pt_last_error();
            return;
        }
        if (args.Callback) args.Callback(const_cast<uint8_t*>(frame), c.width, c.height, args.cbParam);
    }

    // additions (no counterpart in the reference)
    bool SetSkyRgba8(const uint8_t* rgba, int w, int h) { return scene_ && pt_set_sky_rgba8(scene_, rgba, w, h) == PT_OK; }
    // the reference loads "textures\\sky.png" itself (path_tracer.cu:47); here the caller names the file
    bool SetSkyFile(const std::string& png_path) {
        uint8_t* px = nullptr; int w = 0, h = 0;
        if (!scene_ || pt_image_load_png(png_path.c_str(), &px, &w, &h) != PT_OK) { err_ = pt_last_error(); return false; }
        const bool ok = pt_set_sky_rgba8(scene_, px, w, h) == PT_OK;
        pt_image_free(px);
        return ok;
    }
    // pipelined DispatchRay (see there); Flush delivers the frames still in flight
    void SetPipelined(bool on) { if (!on) Flush(); pipelined_ = on; }
    void Flush() { while (!pending_.empty()) DeliverOldest(); }
    void SetSeed(uint64_t seed) { seed_ = seed; }
    pt_scene* Handle() { return scene_; }
    const std::string& LastError() const { return err_; }

private:
    struct Pending { DispatchRayArgs args; int width, height; };
    void DeliverOldest() {
        const uint8_t* frame = nullptr; size_t nbytes = 0;
        const Pending p = pending_.front();
        pending_.erase(pending_.begin());
        if (pt_dispatch_ray_end(scene_, &frame, &nbytes) != PT_OK) { err_ = pt_last_error(); return; }
        if (p.args.Callback) p.args.Callback(const_cast<uint8_t*>(frame), p.width, p.height, p.args.cbParam);
    }
    std::vector<Pending> pending_;
    bool pipelined_ = false;
    pt_scene* scene_ = nullptr;
    std::shared_ptr<MotionalCamera> camera_;
    uint8_t max_recursion_depth_ = 8;     // path_tracer.h:43
    uint64_t seed_ = 0;
    std::string err_;
};
