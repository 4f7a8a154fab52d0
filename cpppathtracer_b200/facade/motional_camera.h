// Facade of include/motional_camera.h:6-61 + the host half of cuSrc/motional_camera.cu:14-200.
// Same public fields, constructors and controls; the device half (RayGen) lives in libpt_b200.so.
#pragma once
#include <mutex>

#include "ray_tracing_common.h"

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
