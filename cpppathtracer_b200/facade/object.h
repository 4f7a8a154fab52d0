// Facade of include/object.h:7-32 -- the tagged primitive the caller fills by hand after memset(0).
#pragma once
#include "material.h"

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
