// Facade of include/material.h:5-35 -- same enum, same public fields, plain data filled by the caller.
#pragma once
#include "ray_tracing_common.h"

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
