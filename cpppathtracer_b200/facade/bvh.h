// Facade of include/bvh.h:8-39.  The reference keeps ONE scene in file-static state (cuSrc/bvh.cu:16-20) behind
// static member functions; the facade keeps that shape (a process-wide current scene) so callers compile
// unchanged, while PathTracer owns the pt_scene handle.
#pragma once
#include <unordered_map>
#include <vector>

#include "object.h"

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
