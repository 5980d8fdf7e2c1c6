/*
 * embree_record.hpp — TEST-ONLY: the Embree scene-construction calls of src/scene.cpp (which only store
 * pointers) recorded into plain structs, shared by scenref.cpp (loader cross-check) and fullref.cpp (the
 * reference's whole program on the CPU).
 */
#pragma once
#include <cstring>
#include <vector>

#include <embree4/rtcore.h>

namespace {
struct Geom {
    RTCGeometryType type;
    const void *vertices = nullptr, *indices = nullptr;
    size_t n_vertices = 0, n_triangles = 0;
    RTCScene instanced = nullptr;
    float xfm[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    void *user = nullptr;
};
struct Scn {
    std::vector<Geom *> geoms;
};
} // namespace

RTCScene rtcNewScene(RTCDevice) { return (RTCScene) new Scn(); }
RTCGeometry rtcNewGeometry(RTCDevice, RTCGeometryType type) {
    Geom *g = new Geom();
    g->type = type;
    return (RTCGeometry)g;
}
void rtcSetSharedGeometryBuffer(RTCGeometry geometry, RTCBufferType type, unsigned int, RTCFormat, const void *ptr, size_t, size_t,
                                size_t itemCount) {
    Geom *g = (Geom *)geometry;
    if (type == RTC_BUFFER_TYPE_VERTEX) {
        g->vertices = ptr;
        g->n_vertices = itemCount;
    } else {
        g->indices = ptr;
        g->n_triangles = itemCount;
    }
}
void rtcCommitGeometry(RTCGeometry) {}
unsigned int rtcAttachGeometry(RTCScene scene, RTCGeometry geometry) {
    Scn *s = (Scn *)scene;
    s->geoms.push_back((Geom *)geometry);
    return (unsigned int)s->geoms.size() - 1;
}
void rtcCommitScene(RTCScene) {}
void rtcSetGeometryTimeStepCount(RTCGeometry, unsigned int) {}
void rtcSetGeometryInstancedScene(RTCGeometry geometry, RTCScene scene) { ((Geom *)geometry)->instanced = scene; }
void rtcSetGeometryTransform(RTCGeometry geometry, unsigned int, RTCFormat, const void *xfm) { memcpy(((Geom *)geometry)->xfm, xfm, 64); }
void rtcSetGeometryUserData(RTCGeometry geometry, void *ptr) { ((Geom *)geometry)->user = ptr; }
