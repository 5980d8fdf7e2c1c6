/*
 * embree4/rtcore.h — TEST-ONLY shim of the Embree 4 declarations the reference's hot path uses
 * (src/trace_ray.hpp:18-30, src/camera.hpp:47-62, src/app.hpp:46-57, src/scene.hpp). Embree itself is a
 * third-party dependency that is not available here; rtcIntersect1 is forwarded to the oracle's
 * brute-force closest-hit search (refshim.cpp), everything else is a stub.
 */
#pragma once
#include <cstdint>
#define RTC_INVALID_GEOMETRY_ID ((unsigned int)-1)
#define RTC_MAX_INSTANCE_LEVEL_COUNT 1
struct RTCRay {
    float org_x, org_y, org_z, tnear;
    float dir_x, dir_y, dir_z, time;
    float tfar;
    unsigned int mask, id, flags;
};
struct RTCHit {
    float Ng_x, Ng_y, Ng_z;
    float u, v;
    unsigned int primID, geomID;
    unsigned int instID[RTC_MAX_INSTANCE_LEVEL_COUNT];
};
struct RTCRayHit {
    RTCRay ray;
    RTCHit hit;
};
typedef struct RTCSceneTy *RTCScene;
typedef struct RTCDeviceTy *RTCDevice;
typedef struct RTCGeometryTy *RTCGeometry;
void rtcIntersect1(RTCScene scene, RTCRayHit *rayhit);
void *rtcGetGeometryUserDataFromScene(RTCScene scene, unsigned int geomID);
inline void rtcReleaseGeometry(RTCGeometry) {}
inline void rtcReleaseScene(RTCScene) {}
inline void rtcReleaseDevice(RTCDevice) {}
inline int rtcSYCLDeviceSelector(const void *) { return 1; }
template <class C>
RTCDevice rtcNewSYCLDevice(const C &, const char *) { return nullptr; }
