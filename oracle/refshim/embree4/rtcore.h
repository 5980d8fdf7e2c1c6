/*
 * embree4/rtcore.h — TEST-ONLY shim of the Embree 4 declarations the reference's hot path uses
 * (src/trace_ray.hpp:18-30, src/camera.hpp:47-62, src/app.hpp:46-57, src/scene.hpp). Embree itself is a
 * third-party dependency that is not available here; rtcIntersect1 is forwarded to the oracle's
 * brute-force closest-hit search (refshim.cpp), everything else is a stub.
 */
#pragma once
#include <cstddef>
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
/* scene construction (src/scene.cpp): recorded by oracle/refshim/scenref.cpp, unused elsewhere */
enum RTCGeometryType { RTC_GEOMETRY_TYPE_TRIANGLE = 0, RTC_GEOMETRY_TYPE_INSTANCE = 121 };
enum RTCBufferType { RTC_BUFFER_TYPE_INDEX = 0, RTC_BUFFER_TYPE_VERTEX = 1 };
enum RTCFormat { RTC_FORMAT_UINT3 = 0x5003, RTC_FORMAT_FLOAT3 = 0x9003, RTC_FORMAT_FLOAT4X4_COLUMN_MAJOR = 0x9244 };
RTCScene rtcNewScene(RTCDevice device);
RTCGeometry rtcNewGeometry(RTCDevice device, RTCGeometryType type);
void rtcSetSharedGeometryBuffer(RTCGeometry geometry, RTCBufferType type, unsigned int slot, RTCFormat format, const void *ptr,
                                size_t byteOffset, size_t byteStride, size_t itemCount);
void rtcCommitGeometry(RTCGeometry geometry);
unsigned int rtcAttachGeometry(RTCScene scene, RTCGeometry geometry);
void rtcCommitScene(RTCScene scene);
void rtcSetGeometryTimeStepCount(RTCGeometry geometry, unsigned int timeStepCount);
void rtcSetGeometryInstancedScene(RTCGeometry geometry, RTCScene scene);
void rtcSetGeometryTransform(RTCGeometry geometry, unsigned int timeStep, RTCFormat format, const void *xfm);
void rtcSetGeometryUserData(RTCGeometry geometry, void *ptr);
inline void rtcReleaseGeometry(RTCGeometry) {}
inline void rtcReleaseScene(RTCScene) {}
inline void rtcReleaseDevice(RTCDevice) {}
inline int rtcSYCLDeviceSelector(const void *) { return 1; }
template <class C>
RTCDevice rtcNewSYCLDevice(const C &, const char *) { return nullptr; }
