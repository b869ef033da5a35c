/* TEST INFRASTRUCTURE (oracle/): stand-in for the OptiX 7 device API used by the reference's envsampling/kernel.cu, so that
 * the program can be compiled and run on the host.  Only the ray-tracing INTRINSICS are replaced; every line of sampling,
 * MIS, BSDF and gradient code that runs is the reference's own.  optixTrace() answers from a brute-force any-hit test over
 * the triangle list registered by the driver (tmin = 0, tmax = 1e16, terminate on first hit; the miss program would set
 * payload 0 to 1). */
#pragma once
typedef unsigned long long OptixTraversableHandle;
typedef unsigned int OptixVisibilityMask;
enum { OPTIX_RAY_FLAG_DISABLE_ANYHIT = 1, OPTIX_RAY_FLAG_DISABLE_CLOSESTHIT = 2, OPTIX_RAY_FLAG_TERMINATE_ON_FIRST_HIT = 4 };

extern thread_local uint3 g_ref_launch_index;
extern uint3 g_ref_launch_dims;
bool ref_any_hit(float3 origin, float3 dir, float tmin, float tmax);

static inline uint3 optixGetLaunchIndex() { return g_ref_launch_index; }
static inline uint3 optixGetLaunchDimensions() { return g_ref_launch_dims; }
static inline void optixSetPayload_0(unsigned int) {}
static inline void optixTrace(OptixTraversableHandle, float3 origin, float3 dir, float tmin, float tmax, float, OptixVisibilityMask,
                              unsigned int, unsigned int, unsigned int, unsigned int, unsigned int& payload0) {
  payload0 = ref_any_hit(origin, dir, tmin, tmax) ? 0u : 1u;
}
