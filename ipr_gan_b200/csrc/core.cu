// core.cu -- library bookkeeping: version, error strings, launch counter.
#include "ipr_common.cuh"

unsigned long long g_ipr_launches = 0ULL;

extern "C" int ipr_version(void) { return 100; }   // 0.1.0

extern "C" uint64_t ipr_launch_count(void) {
    return (uint64_t)__atomic_load_n(&g_ipr_launches, __ATOMIC_RELAXED);
}

extern "C" const char *ipr_strerror(int code) {
    switch (code) {
        case IPR_OK:            return "ok";
        case IPR_E_NULL:        return "ipr: required pointer is NULL";
        case IPR_E_SHAPE:       return "ipr: inconsistent or non-positive dimensions";
        case IPR_E_UNSUPPORTED: return "ipr: shape outside what the sm_100a kernels cover";
        case IPR_E_ALIGN:       return "ipr: pointer alignment requirement violated";
        case IPR_E_WORKSPACE:   return "ipr: workspace too small";
        default: break;
    }
    if (code > 0) return cudaGetErrorString((cudaError_t)code);
    return "ipr: unknown error code";
}
