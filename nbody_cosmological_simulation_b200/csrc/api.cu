// Library-level entry points of libnbody_b200: version, errors, device probe, layout helpers.
#include "common.cuh"

using namespace nb;

extern "C" int nb_abi_version(void) { return NB_ABI_VERSION; }

extern "C" const char* nb_error_string(int status) {
    switch (status) {
        case NB_OK: return "ok";
        case NB_ERR_INVALID_ARGUMENT: return "invalid argument (dim/dtype/mode/size/null pointer/aliasing)";
        case NB_ERR_UNSUPPORTED: return "combination not supported by libnbody_b200 (no fallback exists)";
        case NB_ERR_WORKSPACE_TOO_SMALL: return "workspace too small";
        case NB_ERR_NO_DEVICE: return "no usable sm_100 CUDA device";
        default: break;
    }
    if (status >= NB_ERR_CUDA_BASE) return cudaGetErrorString((cudaError_t)(status - NB_ERR_CUDA_BASE));
    return "unknown libnbody_b200 status";
}

extern "C" int nb_device_info(int* sm_count, int* cc_major, int* cc_minor) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return NB_ERR_NO_DEVICE; }
    int sms = 0, maj = 0, min = 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&maj, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&min, cudaDevAttrComputeCapabilityMinor, dev) != cudaSuccess) {
        cudaGetLastError();
        return NB_ERR_NO_DEVICE;
    }
    if (sm_count) *sm_count = sms;
    if (cc_major) *cc_major = maj;
    if (cc_minor) *cc_minor = min;
    return maj == 10 ? NB_OK : NB_ERR_NO_DEVICE;       // the only image in this library is sm_100a
}

extern "C" int64_t nb_key_from_double(double v) { return key_from_double(v); }
extern "C" double nb_double_from_key(int64_t key) { return double_from_key(key); }

extern "C" int64_t nb_chunk_sources(int dtype) { return dtype == NB_F32 || dtype == NB_F64 ? chunk_sources(dtype) : 0; }
extern "C" int64_t nb_chunk_bytes(int dim, int dtype) {
    (void)dtype;                                         // a unit is 16 B + 16|8 B in both dtypes
    return (dim == 2 || dim == 3) ? chunk_bytes(dim) : 0;
}
extern "C" int64_t nb_num_chunks(int64_t n, int dtype) {
    const int64_t cs = nb_chunk_sources(dtype);
    return (cs == 0 || n <= 0) ? 0 : (n + cs - 1) / cs;
}
extern "C" int64_t nb_packed_bytes(int64_t n, int dim, int dtype) { return nb_num_chunks(n, dtype) * nb_chunk_bytes(dim, dtype); }

// ---- instrumentation events (used with nb_profile_next_force) -------------------------------------------------
extern "C" int nb_event_create(void** event_out) {
    if (!event_out) return NB_ERR_INVALID_ARGUMENT;
    cudaEvent_t e = nullptr;
    const cudaError_t rc = cudaEventCreate(&e);
    *event_out = (void*)e;
    return cuda_status(rc);
}
extern "C" int nb_event_destroy(void* event) { return event ? cuda_status(cudaEventDestroy((cudaEvent_t)event)) : NB_OK; }
extern "C" int nb_event_elapsed_ms(void* start_event, void* stop_event, float* ms_out) {
    if (!start_event || !stop_event || !ms_out) return NB_ERR_INVALID_ARGUMENT;
    cudaError_t rc = cudaEventSynchronize((cudaEvent_t)stop_event);
    if (rc == cudaSuccess) rc = cudaEventElapsedTime(ms_out, (cudaEvent_t)start_event, (cudaEvent_t)stop_event);
    return cuda_status(rc);
}
