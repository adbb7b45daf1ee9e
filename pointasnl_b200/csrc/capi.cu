// capi.cu -- library identification and error text for the C ABI (include/pasnl.h).
#include "common.cuh"

extern "C" const char *pasnl_version(void) { return "pasnl-b200 0.1.0 sm_100a"; }

extern "C" const char *pasnl_error_string(int code) {
    if (code == PASNL_OK) return "ok";
    if (code == PASNL_EINVAL) return "invalid argument (shape / pointer / range check failed)";
    if (code > 0) return cudaGetErrorString((cudaError_t)code);
    return "unknown pasnl error";
}
