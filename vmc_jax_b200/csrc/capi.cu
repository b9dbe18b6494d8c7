// Library-level entry points: version, error strings, last CUDA error.
#include "common.cuh"

static thread_local cudaError_t g_last_cuda = cudaSuccess;
void jvmc_set_last_cuda_error(cudaError_t e) { g_last_cuda = e; }

extern "C" int jvmc_version(void) { return 100; }

extern "C" const char* jvmc_error_string(int code) {
  switch (code) {
    case JVMC_OK: return "ok";
    case JVMC_ERR_ARG: return "invalid argument";
    case JVMC_ERR_CUDA: return cudaGetErrorString(g_last_cuda);
    case JVMC_ERR_UNSUPPORTED: return "unsupported size or configuration";
    case JVMC_ERR_SOLVER: return "cuSOLVER failure";
    default: return "unknown error";
  }
}
