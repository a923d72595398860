// capi.cu — version and status text of the C ABI (include/cloudaae_b200.h).
#include "common.cuh"

extern "C" int caae_abi_version(void) { return CAAE_ABI_VERSION; }

extern "C" const char* caae_status_string(int status) {
  switch (status) {
    case CAAE_OK: return "ok";
    case CAAE_E_BADSHAPE: return "invalid size argument";
    case CAAE_E_NULLPTR: return "null pointer for a non-empty tensor";
    case CAAE_E_SCRATCH: return "scratch buffer required for this size but not provided";
    case CAAE_E_UNSUPPORTED: return "unsupported configuration";
    default: break;
  }
  if (status > 0) return cudaGetErrorString((cudaError_t)status);
  return "unknown cloudaae_b200 status";
}
