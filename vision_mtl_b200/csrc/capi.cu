// Library-level entry points of libvmtl_b200.so.
#include "vmtl_common.cuh"

extern "C" int vmtl_version(void) { return 100; /* 0.1.0 */ }

extern "C" const char* vmtl_strerror(int code) {
  switch (code) {
    case VMTL_OK: return "ok";
    case VMTL_EINVAL: return "invalid argument";
    case VMTL_EALIGN: return "pointer or channel count is not 16-byte friendly";
    case VMTL_ECUDA: return "CUDA runtime / launch failure";
    case VMTL_EUNSUPPORTED: return "shape not covered by the sm_100a kernels";
    case VMTL_EWORKSPACE: return "workspace too small";
    default: return "unknown error";
  }
}

extern "C" int vmtl_sm_count(void) { return vmtl::sm_count(); }
