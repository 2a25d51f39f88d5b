// Version / error strings of the wsdl_b200 C ABI (include/wsdl_b200.h).
#include "common.cuh"

extern "C" int wsdl_version(void) { return WSDL_VERSION; }

extern "C" const char* wsdl_strerror(int rc) {
  switch (rc) {
    case 0: return "ok";
    case WSDL_E_NULL: return "wsdl: required pointer is NULL";
    case WSDL_E_SHAPE: return "wsdl: invalid shape / extent";
    case WSDL_E_ALIGN: return "wsdl: pointer not sufficiently aligned";
    case WSDL_E_WORKSPACE: return "wsdl: workspace too small";
    case WSDL_E_DTYPE: return "wsdl: unknown dtype code";
    case WSDL_E_ARG: return "wsdl: invalid argument";
    default: break;
  }
  if (rc > 0) return cudaGetErrorString((cudaError_t)rc);
  return "wsdl: unknown error";
}
