// tscm_internal.h — the few host helpers shared by the translation units of libtscm_b200.so
// (defined in tscm_b200.cu; not part of the C-ABI, hidden from the dynamic symbol table).
#pragma once

namespace tscm {
namespace internal {

// printf-style text behind tscm_last_error() of the calling thread
__attribute__((visibility("hidden"))) void set_error(const char* fmt, ...);

// cudaSetDevice(device) (-1 = keep the current one) after checking that a CUDA device exists and
// is sm_100 — there is no CPU fallback anywhere in the library.  `what` names the caller in the
// error text.  Returns a TSCM_* code.
__attribute__((visibility("hidden"))) int select_device(int device, const char* what, int* sm_count);

// restores the caller's current device on scope exit
struct DeviceScope {
  int prev = -1;
  DeviceScope();
  ~DeviceScope();
};

}  // namespace internal
}  // namespace tscm
