// C-ABI plumbing shared by every entry point: thread-local error string, device queries, debug knobs.
#include "common.cuh"

namespace mpx {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int num_sms() {
  static int cached = 0;
  if (cached == 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
      cached = n;
    else
      return 148;
  }
  return cached;
}

unsigned long long* g_dbg_stamps = nullptr;
int g_use_pdl = 12;   // decode attention + fused GLU (trip 81: +1.5 % on the whole step; the encoder / policy-head bits lose)
extern int g_debug_swap_lbo_sbo;
extern int g_attn_decode_impl;
extern int g_attn_train_impl;
extern int g_glu_fused_cluster;
extern int g_value_fused_cluster;
extern int g_muon_ns_impl;

}  // namespace mpx

extern "C" const char* mpx_last_error(void) { return mpx::g_err; }
extern "C" int mpx_version(void) { return 100; }
extern "C" int mpx_debug_set_stamps(void* device_buffer) {
  mpx::g_dbg_stamps = reinterpret_cast<unsigned long long*>(device_buffer);
  return MPX_OK;
}
extern "C" int mpx_debug_set(int key, int value) {
  if (key == 0) {
    mpx::g_debug_swap_lbo_sbo = value;
    return MPX_OK;
  }
  if (key == 1) {
    mpx::g_attn_decode_impl = value;
    return MPX_OK;
  }
  if (key == 2) {
    mpx::g_attn_train_impl = value;
    return MPX_OK;
  }
  if (key == 3) {
    mpx::g_glu_fused_cluster = value;
    return MPX_OK;
  }
  if (key == 4) {
    mpx::g_value_fused_cluster = value;
    return MPX_OK;
  }
  if (key == 5) {
    mpx::g_muon_ns_impl = value;
    return MPX_OK;
  }
  if (key == 6) {
    mpx::g_use_pdl = value;
    return MPX_OK;
  }
  mpx::set_error("mpx_debug_set: unknown key %d", key);
  return MPX_ERR_INVALID;
}
