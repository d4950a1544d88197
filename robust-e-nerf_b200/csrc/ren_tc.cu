// ren_tc.cu -- tcgen05 engine (placeholder until the UMMA kernels land).
#include "ren_common.cuh"
int ren_tc_supported(const RenPlan& p) { (void)p; return 0; }
size_t ren_tc_workspace_bytes(const RenPlan& p, long long S) { (void)p; (void)S; return 0; }
int ren_tc_launch_fwd(const RenPlan& p, long long n_rays, int K, const float* o, const float* d,
                      const float* jitter, const float* params, float* stash, float* raw, void* ws,
                      size_t ws_bytes, cudaStream_t st) {
  ren_set_error("tcgen05 engine not built");
  return REN_E_UNSUPPORTED;
}
