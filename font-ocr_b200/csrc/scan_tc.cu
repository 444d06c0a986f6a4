// scan_tc.cu -- placeholder until the tcgen05 kernel lands (next commit): reports "unsupported"
// so FOCR_KERNEL_AUTO takes the SIMT kernel.
#include "scan_tc.cuh"
namespace focr {
int tc_class_build(TcClass &tc, const uint8_t *, uint32_t n_w, uint32_t n_h, uint32_t np, uint32_t n_tpl,
                   const uint32_t *, const TplInfo *)
{
    tc.supported = false; tc.n_w = n_w; tc.n_h = n_h; tc.np = np; tc.n_tpl = n_tpl;
    return 0;
}
void tc_class_release(TcClass &) {}
bool tc_class_supported(const TcClass &tc) { return tc.supported; }
void tc_workspace_release(TcWorkspace &) {}
cudaError_t launch_scan_tc(TcWorkspace &, const TcClass &, const ScanArgs &, int, int, cudaStream_t, int *)
{
    return cudaErrorNotSupported;
}
}  // namespace focr
