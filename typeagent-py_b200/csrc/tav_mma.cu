// tav_mma.cu — placeholder until the tcgen05 kernel lands in this file.
#include "tav_internal.h"
namespace tav {
bool mma_supported(int, int) { return false; }
size_t mma_workspace_bytes(const MmaArgs&) { return 0; }
cudaError_t launch_mma_search(const MmaArgs&, void*, size_t, cudaStream_t, int*) {
    return cudaErrorNotSupported;
}
}  // namespace tav
