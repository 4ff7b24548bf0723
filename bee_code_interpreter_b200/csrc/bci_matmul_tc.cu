// bci_matmul_tc.cu — tcgen05 / TMA path of K3 (see bci_matmul.cu).  Placeholder until the
// 3xTF32 kernel lands: reports "unsupported" so every shape takes the SIMT fp32 kernel.
#include "bci_common.cuh"

namespace bci {

int matmul_tc_supported(int, int, int, const void* const*, const void* const*, void* const*, int) { return 0; }

int matmul_tc_launch(int, const void* const*, const void* const*, void* const*, int, int, int) {
  set_error("tcgen05 matmul path is not built");
  return BCI_ERR_UNSUPPORTED;
}

}  // namespace bci
