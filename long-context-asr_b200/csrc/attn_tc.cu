// placeholder until the tcgen05 flash-attention kernel lands (next commit)
#include "common.cuh"
namespace lcasr {
int attn_tc_available() { return 0; }
int attn_tc_launch(const void*, const void*, const void*, int, int64_t, int, int, int, int64_t, void*, cudaStream_t) {
  return set_error(LCASR_E_UNSUPPORTED, "attention: tcgen05 kernel not built");
}
}  // namespace lcasr
