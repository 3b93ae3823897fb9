#include "common.cuh"
extern "C" int llmseg_attention(const llmseg_attn_params* p, void* stream) {
  (void)p; (void)stream;
  return llmseg::set_error(LLMSEG_ESHAPE, "llmseg_attention: not implemented yet");
}
