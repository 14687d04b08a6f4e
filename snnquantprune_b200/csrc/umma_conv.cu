// placeholder until the tcgen05 kernel lands
#include "common.cuh"
namespace snnqp {
bool umma_conv3x3_supported(const snnqp_block_params &p, const float *att) { return false; }
int launch_conv3x3_umma(const snnqp_block_params &p, const uint8_t *x, const int8_t *wq,
                        const float *scale, const float *bias, uint8_t *spikes, float *u_final,
                        int32_t *acc_dump, cudaStream_t st) {
  return unsupported("tcgen05 conv not built");
}
}
