// placeholder until the packed kernel lands: never eligible, so the general kernel handles everything
#include "common.cuh"
namespace tb {
cudaError_t launch_gotoh_packed(bool, const GotohBatch&, int, cudaStream_t) { return cudaErrorNotSupported; }
cudaError_t gotoh_packed_blocks_per_sm(bool, int* out) { *out = 0; return cudaSuccess; }
int gotoh_packed_warps_per_block() { return 1; }
bool gotoh_packed_eligible(int, int, int, int, int, int) { return false; }
unsigned long long gotoh_packed_ptr_words(int, int) { return 0; }
}
