// Specialised flooding min-sum decoder for BG2, Zc = 96 (see nrldpc_decode_spec.cuh).
#include "nrldpc_decode_spec.cuh"

namespace nrldpc {
NRLDPC_SPEC_INSTANCE(2, 96)
}  // namespace nrldpc
