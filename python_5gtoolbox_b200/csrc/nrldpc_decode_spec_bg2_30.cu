// Specialised flooding min-sum decoder for BG2, Zc = 30 (see nrldpc_decode_spec.cuh).
#include "nrldpc_decode_spec.cuh"

namespace nrldpc {
NRLDPC_SPEC_INSTANCE(2, 30)
}  // namespace nrldpc
