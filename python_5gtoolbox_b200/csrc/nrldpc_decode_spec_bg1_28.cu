// Specialised flooding min-sum decoder for BG1, Zc = 28 (see nrldpc_decode_spec.cuh).
#include "nrldpc_decode_spec.cuh"

namespace nrldpc {
NRLDPC_SPEC_INSTANCE(1, 28)
}  // namespace nrldpc
