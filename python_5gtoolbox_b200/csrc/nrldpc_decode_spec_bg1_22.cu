// Specialised flooding min-sum decoder for BG1, Zc = 22 (see nrldpc_decode_spec.cuh).
#include "nrldpc_decode_spec.cuh"

namespace nrldpc {
NRLDPC_SPEC_INSTANCE(1, 22)
}  // namespace nrldpc
