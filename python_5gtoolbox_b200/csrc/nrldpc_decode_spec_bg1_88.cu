// Specialised flooding min-sum decoder for BG1, Zc = 88 (see nrldpc_decode_spec.cuh).
#include "nrldpc_decode_spec.cuh"

namespace nrldpc {
NRLDPC_SPEC_INSTANCE(1, 88)
}  // namespace nrldpc
