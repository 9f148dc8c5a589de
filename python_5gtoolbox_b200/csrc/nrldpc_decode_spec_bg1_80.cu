// Specialised flooding min-sum decoder for BG1, Zc = 80 (see nrldpc_decode_spec.cuh).
#include "nrldpc_decode_spec.cuh"

namespace nrldpc {
NRLDPC_SPEC_INSTANCE(1, 80)
}  // namespace nrldpc
