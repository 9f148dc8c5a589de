"""Mirror of py5gphy/ldpc/nr_ldpc_raterecover.py on the CUDA rate-recovery kernel (csrc/nrldpc_ratematch.cu)."""
import numpy as np

from .. import engine


def raterecover_ldpc(LLr_fe, Ncb, N, k0, Qm, Zc, K_apo, K):
    """LLr_dn = raterecover_ldpc(LLr_fe, Ncb, N, k0, Qm, Zc, K_apo, K) -- de-interleaving and de-selection
    of one codeblock, py5gphy/ldpc/nr_ldpc_raterecover.py:6-65: repeated positions are averaged, positions
    that were not transmitted are 0, the fillers [K_apo-2Zc, K-2Zc) get 10*max|LLr_fe|.  float64 [N]."""
    x = np.asarray(LLr_fe)
    assert x.size % Qm == 0   # the reference's reshape(E // Qm, Qm) raises otherwise
    x = x.astype(np.float64 if x.dtype != np.float32 else np.float32, copy=False).reshape(-1)
    return engine.raterecover_batch(x, [x.size], Ncb, N, k0, Qm, Zc, K_apo, K, out_f64=True)[0]
