"""Batched Monte-Carlo driver: the drop-in for scripts/internal/sim_ldpc_internal.run_ldpc_simulation.

The reference loop decodes ONE codeblock per pass (scripts/internal/sim_ldpc_internal.py:49-77) and only
looks at its stopping rule when test_count is 1000, 2000, 4000 or 10000 (:67-77).  Here every stretch
between two such checkpoints is generated, encoded and decoded as one batch on the GPU, so the rule,
the BLER values and the pickle layout ([sim_config, test_config_list, test_results_list], :87-91) are
those of the reference.

Sharding: codeblocks and grid points are independent.  With torch.distributed initialised the
work is split by grid point (default: one (decoder setting, SNR) point per rank, one all-reduce at the end) or by
codeblock inside every point (an all-reduce(SUM) of one int64 per checkpoint); see run_ldpc_simulation.

Input generation
  rng="numpy"  (default) the reference's own draws from NumPy's global RNG, in its order (randint
               :247 then normal :253 per codeblock), so np.random.seed(s) reproduces the reference's
               inputs bit for bit; every rank draws the full stream and keeps its share.
  rng="device" Philox bits + noise generated on the GPU (nrldpc_random_bits / nrldpc_awgn_llr) for
               10^6-codeblock points; statistically equivalent, not seed-compatible.
"""
import pickle
import time

import numpy as np

CHECKPOINTS = (1000, 2000, 4000, 10000)   # np.array([200,400,800,2000])*5, sim_ldpc_internal.py:67
FAIL_LIMITS = (50, 25, 10)                # np.array([10,5,2])*5, :68


def test_plan(algo_list, alpha_list, beta_list, mixed_list, L_list):
    """[(flag, algo, L, alpha, beta)] in the reference's nesting order (sim_ldpc_internal.py:15-40)."""
    plan = []
    for algo in algo_list:
        if algo in ['BP', 'BF', 'min-sum']:
            params = [(1, 0)]
        elif algo == 'NMS':
            params = [(a, 0) for a in alpha_list]
        elif algo == 'OMS':
            params = [(1, b) for b in beta_list]
        else:
            params = [(m[0], m[1]) for m in mixed_list]
        for L in L_list:
            for (a, b) in params:
                if algo in ['BF', 'BP', 'min-sum']:
                    flag = '{} L={}'.format(algo, L)
                elif algo == 'NMS':
                    flag = 'NMS-alpha={}-L={}'.format(a, L)
                elif algo == 'OMS':
                    flag = 'OMS-beta={}-L={}'.format(b, L)
                else:
                    flag = 'mixed-MS-[alpha,beta]=[{},{}]-L={}'.format(a, b, L)
                plan.append((flag, algo, L, a, b))
    return plan


def stop_now(test_count, failed_count):
    """The reference's termination check (sim_ldpc_internal.py:69-77)."""
    for cp, lim in zip(CHECKPOINTS[:3], FAIL_LIMITS):
        if test_count == cp and failed_count >= lim:
            return True
    return test_count == CHECKPOINTS[3]


class _Dist:
    def __init__(self):
        self.rank, self.world, self.dist = 0, 1, None
        try:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized():
                self.rank, self.world, self.dist = dist.get_rank(), dist.get_world_size(), dist
        except ImportError:
            pass

    def sum(self, values, device=None):
        if self.world == 1:
            return [int(v) for v in values]
        import torch
        t = torch.tensor([int(v) for v in values], dtype=torch.int64, device=device)
        self.dist.all_reduce(t)
        return t.tolist()


class CudaBackend:
    """generate / encode / decode one batch share on the GPU through the C ABI."""

    def __init__(self, device=None):
        import torch
        self.torch = torch
        self.device = torch.device(device if device is not None else "cuda")
        self.reduce_device = self.device

    def numpy_batch(self, Zc, bgn, snr_db, crcpoly, n, take):
        """n codeblocks drawn exactly like for_test_5g_ldpc_encoder (nr_ldpc_decode.py:247-257); only the
        ones in `take` are encoded.  Returns (blkandcrc int8[m,K], llr float64[m,N])."""
        from . import crc, engine
        K, N = ((22, 66) if bgn == 1 else (10, 50))
        K, N = K * Zc, N * Zc
        crc_len = 24 if crcpoly in ['24A', '24B'] else 16
        sigma = 10 ** (-snr_db / 20)
        bits = np.empty((len(take), K - crc_len), np.int8)
        noise = np.empty((len(take), N))
        want = {b: i for i, b in enumerate(take)}
        for b in range(n):  # the global stream must advance for every codeblock, kept or not
            inb = np.random.randint(2, size=K - crc_len)
            nz = np.random.normal(0, sigma, N)
            if b in want:
                bits[want[b]] = inb
                noise[want[b]] = nz
        if not len(take):
            return np.empty((0, K), np.int8), np.empty((0, N))
        blk = crc.nr_crc_encode_batch(bits, crcpoly)
        dn = engine.encode_batch(blk.copy(), bgn, Zc)
        llr = 2 * ((1 - 2 * dn) + noise) / 10 ** (-snr_db / 10)
        return blk, llr

    def count_failures(self, Zc, bgn, snr_db, crcpoly, algo, L, alpha, beta, n, take, rng, seed, offset):
        from . import engine, _lib
        import ctypes
        torch = self.torch
        K = (22 if bgn == 1 else 10) * Zc
        if not len(take) and rng != "numpy":
            return 0
        if rng == "numpy":
            blk, llr = self.numpy_batch(Zc, bgn, snr_db, crcpoly, n, take)
            if not len(take):
                return 0
            if algo == 'BF':
                ck, _, _ = engine.decode_bf_batch(llr, Zc, bgn, L)
            elif algo == 'BP':
                ck, _, _ = engine.decode_bp_batch(llr, Zc, bgn, L)
            else:
                ck = engine.decode_batch(llr.astype(np.float32), Zc, bgn, L, alpha, beta, True)["ck"]
            return int((ck[:, :K] != blk).any(axis=1).sum())
        # device generation: codeblock id = offset + b draws from its own Philox counter range
        m = len(take)
        crc_len = 24 if crcpoly in ['24A', '24B'] else 16
        A = K - crc_len
        poly = {"24A": 3, "24B": 4, "16": 2}[crcpoly]
        first, stride = offset + take[0], (take[1] - take[0]) if m > 1 else 1
        L_ = _lib.lib()
        counters = torch.zeros(4, dtype=torch.int64, device=self.device)   # accumulated on the device: ONE host read per stretch
        with torch.cuda.device(self.device):
            for i0 in range(0, m, 8192):
                mm = min(8192, m - i0)
                s = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
                bits = torch.empty((mm, A), dtype=torch.int8, device=self.device)
                _lib.check(L_.nrldpc_random_bits_rows(bits.data_ptr(), mm, A, seed, first + i0 * stride, stride, s), "random_bits")
                blk = torch.empty((mm, K), dtype=torch.int8, device=self.device)
                _lib.check(L_.nrldpc_crc_encode(bits.data_ptr(), mm, A, poly, blk.data_ptr(), s), "crc")
                dn = engine.encode_batch(blk, bgn, Zc, fix_fillers=False)   # no fillers in blk: nothing to fix, no copy needed
                llr = torch.empty(dn.shape, dtype=torch.float32, device=self.device)
                _lib.check(L_.nrldpc_awgn_llr_rows(dn.data_ptr(), mm, dn.shape[1], float(snr_db), seed, first + i0 * stride,
                                                   stride, llr.data_ptr(), s), "awgn")
                if algo == 'BF':   # quasi-cyclic bit-flipping kernel, device-resident like the min-sum chain
                    ck, _, it = engine.decode_bf_batch(llr, Zc, bgn, L)
                elif algo == 'BP':   # quasi-cyclic sum-product kernel: float64 arithmetic on the device-resident fp32 LLRs
                    ck, _, it = engine.decode_bp_batch(llr, Zc, bgn, L)
                else:
                    r = engine.decode_batch(llr, Zc, bgn, L, alpha, beta, True)
                    ck, it = r["ck"], r["iters"]
                engine.count_errors(blk, ck, K, it, counters)
        fails = int(counters[1].item())
        return fails


def _run_point(be, Zc, bgn, snr_db, crcpoly, algo, L, alpha, beta, rng, seed, point, rank, world, reduce):
    """One (decoder setting, SNR) point with the reference's stopping rule (sim_ldpc_internal.py:46-77); rank / world
    split every stretch between two checkpoints by codeblock (world = 1: the whole stretch)."""
    test_count = failed_count = 0
    while True:
        nxt = next(c for c in CHECKPOINTS if c > test_count)
        n = nxt - test_count
        take = list(range(rank, n, world))  # this rank's share of the stretch
        f = be.count_failures(Zc, bgn, snr_db, crcpoly, 'min-sum' if algo in ('NMS', 'OMS', 'mixed-MS') else algo,
                              L, alpha, beta, n, take, rng, seed, point * CHECKPOINTS[-1] + test_count)
        failed_count += reduce(f)
        test_count = nxt
        if stop_now(test_count, failed_count):
            return test_count, failed_count


def run_ldpc_simulation(Zc, bgn, crcpoly, algo_list, alpha_list, beta_list, mixed_list, L_list, snr_db_list, filename,
                        *, rng=None, seed=0x5601, shard=None, backend=None, verbose=True):
    """Same call and pickle output as scripts/internal/sim_ldpc_internal.run_ldpc_simulation (:9-91).

    rng    "numpy" (default; NRLDPC_SIM_RNG overrides) or "device", see the module docstring.
    shard  how the work is split over the ranks of an initialised torch.distributed job:
           "codeblock"  every rank walks every point and takes every W-th codeblock of a stretch; with rng="numpy" every
                        rank draws the reference's whole global stream, so the inputs are those of the seeded reference
                        bit for bit on any number of ranks -- and the host RNG, not the GPU, sets the pace;
           "point"      (default when W > 1) one (decoder setting, SNR) point per rank, round robin: BASELINE config #3's
                        "sharded by SNR point".  Every point draws from its own stream, seeded from (base, point index)
                        with base = one draw from the global RNG on rank 0, so the result does not depend on W; the
                        stopping rule of a point is the reference's.  One all-reduce of int64[points x 2] at the end.
    """
    import os
    d = _Dist()
    be = backend if backend is not None else CudaBackend()
    rng = rng or os.environ.get("NRLDPC_SIM_RNG", "numpy")
    shard = shard or ("point" if d.world > 1 else "codeblock")
    assert rng in ("numpy", "device") and shard in ("codeblock", "point")
    plan = test_plan(algo_list, alpha_list, beta_list, mixed_list, L_list)
    points = [(flag, algo, L, alpha, beta, snr_db) for flag, algo, L, alpha, beta in plan for snr_db in snr_db_list]
    results = []
    if shard == "codeblock":
        for p, (flag, algo, L, alpha, beta, snr_db) in enumerate(points):
            start = time.time()
            tc, fc = _run_point(be, Zc, bgn, snr_db, crcpoly, algo, L, alpha, beta, rng, seed, p, d.rank, d.world,
                                lambda f: d.sum([f], getattr(be, "reduce_device", None))[0])
            results.append((tc, fc, time.time() - start))
    else:
        base = int(np.random.randint(0, 2 ** 31 - 1)) if d.rank == 0 else 0
        base = d.sum([base], getattr(be, "reduce_device", None))[0]
        mine = []
        for p, (flag, algo, L, alpha, beta, snr_db) in enumerate(points):
            if p % d.world != d.rank:
                mine += [0, 0, 0]
                continue
            start = time.time()
            if rng == "numpy":
                np.random.seed(np.random.SeedSequence([base, p]).generate_state(4))
            tc, fc = _run_point(be, Zc, bgn, snr_db, crcpoly, algo, L, alpha, beta, rng, seed + base, p, 0, 1, lambda f: f)
            mine += [tc, fc, int(1e3 * (time.time() - start))]
        tot = d.sum(mine, getattr(be, "reduce_device", None))
        results = [(tot[3 * p], tot[3 * p + 1], tot[3 * p + 2] / 1e3) for p in range(len(points))]
    test_results_list, test_config_list = [], []
    for i, (flag, algo, L, alpha, beta) in enumerate(plan):
        test_config_list.append(flag)
        bler_result = []
        for j, snr_db in enumerate(snr_db_list):
            tc, fc, sec = results[i * len(snr_db_list) + j]
            bler_result.append(fc / tc)
            if verbose and d.rank == 0:
                print("finish test {}, Zc {}, bgn{},snr_db={}, test_count={},failed_count={},bler={:2.5f},elpased time: {:6.2f}".
                      format(flag, Zc, bgn, snr_db, tc, fc, fc / tc, sec))
        test_results_list.append(bler_result)
    sim_config = {'Zc': Zc, 'bgn': bgn}
    if d.rank == 0 and filename:
        with open(filename, 'wb') as handle:
            pickle.dump([sim_config, test_config_list, test_results_list], handle, protocol=pickle.HIGHEST_PROTOCOL)
    return sim_config, test_config_list, test_results_list


# ------------------------------------------------------------------ fixed-count BLER / BER curves (BASELINE config #5)

def shard_range(n, rank, world):
    """Contiguous share [lo, hi) of n codeblocks for `rank` of `world` (sizes differ by at most one)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def _device_point_counters(device, Zc, bgn, snr_db, crcpoly, L, alpha, beta, lo, hi, seed, early_term=True, chunk=16384,
                           packed=None):
    """{codeblocks, block errors, bit errors, iterations} of codeblocks lo..hi-1 of one SNR point, generated
    (Philox, codeblock id = its index), encoded, decoded and counted on the device.

    packed (default: whenever Zc is a multiple of 32): bits, CRC, codeword and decisions stay bit-packed between the
    kernels (SURVEY 8(d)'s K/8 + N/8 bytes per codeblock instead of a byte per bit); same Philox counters, so the
    codeblocks, the noise and therefore the counters are those of the byte-per-bit chain."""
    import ctypes
    import torch
    from . import engine, _lib
    K = (22 if bgn == 1 else 10) * Zc
    crc_len = 24 if crcpoly in ['24A', '24B'] else 16
    A = K - crc_len
    poly = {"24A": 3, "24B": 4, "16": 2}[crcpoly]
    counters = torch.zeros(4, dtype=torch.int64, device=device)
    L_ = _lib.lib()
    if packed is None:
        packed = Zc % 32 == 0
    assert not packed or Zc % 32 == 0
    with torch.cuda.device(device):
        for i0 in range(lo, hi, chunk):
            mm = min(chunk, hi - i0)
            if packed:
                blk = engine.random_bits_packed(mm, A, seed, device, first_id=i0, row_words=K // 32)
                engine.crc_attach_packed(blk, A, crcpoly)
                dn = engine.encode_packed(blk, bgn, Zc)
                llr = engine.awgn_llr_packed(dn, dn.shape[1] * 32, snr_db, seed, first_id=i0)
                r = engine.decode_batch(llr, Zc, bgn, L, alpha, beta, early_term, want_ck=False, want_info=True)
                engine.count_errors_packed(blk, r["info"], K, r["iters"], counters)
                continue
            s = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
            bits = torch.empty((mm, A), dtype=torch.int8, device=device)
            _lib.check(L_.nrldpc_random_bits_rows(bits.data_ptr(), mm, A, seed, i0, 1, s), "random_bits")
            blk = torch.empty((mm, K), dtype=torch.int8, device=device)
            _lib.check(L_.nrldpc_crc_encode(bits.data_ptr(), mm, A, poly, blk.data_ptr(), s), "crc")
            dn = engine.encode_batch(blk, bgn, Zc, fix_fillers=False)   # no fillers in blk: nothing to fix, no copy needed
            llr = torch.empty(dn.shape, dtype=torch.float32, device=device)
            _lib.check(L_.nrldpc_awgn_llr_rows(dn.data_ptr(), mm, dn.shape[1], float(snr_db), seed, i0, 1, llr.data_ptr(), s), "awgn")
            r = engine.decode_batch(llr, Zc, bgn, L, alpha, beta, early_term)
            engine.count_errors(blk, r["ck"], K, r["iters"], counters)
    return counters


def bler_curve(Zc, bgn, snr_db_list, n_per_point, L, alpha=1.0, beta=0.0, crcpoly='24A', *, seed=0x5601, device=None,
               point_counters=None):
    """BLER / BER / mean iterations of the min-sum decoder at every SNR of `snr_db_list` with a FIXED number of
    codeblocks per point (10^6 in BASELINE config #5), device-generated inputs.

    Sharding: rank k of W takes a contiguous share of every point's codeblocks (a codeblock's
    bits and noise depend only on (seed, point, index), so any W sees the same codeblocks); the ONLY collective
    is one all-reduce(SUM) of int64[n_points x 4] = {codeblocks, block errors, bit errors, iterations} at the
    end.  Returns a list of dicts, one per SNR point.  `point_counters` replaces the device worker (tests)."""
    d = _Dist()
    if point_counters is None:
        import torch
        dev = torch.device(device if device is not None else "cuda")
        def point_counters(p, snr_db, lo, hi):
            return _device_point_counters(dev, Zc, bgn, snr_db, crcpoly, L, alpha, beta, lo, hi, seed + 7919 * p).tolist()
        reduce_device = dev
    else:
        reduce_device = None
    rows = []
    for p, snr_db in enumerate(snr_db_list):
        lo, hi = shard_range(n_per_point, d.rank, d.world)
        rows.extend(int(v) for v in (point_counters(p, snr_db, lo, hi) if hi > lo else [0, 0, 0, 0]))
    total = d.sum(rows, reduce_device)
    K = (22 if bgn == 1 else 10) * Zc
    out = []
    for p, snr_db in enumerate(snr_db_list):
        n, be, bit, it = total[4 * p:4 * p + 4]
        out.append({"snr_db": snr_db, "codeblocks": n, "block_errors": be, "bit_errors": bit, "bler": be / max(n, 1),
                    "ber": bit / max(n * K, 1), "mean_iters": it / max(n, 1)})
    return out
