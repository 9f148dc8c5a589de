#!/usr/bin/env python3
"""bench.py -- decoded-info Gbit/s of the batched NR-LDPC min-sum decoder (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" is one pass of the hot path (nr_decode_ldpc semantics, BG1 Zc=384 R=1/3, NMS alpha=0.8,
exactly 10 flooding iterations, early termination off) over one batch of synthetic codeblocks:
random bits -> CUDA encoder -> BPSK/AWGN LLRs (Philox, +1 dB), all generated on the device.
  value  : whole-job Gbit/s of decoded info bits (K = 8448 per codeblock) with the LLRs resident in HBM
  e2e    : the same metric through the host-buffer C-ABI call (nrldpc_decode_minsum_host): pinned
           host LLRs in, packed info bits + status + iteration counts out, copies inside the timing
  roofline / cpu_baseline / clocks : see DESIGN.md "Measurement"
`--impl reference` times the CPU restatement of the reference (oracle/, OpenMP over all host cores)
on a bounded sample of the same workload.  The reference itself is pure Python (26 s per codeblock).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BGN, ZC, MAX_ITER, ALPHA, BETA, SNR_DB = 1, 384, 10, 0.8, 0.0, 1.0
K_INFO, N_CODED = 22 * ZC, 66 * ZC
ALGO_BYTES_PER_CB = 4 * N_CODED + K_INFO // 8   # fp32 LLRs in + packed info bits out = 102432
METRIC = "decoded_info_gbit_per_s"
WORKLOAD = "BG1 Zc=384 R=1/3 NMS(alpha=0.8) min-sum, 10 flooding iterations, no early termination, BPSK/AWGN +1 dB"


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.stop = index, [], threading.Event()
        self.t = threading.Thread(target=self.run, daemon=True)

    def run(self):
        while not self.stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                self.rows.append([c.strip() for c in out.strip().split(",")])
            except Exception:
                pass
            self.stop.wait(0.1)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.t.join(timeout=6)

    def summary(self):
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 6 for n, v in zip(names, r[2:6]) if v.lower().startswith("active")})
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def cpu_reference_run(steps, warmup, sample_cbs=None, early_term=0):
    """Time the CPU restatement of the reference (oracle/) with every host thread on a bounded sample."""
    import numpy as np
    from oracle import oracle as O
    O.build()
    O.set_num_threads(len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1))
    threads = O.num_threads()
    n = sample_cbs or max(threads * 100, 64)   # ~13-15 ms per codeblock per core -> 1-2 s per step
    rng = np.random.default_rng(0x5601)
    ck = rng.integers(0, 2, (n, K_INFO)).astype("i1")
    dn = O.encode_batch(ck, BGN, ZC)
    sigma = 10 ** (-SNR_DB / 20)
    llr = (2 * ((1 - 2 * dn.astype("f8")) + rng.normal(0, sigma, dn.shape)) / sigma ** 2).astype("f4").astype("f8")
    for _ in range(warmup):
        O.decode_batch(llr[: max(threads, 1)], ZC, BGN, MAX_ITER, "min-sum", ALPHA, BETA, early_term, np.float64)
    t0 = time.perf_counter()
    for _ in range(steps):
        O.decode_batch(llr, ZC, BGN, MAX_ITER, "min-sum", ALPHA, BETA, early_term, np.float64)
    dt = time.perf_counter() - t0
    gbps = steps * n * K_INFO / dt / 1e9
    return gbps, threads, n, dt / steps


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # --steps / --warmup are honoured as long as the run stays inside its time budget (REF_BUDGET_S, default 150 s):
    # a step is a bounded sample of the workload (100 codeblocks per host thread, ~1.5 s), probed by the warm-up step
    budget = float(os.environ.get("REF_BUDGET_S", "150"))
    t0 = time.perf_counter()
    _, threads, n, probe = cpu_reference_run(1, 0)
    steps = max(1, min(args.steps, int((budget - (time.perf_counter() - t0)) / max(probe, 1e-3)) - min(args.warmup, 2)))
    warmup = max(0, min(args.warmup, 2))
    gbps, threads, n, sps = cpu_reference_run(steps, warmup)
    sample = (f"{n} codeblocks per step x {steps} steps of the same workload, float64, {threads} OpenMP threads "
              f"({'all' if steps == args.steps else 'time budget: fewer than'} the {args.steps} steps asked for)")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": gbps, "unit": "Gbit/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": sps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "codeblocks_per_step": n,
                   "note": "CPU port of the reference algorithm (oracle/, C + OpenMP, ~13-15 ms per codeblock per core); the "
                           "reference itself is pure Python at ~26 s per codeblock per core (BASELINE.md) and cannot "
                           "travel to this box"},
        "cpu_baseline": {"value": gbps, "unit": "Gbit/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": gbps, "unit": "Gbit/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from python_5gtoolbox_b200 import engine, _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # stdout carries the JSON line only: NCCL prints its version banner (NCCL_DEBUG=VERSION/INFO) to fd 1
        # when the communicator is created, so fd 1 points at stderr until the first collective is done
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    B = args.batch
    seed = 0x5601 + rank

    # ---- synthetic workload, generated on the device in slices (encoder + Philox AWGN)
    llr = torch.empty((B, N_CODED), dtype=torch.float32, device=dev)
    sent = torch.empty((B, K_INFO), dtype=torch.int8, device=dev)
    sl = 4096
    for b0 in range(0, B, sl):
        nb = min(sl, B - b0)
        ck = engine.random_bits(nb, K_INFO, seed=seed, device=dev, offset=b0 * (K_INFO // 128 + 1))
        dn = engine.encode_batch(ck, BGN, ZC)
        engine.awgn_llr(dn, SNR_DB, seed=seed, offset=b0 * (N_CODED // 4), out=llr[b0:b0 + nb])
        sent[b0:b0 + nb] = ck
    info = torch.empty((B, (K_INFO + 31) // 32), dtype=torch.int32, device=dev)
    status = torch.empty((B,), dtype=torch.uint8, device=dev)
    iters = torch.empty((B,), dtype=torch.int32, device=dev)
    L = _lib.lib()
    stream = torch.cuda.current_stream()

    def step():
        _lib.check(L.nrldpc_decode_minsum(llr.data_ptr(), B, BGN, ZC, MAX_ITER, ALPHA, BETA, 0, None, info.data_ptr(),
                                          status.data_ptr(), iters.data_ptr(), stream.cuda_stream), "decode")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        ev0.record(stream)
        for _ in range(args.steps):
            step()
        ev1.record(stream)
        barrier()
    ms = torch.tensor([ev0.elapsed_time(ev1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    cbs_per_s = world * B * args.steps / (ms * 1e-3)
    value = cbs_per_s * K_INFO / 1e9

    # ---- secondary runs of SURVEY 8(d) (outside the timed region, reported in `config`): the reference's early
    # termination at +1 dB / -3 dB, and the OMS / mixed-MS parameter sets, same batch, one launch each
    extra = []
    if rank == 0 and not args.no_extra:
        bits_b = B * K_INFO
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        llr_low = None
        for name, snr, et, al, be in [("early_term +1 dB NMS(0.8)", SNR_DB, 1, ALPHA, 0.0), ("early_term -3 dB NMS(0.8)", -3.0, 1, ALPHA, 0.0),
                                      ("fixed 10 it OMS(beta=0.5)", SNR_DB, 0, 1.0, 0.5), ("fixed 10 it mixed(0.8,0.3)", SNR_DB, 0, 0.8, 0.3)]:
            x = llr
            if snr != SNR_DB:
                if llr_low is None:   # same codewords, other noise level: LLR = 2y/sigma^2 rescaled from stored bits
                    llr_low = torch.empty_like(llr)
                    for b0 in range(0, B, sl):
                        nb = min(sl, B - b0)
                        dn = engine.encode_batch(sent[b0:b0 + nb].clone(), BGN, ZC)
                        engine.awgn_llr(dn, snr, seed=seed + 7, offset=b0 * (N_CODED // 4), out=llr_low[b0:b0 + nb])
                x = llr_low
            for rep in range(2):
                e0.record(stream)
                _lib.check(L.nrldpc_decode_minsum(x.data_ptr(), B, BGN, ZC, MAX_ITER, al, be, et, None, info.data_ptr(),
                                                  status.data_ptr(), iters.data_ptr(), stream.cuda_stream), "decode")
                e1.record(stream)
                torch.cuda.synchronize()
            extra.append({"run": name, "gbit_per_s": bits_b / (e0.elapsed_time(e1) * 1e-3) / 1e9,
                          "mean_iters": float(iters.float().mean()), "parity_ok_frac": float(status.float().mean())})
        del llr_low
        step()   # restore the headline outputs for the correctness check below
        torch.cuda.synchronize()

    other = []
    agreement = []
    if rank == 0 and not args.no_extra:
        other = other_kernel_lines(torch, engine, dev, peaks()[0])
        agreement = fp32_vs_fp64_agreement(np, torch, engine, dev)

    # correctness of what was timed: decoded info bits vs what was sent; counters reduced over ranks
    got = ((info.view(torch.uint8).unsqueeze(-1) >> torch.arange(8, device=dev, dtype=torch.uint8)) & 1).reshape(B, -1)[:, :K_INFO]
    blk_err = (got != sent.to(torch.uint8)).any(1)
    counters = torch.stack([torch.tensor(B, device=dev), blk_err.sum(), (got != sent.to(torch.uint8)).sum(),
                            iters.sum(), status.sum()]).to(torch.int64)
    if world > 1:
        dist.all_reduce(counters)  # the path's only collective: 5 int64 counters
    cnt = counters.tolist()

    # ---- e2e through the host-buffer C-ABI entry point, pinned host memory
    Be = min(args.e2e_batch, B)
    # pinned buffers are placed on the NUMA node next to this rank's GPU (process affinity while they are allocated)
    cores_before = os.sched_getaffinity(0) if hasattr(os, "sched_getaffinity") else None
    bound = engine.bind_host_to_device(local)
    h_llr = torch.empty((Be, N_CODED), dtype=torch.float32).pin_memory()
    h_llr.copy_(llr[:Be])
    h_info = torch.empty((Be, (K_INFO + 31) // 32), dtype=torch.int32).pin_memory()
    h_st = torch.empty((Be,), dtype=torch.uint8).pin_memory()
    h_it = torch.empty((Be,), dtype=torch.int32).pin_memory()

    def e2e_step():
        _lib.check(L.nrldpc_decode_minsum_host(h_llr.data_ptr(), Be, BGN, ZC, MAX_ITER, ALPHA, BETA, 0, None,
                                               h_info.data_ptr(), h_st.data_ptr(), h_it.data_ptr()), "decode_host")

    e2e_step()
    barrier()
    e_steps = max(1, min(args.steps, 3))
    t0 = time.perf_counter()
    for _ in range(e_steps):
        e2e_step()
    barrier()
    e_dt = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(e_dt, op=dist.ReduceOp.MAX)
    e2e_val = world * Be * e_steps * K_INFO / float(e_dt.item()) / 1e9
    assert torch.equal(h_info.to(dev), info[:Be]), "host path and device path disagree"

    # ---- the same call from PAGEABLE host memory (what a NumPy caller of engine.decode_batch hands over): the library
    # stages it through its ring of pinned slots with its copy threads
    p_llr = np.empty((Be, N_CODED), np.float32)
    p_llr[...] = h_llr.numpy()
    p_info = np.empty((Be, (K_INFO + 31) // 32), np.int32)
    p_st, p_it = np.empty(Be, np.uint8), np.empty(Be, np.int32)

    def e2e_pageable_step():
        _lib.check(L.nrldpc_decode_minsum_host(p_llr.ctypes.data, Be, BGN, ZC, MAX_ITER, ALPHA, BETA, 0, None,
                                               p_info.ctypes.data, p_st.ctypes.data, p_it.ctypes.data), "decode_host")

    e2e_pageable_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e_steps):
        e2e_pageable_step()
    barrier()
    p_dt = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(p_dt, op=dist.ReduceOp.MAX)
    e2e_pageable = world * Be * e_steps * K_INFO / float(p_dt.item()) / 1e9
    assert np.array_equal(p_info, h_info.numpy()), "pageable and pinned host paths disagree"

    # ---- the same call with the LLRs stored as IEEE half floats in pinned host memory (nrldpc_decode_minsum_host_f16): half the
    # bytes on the host link, widened on the device.  A different INPUT FORMAT (the LLRs are rounded to 11 significant bits
    # before they reach the decoder), reported next to -- not instead of -- the fp32 figure.
    h16 = torch.empty((Be, N_CODED), dtype=torch.float16).pin_memory()
    h16.copy_(h_llr)
    h_info16 = torch.empty_like(h_info).pin_memory()

    def e2e_f16_step():
        _lib.check(L.nrldpc_decode_minsum_host_f16(h16.data_ptr(), Be, BGN, ZC, MAX_ITER, ALPHA, BETA, 0, None,
                                                   h_info16.data_ptr(), h_st.data_ptr(), h_it.data_ptr()), "decode_host_f16")

    e2e_f16_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e_steps):
        e2e_f16_step()
    barrier()
    f_dt = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(f_dt, op=dist.ReduceOp.MAX)
    e2e_f16 = world * Be * e_steps * K_INFO / float(f_dt.item()) / 1e9
    f16_word_agreement = float((h_info16 == h_info).float().mean())
    if bound is not None:
        os.sched_setaffinity(0, cores_before)

    link = host_link_ceiling(torch, dev, h_llr, barrier, world, dist)
    tb_line = transport_block_line(np, engine) if (rank == 0 and not args.no_extra) else None
    if rank == 0:
        peak, peak_src = peaks()
        per_gpu_cbs = cbs_per_s / world
        achieved = per_gpu_cbs * ALGO_BYTES_PER_CB / 1e9
        G, nt, smem = (ctypes_int() for _ in range(3))
        L.nrldpc_decode_minsum_geometry(BGN, ZC, G, nt, smem)
        out = {
            "metric": METRIC, "value": value, "unit": "Gbit/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "codeblocks_per_gpu_per_step": B, "info_bits_per_codeblock": K_INFO,
                       "l2_policy": f"inputs larger than L2 ({B * N_CODED * 4 / 2**30:.1f} GiB of LLRs per step)",
                       "kernel_geometry": {"codeblocks_per_cta": G.value, "threads": nt.value, "smem_bytes": smem.value},
                       "block_error_rate": cnt[1] / cnt[0], "mean_iters": cnt[3] / cnt[0], "parity_ok_frac": cnt[4] / cnt[0],
                       "other_runs_1gpu_untimed_region": extra, "other_kernels": other,
                       "fp32_vs_float64_reference_agreement": agreement,
                       "transport_block": tb_line},
            "e2e": {"value": e2e_val, "unit": "Gbit/s", "h2d_bytes_per_step": Be * N_CODED * 4,
                    "d2h_bytes_per_step": Be * ((K_INFO + 31) // 32 * 4 + 1 + 4), "codeblocks_per_step": Be,
                    "host_cores_bound_to_gpu_numa_node": len(bound) if bound is not None else None,
                    "host_memory": "pinned", "pageable_value": e2e_pageable, "pageable_over_pinned": e2e_pageable / e2e_val,
                    "host_link_ceiling": link,
                    "half_precision_llr_input": {"value": e2e_f16, "unit": "Gbit/s", "h2d_bytes_per_step": Be * N_CODED * 2,
                                                 "info_words_equal_to_fp32_input": f16_word_agreement,
                                                 "note": "nrldpc_decode_minsum_host_f16: the caller stores its LLRs as IEEE half floats; "
                                                         "another input format, not the headline"},
                    "pageable_note": "same call, LLRs in pageable NumPy memory: staged through the library's pinned ring by its copy threads"},
            "gpu_launches": args.steps,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": TRAFFIC_BYTES_PER_CB * B, "peak_source": peak_src,
                         "traffic_source": f"{NCU_CAPTURE}: (dram__bytes_read.sum + dram__bytes_write.sum) / 592 codeblocks x this "
                                           "launch's codeblocks (cited from the committed capture, not measured by this run)",
                         "algorithmic_bytes_per_launch": B * ALGO_BYTES_PER_CB,
                         "kernel": "decode_spec_kernel<Code<1,384>,ET=0,B0=1>", "kernel_ms": ms / args.steps,
                         "note": "HBM is not the binding roof of this kernel (10 on-chip iterations per byte): the check "
                                 "pass is bound by the SM ALU pipe / issue slots, the variable pass by shared-memory "
                                 "wavefronts; see DESIGN.md and profiles/"},
            "clocks": clk.summary(),
        }
        # the roof that actually binds: warp-instruction issue slots (148 SMs x 4 schedulers x SM clock)
        clk_mhz = out["clocks"].get("sm_mhz") or 1965
        sms = torch.cuda.get_device_properties(dev).multi_processor_count
        issue_peak = sms * 4 * clk_mhz * 1e6
        out["roofline_onchip"] = {"bound": "issue_slots", "achieved": per_gpu_cbs * WARP_INSTR_PER_CB / 1e9,
                                  "peak": issue_peak / 1e9, "unit": "G warp-instr/s", "frac": per_gpu_cbs * WARP_INSTR_PER_CB / issue_peak,
                                  "warp_instr_per_codeblock": WARP_INSTR_PER_CB,
                                  "source": f"smsp__inst_executed.sum of {NCU_CAPTURE} / 592 codeblocks (cited, not measured by this run); "
                                            "ALU pipe 65 %, issue 80 %, smem wavefronts 63 % busy in that capture"}
        if not args.no_cpu and world == 1:   # reported at N=1 only (rank 0); ~10 s of CPU work on all host cores
            ncores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
            gbps, threads, n, sps = cpu_reference_run(1, 0, sample_cbs=max(400 * ncores, 256))
            out["cpu_baseline"] = {"value": gbps, "unit": "Gbit/s", "cores": threads, "kind": "port",
                                   "sample": f"{n} codeblocks of the same workload, float64 C port of the reference (oracle/), "
                                             f"{threads} OpenMP threads, {sps:.2f} s"}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def host_link_ceiling(torch, dev, h_llr, barrier, world, dist):
    """What the box's host links can carry: the e2e step's LLR bytes as plain pinned cudaMemcpyAsync H2D copies (no
    kernels), all ranks at once, max over ranks -> GB/s over all links and the e2e value that bandwidth would allow."""
    d = torch.empty_like(h_llr, device=dev)
    chunk = 331   # the host path's chunk: 32 MiB of LLRs
    def run():
        for b0 in range(0, h_llr.shape[0], chunk):
            d[b0:b0 + chunk].copy_(h_llr[b0:b0 + chunk], non_blocking=True)
    run()
    barrier()
    t0 = time.perf_counter()
    for _ in range(3):
        run()
    barrier()
    dt = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    nbytes = h_llr.numel() * 4
    gbs = world * 3 * nbytes / float(dt.item()) / 1e9
    return {"h2d_gb_per_s_all_ranks": gbs, "e2e_gbit_per_s_at_that_rate": gbs * 1e9 / (4 * N_CODED) * K_INFO / 1e9,
            "how": "plain pinned cudaMemcpyAsync of the e2e step's LLR bytes, 32 MiB chunks, all ranks concurrently, 3 passes"}


def other_kernel_lines(torch, engine, dev, hbm_peak):
    """Encoder, bit-flipping and sum-product lines (SURVEY 8(d) byte conventions), CUDA events, batches larger than L2."""
    out = []
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]

    def timed(f, reps=5):
        f()
        torch.cuda.synchronize()
        ev[0].record()
        for _ in range(reps):
            f()
        ev[1].record()
        torch.cuda.synchronize()
        return ev[0].elapsed_time(ev[1]) / reps

    for bgn, Zc, B in [(1, 384, 1 << 16), (2, 384, 1 << 16), (1, 208, 1 << 16)]:
        K, N, Nf, M = engine.dims(bgn, Zc)
        ck = engine.random_bits(B, K, seed=11, device=dev)
        ms = timed(lambda: engine.encode_batch(ck, bgn, Zc, fix_fillers=False))
        api, packed = B * (K + N), B * (K + N) // 8
        out.append({"kernel": f"encode BG{bgn} Zc={Zc}", "codeblocks": B, "ms": ms, "info_tbit_per_s": B * K / ms / 1e9,
                    "bytes_int8_api": api, "gb_per_s_int8_api": api / ms / 1e6, "frac_hbm_int8_api": api / ms / 1e6 / hbm_peak,
                    "bytes_packed_8d": packed, "gb_per_s_packed_8d": packed / ms / 1e6, "frac_hbm_packed_8d": packed / ms / 1e6 / hbm_peak})
        del ck
    # the bit-packed encoder: what moves IS SURVEY 8(d)'s K/8 + N/8 bytes per codeblock (batch > L2: 1 << 18 codeblocks = 1.1 GB)
    for bgn, Zc, B in [(1, 384, 1 << 18), (2, 384, 1 << 18)]:
        K, N, Nf, M = engine.dims(bgn, Zc)
        ckw = engine.random_bits_packed(B, K, seed=11, device=dev)
        ms = timed(lambda: engine.encode_packed(ckw, bgn, Zc))
        packed = B * (K + N) // 8
        out.append({"kernel": f"encode_packed BG{bgn} Zc={Zc}", "codeblocks": B, "ms": ms, "info_tbit_per_s": B * K / ms / 1e9,
                    "bytes_packed_8d": packed, "gb_per_s_packed_8d": packed / ms / 1e6, "frac_hbm_packed_8d": packed / ms / 1e6 / hbm_peak})
        del ckw
    bgn, Zc, B, L = 1, 384, 1 << 14, 20
    K, N, Nf, M = engine.dims(bgn, Zc)
    ck = engine.random_bits(B, K, seed=12, device=dev)
    llr = engine.awgn_llr(engine.encode_batch(ck, bgn, Zc), 1.0, seed=13)
    ms = timed(lambda: engine.decode_bf_batch(llr, Zc, bgn, L), reps=3)
    nbytes = B * (4 * N + Nf)
    out.append({"kernel": f"bit-flipping BG{bgn} Zc={Zc} L={L} (nothing converges at +1 dB)", "codeblocks": B, "ms": ms,
                "info_gbit_per_s": B * K / ms / 1e6, "bytes_8d": nbytes, "gb_per_s": nbytes / ms / 1e6, "frac_hbm": nbytes / ms / 1e6 / hbm_peak})
    # sum-product (algo='BP', float64 like the reference): issue- and latency-bound, profiles/r2_bp_ncu_summary.md
    Bp, Lp = 1 << 12, 10
    ms = timed(lambda: engine.decode_bp_batch(llr[:Bp], Zc, bgn, Lp), reps=2)
    nbytes = Bp * (4 * N + Nf)
    out.append({"kernel": f"sum-product BG{bgn} Zc={Zc} L={Lp} (early termination, +1 dB)", "codeblocks": Bp, "ms": ms,
                "info_gbit_per_s": Bp * K / ms / 1e6, "bytes_8d": nbytes, "gb_per_s": nbytes / ms / 1e6, "frac_hbm": nbytes / ms / 1e6 / hbm_peak})
    return out


def fp32_vs_fp64_agreement(np, torch, engine, dev, n=1536):
    """Codeblock agreement of the fp32 kernel with the float64 arithmetic of the reference (the generic float64 kernel,
    bit-identical to the reference on every golden vector) on identical fp32-representable LLRs: hard bits, status and
    iteration count all equal.  North-star bar: >= 99.99 % -- met where codeblocks converge; the non-converged ones of
    a BLER ~ 1 point differ in a bit or two at the 1e-3 level."""
    out = []
    K, N, Nf, M = engine.dims(BGN, ZC)
    ck = engine.random_bits(n, K_INFO, seed=21, device=dev)
    dn = engine.encode_batch(ck, BGN, ZC)
    for snr in (-3.0, -0.15, 1.0):
        llr = engine.awgn_llr(dn, snr, seed=22)
        r = engine.decode_batch(llr, ZC, BGN, MAX_ITER, ALPHA, BETA, True)
        c64, s64, i64 = engine.decode_ref_batch(llr.cpu().numpy().astype(np.float64), ZC, BGN, MAX_ITER, "min-sum", ALPHA, BETA, True, f64=True)
        c32, s32, i32 = r["ck"].cpu().numpy(), r["status"].cpu().numpy().astype(bool), r["iters"].cpu().numpy()
        same = (c32 == c64).all(1) & (s32 == s64) & (i32 == i64)
        conv = s64
        blk32, blk64 = (c32[:, :K_INFO] != ck.cpu().numpy()).any(1), (c64[:, :K_INFO] != ck.cpu().numpy()).any(1)
        out.append({"snr_db": snr, "codeblocks": n, "bler_float64": float((c64[:, :K_INFO] != ck.cpu().numpy()).any(1).mean()),
                    "agreement": float(same.mean()), "agreement_converged": float(same[conv].mean()) if conv.any() else None,
                    "agreement_not_converged": float(same[~conv].mean()) if (~conv).any() else None,
                    "status_and_iterations_equal": float(((s32 == s64) & (i32 == i64)).mean()),
                    "block_error_decisions_equal": float((blk32 == blk64).mean()),
                    "differing_bits_total": int((c32 != c64).sum()), "bits_total": int(c64.size)})
    return out


def transport_block_line(np, engine):
    """BASELINE config #4 through the reference-facing functions, host buffers in and out: a 273-PRB 256QAM PDSCH
    transport block (TBS 966 896, 115 codeblocks of BG1 Zc=384, E = 10 939 per codeblock, fillers 16)."""
    from python_5gtoolbox_b200.nr_pdsch import nr_dlsch, nr_dlsch_decode
    rng = np.random.default_rng(4)
    cfg = {"L": 10, "algo": "min-sum", "alpha": 0.8, "beta": 0.0}
    A, R, Qm, NL, G, LBRM = 966896, 948, 8, 4, 273 * 12 * 12 * 8 * 4, 10 ** 9
    trblk = rng.integers(0, 2, A).astype("i1")
    g = nr_dlsch.DLSCHEncode(trblk, A, Qm, R, NL, 0, LBRM, G)
    sigma = 10 ** (-6.5 / 20)
    llr = (2 * ((1 - 2 * g.astype("f4")) + rng.normal(0, sigma, G).astype("f4")) / sigma ** 2).astype("f4")

    def t(f, reps=20):
        for _ in range(4):
            r = f()
        t0 = time.perf_counter()
        for _ in range(reps):
            r = f()
        return (time.perf_counter() - t0) / reps, r

    t_enc, g2 = t(lambda: nr_dlsch.DLSCHEncode(trblk, A, Qm, R, NL, 0, LBRM, G))
    t_dec, (st, tb, new) = t(lambda: nr_dlsch_decode.DLSCHDecode(llr, A, Qm, R, NL, 0, LBRM, cfg))
    ok = bool(st) and np.array_equal(tb, trblk) and np.array_equal(g, g2)
    return {"workload": "BASELINE config #4: DLSCHEncode / DLSCHDecode of a 273-PRB 256QAM 4-layer transport block (TBS 966896 = 115 "
                        "codeblocks BG1 Zc=384), NumPy host buffers in and out, float64 [115,25344] soft buffer returned",
            "ok": ok, "DLSCHDecode_ms": t_dec * 1e3, "DLSCHEncode_ms": t_enc * 1e3, "decode_tb_per_s": 1 / t_dec,
            "decode_gbit_per_s_tb_bits": A / t_dec / 1e9, "encode_gbit_per_s_tb_bits": A / t_enc / 1e9,
            "h2d_bytes_per_tb": int(llr.nbytes), "d2h_bytes_per_tb": int(new.nbytes + tb.nbytes),
            "note": "the 23.3 MB float64 soft buffer (HARQ state, the reference's return value) is stored by the decoder straight "
                    "into pinned host memory and bounds the call at the PCIe rate"}


def run_mc(args):
    """--workload mc: the device-resident Monte-Carlo chain of BASELINE config #5 (sim.bler_curve: Philox bits -> CRC24A ->
    encode -> BPSK/AWGN -> min-sum decode with the reference's early termination -> error counters), sharded over the
    ranks by codeblock index; the only collective is the final all-reduce of int64[points x 4].  A codeblock's bits and
    noise depend only on (seed, point, index): the summed counters are identical for every number of GPUs."""
    import torch
    import torch.distributed as dist
    from python_5gtoolbox_b200 import sim
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    snrs, n = [0.0, 1.0, 2.0], args.mc_codeblocks

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(1, min(args.warmup, 2))):
        sim.bler_curve(ZC, BGN, snrs, min(n, 16384 * world), MAX_ITER, ALPHA, BETA, device=dev)
    barrier()
    with ClockSampler(local) as clk:
        t0 = time.perf_counter()
        for _ in range(args.steps):
            rows = sim.bler_curve(ZC, BGN, snrs, n, MAX_ITER, ALPHA, BETA, device=dev)
        barrier()
        dt = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    sec = float(dt.item()) / args.steps
    total = n * len(snrs)
    if rank == 0:
        print(json.dumps({
            "metric": METRIC, "value": total * K_INFO / sec / 1e9, "unit": "Gbit/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "Monte-Carlo chain (BASELINE config #5 analogue on AWGN): BG1 Zc=384, Philox bits -> CRC24A -> encode -> "
                                   "BPSK/AWGN -> NMS(0.8) min-sum, L=10, early termination -> error counters, all on the device",
                       "snr_db": snrs, "codeblocks_per_point": n, "codeblocks_per_s": total / sec,
                       "counters": [{k: r[k] for k in ("snr_db", "codeblocks", "block_errors", "bit_errors", "bler", "mean_iters")} for r in rows],
                       "collective": "one all_reduce(SUM) of int64[3 x 4] per curve (NCCL)"},
            "gpu_launches": args.steps * len(snrs) * 6 * (-(-(n // world) // 16384)), "clocks": clk.summary()}))
    if world > 1:
        dist.destroy_process_group()


def ctypes_int():
    import ctypes
    return ctypes.c_int()


# Cited, not measured by this run: the committed `ncu --set full` capture of the same kernel, profiles/prof_r2_decode.ncu-rep
# (592 codeblocks; raw page profiles/prof_r2_decode_raw.csv, summary profiles/r2_ncu_summary.md):
#   dram__bytes_read.sum + dram__bytes_write.sum = 60 185 088 B + 399 616 B;  smsp__inst_executed.sum = 469 770 352
NCU_CAPTURE = "profiles/prof_r2_decode.ncu-rep"
TRAFFIC_BYTES_PER_CB = (60185088 + 399616) / 592
WARP_INSTR_PER_CB = 469770352 / 592


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=1 << 16, help="codeblocks per GPU per step")
    ap.add_argument("--e2e-batch", type=int, default=1 << 13)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the secondary early-termination / OMS / mixed runs")
    ap.add_argument("--workload", default="decode", choices=["decode", "mc"],
                    help="decode: the headline decoder bench (default); mc: the device-resident Monte-Carlo chain, sharded over the ranks")
    ap.add_argument("--mc-codeblocks", type=int, default=200000, help="--workload mc: codeblocks per SNR point (whole job)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "mc":
        run_mc(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
