#!/usr/bin/env python3
"""bench.py -- BN254 batch-verify throughput on B200 (BASELINE.json configs[1]) and the CPU reference arm.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl engine|reference] [--n ITEMS] [--no-extras]

A step = one pass of ECDSA::verify over a batch of 2^20 independent (32-byte msg, sig, pk) triples per GPU
(weak scaling: every rank verifies its own 2^20 triples, no data-path collective).  Prints ONE JSON line:
  value       verifies/s, whole job, inputs resident in HBM when the timed region starts (CUDA events on the
              engine's stream, max over ranks)
  e2e         the same metric through the public host-buffer entry point (bn254_verify_batch via
              bn254_b200.engine.verify_batch): pinned host inputs, H2D + kernels + D2H of the verdicts timed
  roofline    INT32 multiply-pipe roofline of the dominant kernel (k_coop4_run: Miller accumulation + final
              exponentiation): achieved = algorithmic IMAD32 (264 per Fq product) per second over the
              kernel's CUDA-event time, peak = the IMAD issue rate measured live by tools/microbench.bin on this GPU
              (the path is integer-issue bound: a verify reads 224 B and does ~5.8 M IMAD32-equivalents, so neither
              HBM nor tensor peak applies; the HBM figures are reported beside it to show that)
  cpu_baseline  the oracle (C restatement of the dependency's algorithms) timed on the host cores, bounded sample
  configs     the other BASELINE.json configs measured in the same run at this N (not the headline): config 3 hash + sign,
              config 4 same-message aggregate (keys sharded over the ranks, one all-gather), config 5 distinct-message
              aggregate (2^22 pairs sharded over the ranks, one all-gather, ONE final exponentiation: strong scaling),
              pairings/s (bn254_pairing_check_batch, k = 1), the untrusted-input policy, and a small-batch latency table
`--impl reference` times the CPU path alone (the Rust crate cannot be built here: no cargo/rustc; oracle/ port).
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

# Fq-product equivalents per unit, the fixed numerators of SURVEY.md 8(d), split by the kernel that does the work
M_HASH = 774            # hash kernels: try-and-increment, 2.116 expected tries
M_LINES = 3083          # k_verify_lines: G2 doubling / addition steps (64 x 28 + 23 x 41) + scaling of the -G2 lines (87 x 4)
M_COOP = 18028          # k_coop4_run: f^2 chain 2 304 + 2 x 87 sparse products x 39 + final exponentiation 8 938
M_VERIFY = M_HASH + M_LINES + M_COOP   # 21 885
M_PAIRING = 17370       # one pairing: Miller loop 8 432 + final exponentiation 8 938
M_DISTINCT_PAIR = 6902  # distinct-message aggregate, per pair (hash + variable-Q Miller, squarings amortised)
M_SIGN = 774 + 2610     # hash + 254-bit variable-base G1 scalar multiplication (w = 4 signed window count of SURVEY 8d)
IMAD_PER_M = 264        # IMAD32 issue slots per Fq product (an IMAD.WIDE.U32.X costs two: profiles/r01_tuning_log.md)
METRIC = "bn254_verifies_per_sec"
UNIT = "verifies/s"
WORKLOAD = "batch verify 2^20 independent (32-byte msg, sig, pk) triples per GPU (BASELINE configs[1])"


def shared_config(n):
    """`config` of BOTH arms (identical keys and values, so that a driver comparing the two lines sees the same configuration); what is
    specific to one arm lives next to it (`engine`, `cpu_baseline.sample`)."""
    return {"workload": WORKLOAD, "triples_per_gpu": n, "msg_len": 32, "input_policy": "typed (already-decoded points, SURVEY 8d config 2)",
            "l2": "engine arm: inputs + line-set workspace (%.1f GB per step, written and read once) larger than L2; CPU arm: each step verifies a "
                  "bounded sample of the workload (cpu_baseline.sample)" % ((224 + 50112 + 2304) * n / 1e9)}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--n", type=int, default=1 << 20, help="triples per GPU per step")
    ap.add_argument("--cpu-sample", type=int, default=0, help="verifies in the CPU baseline sample (0 = auto)")
    ap.add_argument("--no-extras", action="store_true", help="skip the `configs` record (tuning runs)")
    ap.add_argument("--distinct-log2", type=int, default=22, help="total pairs of config 5 (all ranks together)")
    return ap.parse_args()


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.samples, self.reasons, self.max_mhz = index, False, [], set(), None

    def run(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            names = {
                pynvml.nvmlClocksEventReasonSwPowerCap if hasattr(pynvml, "nvmlClocksEventReasonSwPowerCap") else 0x4: "sw_power_cap",
                0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown",
            }
            while not self.stop_flag:
                self.samples.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                try:
                    r = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    r = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
                time.sleep(0.1)
        except Exception as e:  # no NVML: report that instead of inventing clocks
            self.reasons.add("nvml_unavailable:%s" % type(e).__name__)

    def result(self):
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def cpu_baseline(n_sample, msgs, sigs, pks, threads):
    """Oracle verify on the host cores over the first n_sample triples of the workload (the checker timed, never shipped)."""
    import oracle_lib as O
    O.verify_batch(msgs[:32 * 8], 32, sigs[:64 * 8], pks[:128 * 8], 8, threads)  # warm (one-time init)
    t = time.perf_counter()
    st = O.verify_batch(msgs[:32 * n_sample], 32, sigs[:64 * n_sample], pks[:128 * n_sample], n_sample, threads)
    dt = time.perf_counter() - t
    return n_sample / dt, dt, st


def cpu_single_thread_ms(msgs, sigs, pks, k=24):
    """one CPU core, one verify at a time: the latency a caller of the crate's own one-item API sees (ms per verify)"""
    import oracle_lib as O
    O.verify(msgs[:32], sigs[:64], pks[:128])
    t = time.perf_counter()
    for i in range(k):
        assert O.verify(msgs[32 * i:32 * i + 32], sigs[64 * i:64 * i + 64], pks[128 * i:128 * i + 128]) == 0
    return (time.perf_counter() - t) / k * 1e3


def run_reference(args):
    """CPU arm: the reference's own CPU implementation cannot be built here (Rust; dependency not vendored), so this
    times the oracle port with every host thread, each step a bounded sample of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle_lib as O
    import synth
    threads = os.cpu_count() or 1
    n = args.cpu_sample or max(256, 4096 * threads)  # ~10 s of CPU work per step at ~0.5 k verifies/s/thread
    msgs, sks = synth.messages(n, 32, seed=1), synth.secret_keys(n, seed=2)
    sigs, st = O.sign_batch(msgs, 32, sks, n, threads)
    pks = O.derive_pk_g2_batch(sks, n, threads)
    for _ in range(max(1, min(args.warmup, 1))):
        O.verify_batch(msgs, 32, sigs, pks, min(n, 64), threads)
    t = time.perf_counter()
    for _ in range(args.steps):
        out = O.verify_batch(msgs, 32, sigs, pks, n, threads)
    dt = time.perf_counter() - t
    assert out == bytes(n)
    v = n * args.steps / dt
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32",
        "data": "synthetic",
        "config": shared_config(1 << 20), "reference_sample_per_step": n,
        "cpu_baseline": {"value": v, "per_core": v / threads, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": "%d verifies per step x %d steps, oracle/bn254_oracle.c on %d threads" % (n, args.steps, threads)},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def imad_peak():
    """Measured INT32 IMAD issue rate of this GPU (ops/s) from tools/microbench.bin; None if it cannot run."""
    exe = os.path.join(ROOT, "tools", "microbench.bin")
    try:
        out = subprocess.run([exe], capture_output=True, text=True, timeout=300, check=True).stdout.strip().splitlines()[-1]
        return json.loads(out)
    except Exception as e:
        return {"error": "%s: %s" % (type(e).__name__, e)}


def run_engine(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import bn254_b200
    from bn254_b200 import dist as D
    from bn254_b200 import engine as E
    from bn254_b200._native import I, S
    import synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = E.context(local)
    # configs[1]: "inputs are already-decoded affine points (the Rust API takes &Signature / &PublicKey)": values of
    # the crate's types, i.e. the engine's typed input policy.  The untrusted policy (the engine's default, with the r-torsion test
    # per key) is measured separately in `configs`.
    E.set_input_policy(E.INPUTS_TYPED, ctx=ctx)
    n = args.n

    # ---- synthetic workload: seeded messages and keys; signatures and keys made by the engine
    seed = 1 + 1000 * rank
    msgs, sks = synth.messages(n, 32, seed=seed), synth.secret_keys(n, seed=seed + 1)
    sigs, st = E.sign_batch(msgs, 32, sks, ctx=ctx)
    assert not any(st)
    pks = E.derive_pk_g2_batch(sks, ctx=ctx)

    def dev(b):
        return torch.frombuffer(bytearray(b), dtype=torch.uint8).cuda()

    def pinned(b):
        return torch.frombuffer(bytearray(b), dtype=torch.uint8).pin_memory()

    d_msgs, d_sigs, d_pks = dev(msgs), dev(sigs), dev(pks)
    d_st = torch.empty(n, dtype=torch.uint8, device="cuda")
    h_msgs, h_sigs, h_pks = pinned(msgs), pinned(sigs), pinned(pks)
    h_st = torch.empty(n, dtype=torch.uint8).pin_memory()
    stream = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", local))
    torch.cuda.synchronize()

    def step_dev():
        ctx.call("bn254_verify_batch_dev", d_msgs, S(32), d_sigs, d_pks, S(n), d_st)

    def step_e2e():
        ctx.call("bn254_verify_batch", h_msgs, S(32), h_sigs, h_pks, S(n), h_st)

    def barrier():
        torch.cuda.synchronize()
        ctx.sync()
        if world > 1:
            dist.barrier()

    def timed(fn, steps):
        """CUDA-event time (ms) of `steps` calls on the engine's stream, barrier + synchronize on both sides, max over ranks"""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        ctx.sync()
        e1.synchronize()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        barrier()
        return ms

    for _ in range(args.warmup):
        step_dev()
    barrier()
    nocheck = bool(os.environ.get("BN254_BENCH_NOCHECK"))  # timing-only ablation builds (tuning; their results are wrong on purpose)
    assert nocheck or bytes(d_st.cpu().numpy().tobytes()) == bytes(n), "engine rejected valid signatures"

    sampler = ClockSampler(local)
    sampler.start()
    launches0 = ctx.launch_count
    ms_total = timed(step_dev, args.steps)
    launches = ctx.launch_count - launches0
    # per-phase device time of the same pipeline (events between its kernels), one extra profiled step
    phase = (ctypes.c_float * 3)()
    ctx.call("bn254_set_profiling", I(1))
    step_dev()
    ctx.call("bn254_phase_ms", phase)
    ctx.call("bn254_set_profiling", I(0))
    # end to end through the host-buffer entry point
    step_e2e()
    ms_e2e = timed(step_e2e, args.steps)
    sampler.stop_flag = True
    sampler.join(timeout=2)
    assert nocheck or bytes(h_st.numpy().tobytes()) == bytes(n)

    value = world * n * args.steps / (ms_total * 1e-3)
    e2e = world * n * args.steps / (ms_e2e * 1e-3)

    # ------------------------------------------------------------------------------------------ the other configs, same run
    configs = None
    if not args.no_extras:
        configs = {}
        reps = 2

        # config 3: batch hash_to_g1 + sign, distinct keys (device-resident, weak scaling like the headline)
        d_sks = dev(sks)
        d_out = torch.empty(64 * n, dtype=torch.uint8, device="cuda")
        sign_step = lambda: ctx.call("bn254_sign_batch_dev", d_msgs, S(32), d_sks, S(n), d_out, d_st)
        hash_step = lambda: ctx.call("bn254_hash_to_g1_batch_dev", d_msgs, S(32), S(n), d_out, d_st)
        sign_step()
        ms_sign = timed(sign_step, reps) / reps
        assert nocheck or bytes(d_out.cpu().numpy().tobytes()) == sigs
        hash_step()
        ms_hash = timed(hash_step, reps) / reps
        configs["config3_hash_sign"] = {"signs_per_sec": world * n / (ms_sign * 1e-3), "ms_per_step": ms_sign, "messages_per_gpu": n,
                                        "hash_to_g1_per_sec": world * n / (ms_hash * 1e-3), "scaling": "weak",
                                        "checked": "signatures equal the ones the workload was built from (oracle-sampled in tests)"}
        del d_out

        # pairings/s: bn::pairing_batch of ONE pair per item through the cooperative machine (Miller loop + final exponentiation)
        ctx.call("bn254_pairing_check_batch_dev", d_sigs, d_pks, S(1), S(n), d_st)
        ms_pair = timed(lambda: ctx.call("bn254_pairing_check_batch_dev", d_sigs, d_pks, S(1), S(n), d_st), reps) / reps
        assert nocheck or bytes(d_st.cpu().numpy().tobytes()) == bytes([9]) * n  # e(sig, pk) is not one; every item decoded
        configs["pairings"] = {"pairings_per_sec": world * n / (ms_pair * 1e-3), "ms_per_step": ms_pair, "pairs_per_item": 1, "items_per_gpu": n,
                               "entry_point": "bn254_pairing_check_batch_dev"}

        # the engine's default input policy: every key gets the r-torsion test, (0, 0) is rejected
        E.set_input_policy(E.INPUTS_UNTRUSTED, ctx=ctx)
        step_dev()
        ms_strict = timed(step_dev, reps) / reps
        assert nocheck or bytes(d_st.cpu().numpy().tobytes()) == bytes(n)
        E.set_input_policy(E.INPUTS_TYPED, ctx=ctx)
        configs["config2_untrusted_inputs"] = {"verifies_per_sec": world * n / (ms_strict * 1e-3), "ms_per_step": ms_strict,
                                               "note": "bn254_set_input_policy default: from_uncompressed semantics incl. G2 r-torsion test per key"}

        # NOT config 2: the same triples when the keys are a fixed set known in advance (validators): their line coefficients are
        # cached once (bn254_key_lines_prepare_dev, 16.7 KB per key) and a verify only scales them -- k_verify_lines disappears
        t0 = time.perf_counter()
        cache = E.KeyLineCache(pks, ctx=ctx)
        t_cache = (time.perf_counter() - t0) * 1e3
        assert nocheck or not any(cache.key_status())
        cstep = lambda: cache.verify_dev(d_msgs, 32, d_sigs, n, d_st)
        cstep()
        ms_cached = timed(cstep, reps) / reps
        assert nocheck or bytes(d_st.cpu().numpy().tobytes()) == bytes(n)
        configs["fixed_key_set_cached_lines"] = {"verifies_per_sec": world * n / (ms_cached * 1e-3), "ms_per_step": ms_cached, "keys": n,
                                                 "cache_bytes_per_key": 16704, "cache_build_ms_incl_upload": t_cache,
                                                 "note": "additional entry point bn254_verify_batch_cached_dev; same statuses as verify_batch; not the headline config"}
        del cache

        # config 4: same-message aggregate, 2^20 keys in total sharded over the ranks (strong scaling)
        n4 = (1 << 20) // world
        msg4 = synth.messages(1, 32, seed=77)
        d_msg4 = dev(msg4)
        d_sig4 = torch.empty(64 * n4, dtype=torch.uint8, device="cuda")
        ctx.call("bn254_sign_batch_dev", dev(msg4 * n4), S(32), d_sks[:32 * n4], S(n4), d_sig4, d_st[:n4])
        same = D.SameMessageAggregate(ctx)
        same.step(d_msg4, 32, d_sig4, d_pks[:128 * n4], n4)
        v4 = same.status()
        ms4 = timed(lambda: same.step(d_msg4, 32, d_sig4, d_pks[:128 * n4], n4), 3) / 3
        configs["config4_same_message"] = {"keys_total": n4 * world, "ms_per_step": ms4, "points_per_sec": 2 * n4 * world / (ms4 * 1e-3),
                                           "verdict": v4, "scaling": "strong", "exchange_bytes_per_rank": 256,
                                           "path": "per-rank bn254_g{1,2}_sum_dev -> all_gather_into_tensor -> bn254_aggregate_verify_same_msg_dev"}
        del d_sig4

        # config 5: distinct-message aggregate, 2^22 pairs in total sharded over the ranks, ONE final exponentiation
        total5 = 1 << args.distinct_log2
        n5 = total5 // world
        msgs5, sks5 = synth.messages(n5, 32, seed=500 + rank), synth.secret_keys(n5, seed=600 + rank)
        d_m5, d_k5 = dev(msgs5), dev(sks5)
        d_s5 = torch.empty(64 * n5, dtype=torch.uint8, device="cuda")
        d_st5 = torch.empty(n5, dtype=torch.uint8, device="cuda")
        ctx.call("bn254_sign_batch_dev", d_m5, S(32), d_k5, S(n5), d_s5, d_st5)
        d_p5 = dev(E.derive_pk_g2_batch(sks5, ctx=ctx))
        agg = D.DistinctAggregate(ctx)
        agg.step(d_m5, 32, d_p5, d_s5, n5)
        v5 = agg.status()
        ms5 = timed(lambda: agg.step(d_m5, 32, d_p5, d_s5, n5), reps) / reps
        # one forged signature on the last rank: every rank must reject
        if rank == world - 1:
            d_s5[64 * 5:64 * 6] = d_s5[64 * 6:64 * 7].clone()
        agg.step(d_m5, 32, d_p5, d_s5, n5)
        v5_bad = agg.status()
        configs["config5_distinct_messages"] = {
            "pairs_total": n5 * world, "pairs_per_sec": n5 * world / (ms5 * 1e-3), "ms_per_step": ms5, "verdict": v5,
            "verdict_with_forged_signature_on_last_rank": v5_bad, "scaling": "strong", "exchange_bytes_per_rank": 448,
            "roofline_frac_whole_step": None,  # filled on rank 0 once the IMAD peak is known
            "path": "bn254_distinct_payload_dev -> all_gather_into_tensor (device) -> bn254_finish_distinct_dev (cooperative final exponentiation)"}
        del d_m5, d_k5, d_s5, d_p5, d_st5

        # small-batch latency of the crate's own call shape: host buffers in, verdicts out, one call at a time
        lat = {}
        for nl in (1, 2, 4, 32, 1024, 4736, 1 << 14):
            if nl > n:
                continue
            hm, hs_, hp = h_msgs[:32 * nl], h_sigs[:64 * nl], h_pks[:128 * nl]
            call = lambda: ctx.call("bn254_verify_batch", hm, S(32), hs_, hp, S(nl), h_st[:nl])
            call()
            ts = []
            for _ in range(7):
                t0 = time.perf_counter()
                call()
                ts.append((time.perf_counter() - t0) * 1e3)
            assert nocheck or bytes(h_st[:nl].numpy().tobytes()) == bytes(nl)
            lat[str(nl)] = {"ms_median": sorted(ts)[len(ts) // 2], "ms_min": min(ts)}
        configs["latency_verify_host_buffers"] = lat

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    cal = imad_peak()
    peak = cal.get("imad_lo", {}).get("gops") if isinstance(cal, dict) else None
    # phase[0] hash, phase[1] line sets, phase[2] cooperative Miller + final exponentiation (one profiled step)
    coop_ms = phase[2]
    achieved = (n * M_COOP * IMAD_PER_M / (coop_ms * 1e-3) / 1e9) if coop_ms > 0 else None
    step_ms = ms_total / args.steps
    traffic, hbm, traffic_src = None, None, None
    try:  # DRAM bytes of one k_coop4_run launch from the committed ncu capture (NOT measured in this run)
        for cand in ("r02_traffic.json", "r01_traffic.json"):
            if os.path.exists(os.path.join(ROOT, "profiles", cand)):
                traffic_src = "profiles/" + cand
                break
        tj = json.load(open(os.path.join(ROOT, traffic_src)))
        per_item = tj["k_coop4_run"]["dram_bytes_per_launch"] / tj["k_coop4_run"]["items_per_launch"]
        chunk = 1 << int(os.environ.get("BN254_COOP_CHUNK_LOG2", "19"))
        launches_per_step = max(1, (n + chunk - 1) // chunk)
        traffic = per_item * min(n, chunk)
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
        hbm = {"achieved_gbs": per_item * n / (coop_ms * 1e-3) / 1e9, "peak_gbs": peaks.get("hbm_gbs", 6650.0),
               "peak_source": "MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback 6.65 TB/s", "launches_per_step": launches_per_step}
        hbm["frac"] = hbm["achieved_gbs"] / hbm["peak_gbs"]
    except Exception:
        pass
    roof = {
        "bound": "int32_imad", "kernel": "k_coop4_run", "achieved": achieved, "peak": peak, "unit": "GIMAD32/s",
        "frac": (achieved / peak if achieved and peak else None), "traffic": traffic,
        "traffic_source": ("committed ncu capture %s scaled to this launch's item count; not measured in this run" % traffic_src) if traffic else None,
        "peak_source": "tools/microbench.bin mad.lo.u32 chain measured in this run (MEASURED_PEAKS.json has no int32 figure)",
        "algorithmic_per_unit": {"hash kernels (k_hash_round / k_hash_tail)": M_HASH * IMAD_PER_M, "k_verify_lines": M_LINES * IMAD_PER_M,
                                 "k_coop4_run": M_COOP * IMAD_PER_M, "unit": "IMAD32 per verify"},
        "whole_step_frac": (n * M_VERIFY * IMAD_PER_M / (step_ms * 1e-3) / 1e9 / peak) if peak else None,
        "phase_ms": {"hash_to_g1": phase[0], "line_sets": phase[1], "miller_and_final_exp": phase[2]},
        "phase_frac": ({"hash_to_g1": n * M_HASH * IMAD_PER_M / (phase[0] * 1e-3) / 1e9 / peak,
                        "line_sets": n * M_LINES * IMAD_PER_M / (phase[1] * 1e-3) / 1e9 / peak} if peak and phase[0] > 0 and phase[1] > 0 else None),
        "hbm": hbm, "calibration": cal,
    }
    if configs and peak:
        c5 = configs["config5_distinct_messages"]
        c5["roofline_frac_whole_step"] = c5["pairs_per_sec"] / world * M_DISTINCT_PAIR * IMAD_PER_M / 1e9 / peak
        configs["pairings"]["roofline_frac_whole_step"] = configs["pairings"]["pairings_per_sec"] / world * M_PAIRING * IMAD_PER_M / 1e9 / peak
        configs["config3_hash_sign"]["roofline_frac_whole_step"] = configs["config3_hash_sign"]["signs_per_sec"] / world * M_SIGN * IMAD_PER_M / 1e9 / peak
    threads = os.cpu_count() or 1
    cpu_line = None
    if world == 1:  # the CPU baseline is reported by the single-GPU run only
        n_cpu = args.cpu_sample or min(n, max(256, 5120 * threads))  # bounded sample: ~10 s of CPU work
        cpu_v, cpu_dt, cpu_st = cpu_baseline(n_cpu, msgs, sigs, pks, threads)
        assert cpu_st == bytes(n_cpu)
        cpu_line = {"value": cpu_v, "per_core": cpu_v / threads, "unit": UNIT, "cores": threads, "kind": "port",
                    "sample": "first %d triples of the workload, oracle/bn254_oracle.c on %d threads, %.1f s" % (n_cpu, threads, cpu_dt)}
        if configs:
            one = cpu_single_thread_ms(msgs, sigs, pks)
            lat = configs["latency_verify_host_buffers"]
            lat["cpu_port_one_core_ms_per_verify"] = one
            # the batch size from which one GPU call beats one CPU core doing the items one after the other
            lat["crossover_items_vs_one_cpu_core"] = next((int(k) for k in ("1", "2", "4", "32", "1024", "4736", "16384") if k in lat and lat[k]["ms_median"] < int(k) * one), None)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32",
        "data": "synthetic",
        "config": shared_config(n),
        "engine": {"pairing_kernels": "cooperative machine: six warps per 32-item group, four groups per block (one per SM sub-partition), workspace chunks of 2^19 items",
                   "small_batches": "counter-parallel hash, four-warp cooperative G2 walk, eighteen-warp machine blocks, producer and machine pipelined (latency_verify_host_buffers)"},
        "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": 224 * n, "d2h_bytes_per_step": n, "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": launches, "clocks": sampler.result(), "roofline": roof,
        "cpu_baseline": cpu_line, "configs": configs,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_engine(a)
