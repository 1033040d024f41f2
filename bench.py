#!/usr/bin/env python
"""bench.py -- bootstraps/s (functional_bootstrap + tlwe_keyswitch per ciphertext) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload level1|level2] [--batch B]
    python bench.py --impl reference ...      # the reference's CPU implementation on the host cores

One step = one pass of the hot path over one batch of synthetic ciphertexts (per GPU: `--batch`,
default 4096 = BASELINE configs[1]).  For N > 1 launch with torchrun; ciphertexts are sharded across
ranks (weak scaling: per-GPU batch fixed), keys are made on rank 0 and broadcast once over NCCL.
Rank 0 prints ONE JSON line.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "bootstraps/sec (PBS+KS)"
UNIT = "bootstraps/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="level1", choices=["level1", "level2", "set1"])
    ap.add_argument("--batch", type=int, default=4096, help="ciphertexts per GPU per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target CPU work for the cpu_baseline sample")
    ap.add_argument("--no-extras", action="store_true", help="skip parity / e2e_handles / extra.* (rank 0, N = 1 only)")
    ap.add_argument("--strong-batch", type=int, default=65536, help="configs[2]: global Level-2 batch split over the ranks (N > 1)")
    return ap.parse_args()


def workload_config(args, P, world):
    return {
        "workload": f"batched functional_bootstrap+tlwe_keyswitch, {args.workload} params "
                    f"(n={P.n}, N={P.N}, k={P.k}, l={P.l}, Bg_bit={P.Bg_bit}, t={P.t}, base_bit={P.base_bit}), "
                    f"batch {args.batch} per GPU",
        "baseline_config": "configs[1]" if args.workload == "level1" else ("configs[2]" if args.workload == "level2" else "SET_1"),
        "batch_per_gpu": args.batch,
        "global_batch": args.batch * world,
        "torus_base": 4,
        "parallelism": f"ciphertext-sharded x{world}, keys broadcast once",
        "l2_policy": "512 MiB scratch buffer rewritten between timed steps (L2 flush); keys+batch also exceed L2",
    }


# ------------------------------------------------------------------------------------------------
# clocks sampling (B200_PROFILING.md recipe), during the timed region
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); power.append(float(r[3]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
# CPU baseline: the unmodified reference on the host cores (oracle/_ref; checker side only)
# ------------------------------------------------------------------------------------------------
def _run_ref_exe(exe, P, threads, ops_per_thread, steps, warmup):
    cmd = [exe] + [str(x) for x in (P.n, P.N, P.k, P.l, P.Bg_bit, P.t, P.base_bit, P.lwe_sigma, P.rlwe_sigma,
                                    threads, ops_per_thread, steps, warmup)]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=1500)
    if out.returncode not in (0, 1) or not out.stdout.strip():
        return {"error": (out.stderr or "no output")[-300:]}
    return json.loads(out.stdout.strip().splitlines()[-1])


def run_reference_cpu(P, threads, ops_per_thread, steps, warmup, both=False):
    """Times the UNMODIFIED reference (oracle/_ref) on the host cores.  Two builds of the same sources exist for the
    AVX-512 ISA level: the shared library + driver (`avx512`) and the reference's own benchmark recipe, one whole-program
    LTO binary (Makefile.def:2-6, `avx512_lto`); the faster one is the baseline, `both` also reports the other."""
    from oracle import ref as reflib
    variant = reflib.best_variant()
    if variant is None:
        return None
    names = [variant]
    if variant == "avx512" and os.path.exists(os.path.join(reflib.REF_DIR, "ref_bench_avx512_lto")):
        names = ["avx512_lto", "avx512"] if both else ["avx512_lto"]
    results = {}
    for nm in names:
        exe = os.path.join(reflib.REF_DIR, f"ref_bench_{nm}")
        if os.path.exists(exe):
            results[nm] = _run_ref_exe(exe, P, threads, ops_per_thread, steps, warmup)
    good = {k: v for k, v in results.items() if v and "error" not in v}
    if not good:
        return {"error": str(results), "variant": variant}
    best = max(good, key=lambda k: good[k]["pbs_ks_per_s"])
    res = dict(good[best])
    res["variant"] = best
    res["all_variants"] = {k: v["pbs_ks_per_s"] for k, v in good.items()}
    return res


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def reference_arm(args, P, rank, world):
    """`--impl reference`: rank 0 times the reference CPU implementation with all host threads."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    per_op_ms = {"level1": 25.0, "level2": 60.0, "set1": 13.0}[args.workload]
    # bounded sample: each step ~ args.cpu_seconds / (steps + warmup) of wall time on all cores
    total_steps = args.steps + args.warmup
    ops = max(1, int(args.cpu_seconds * 6 / total_steps * 1e3 / per_op_ms))
    res = run_reference_cpu(P, cores, ops, args.steps, args.warmup)
    cfg = workload_config(args, P, world)
    if not res or "error" in res:
        print(json.dumps({"impl": "reference", "unavailable": f"oracle/_ref driver failed: {res}"}))
        return
    sample = f"{res['ops_per_step']} PBS+KS per step ({ops} per thread x {cores} threads), same parameters, variant {res['variant']}"
    v = res["pbs_ks_per_s"]
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64+u64", "data": "synthetic", "config": cfg,
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "reference", "sample": sample,
                             "cpu": cpu_model(), "wrong_results": res["wrong"]},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# rank 0, N = 1: the drop-in handle path on reference-made keys (e2e_handles), the parity gate, the other configs
# ------------------------------------------------------------------------------------------------
def handle_path_and_parity(args, P, api, B, torus_base, value):
    """`parity`: the CUDA path against the unmodified reference on the SAME keys at the benchmarked parameters
    (oracle/parity.py: messages identical, bootstrap phase difference beside the reference-vs-reference yardstick, key switch
    bit-exact).  `e2e_handles`: the metric through functional_bootstrap_keyswitch_batch over arrays of the reference's own
    TLWE handles (host gather / H2D / kernels / D2H / host scatter inside the timed region)."""
    out = {"parity": None, "e2e_handles": None}
    try:
        import ctypes as C
        from mosfhet_b200 import abi
        from oracle import parity, ref as reflib
        if not reflib.available():
            out["parity"] = {"unavailable": "oracle/_ref not built"}
            return out
        out["parity"] = parity.reference_parity(P, api, n_inputs=64, ks_count=640, policies=((0, "auto"), (5, "k1q")))
        S = parity.ReferenceSetup(P, B, torus_base)
        R = S.R
        api.set_host_fft_layout(R.layout)
        api.register_bootstrap_key(S.bk)
        api.register_ks_key(S.ksk)
        outs = [R.tlwe_alloc_sample(P.n) for _ in range(B)]
        a_out, a_in = abi.handle_array(outs, abi.TLWE), abi.handle_array(S.inputs, abi.TLWE)
        a_tv = abi.handle_array([S.tv], abi.TRLWE)
        fn = api.lib().functional_bootstrap_keyswitch_batch

        def call():
            fn(a_out, a_tv, 1, a_in, S.bk, S.ksk, torus_base, B)
        for _ in range(2):
            call()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            call()
        dt = time.perf_counter() - t0
        dec = S.decode(S.phases(outs, S.key_tlwe))
        wrong = int((dec % (2 * torus_base) != (3 * S.msgs + 1) % torus_base).sum())
        v = B * args.steps / dt
        out["e2e_handles"] = {"value": v, "unit": UNIT, "ms_per_step": dt / args.steps * 1e3, "wrong_results": wrong,
                              "fraction_of_device_value": v / value,
                              "path": "functional_bootstrap_keyswitch_batch(TLWE*, ...) on keys and inputs made by the reference library",
                              "h2d_bytes_per_step": B * (P.n + 1) * 8 + (P.k + 1) * P.N * 8, "d2h_bytes_per_step": B * (P.n + 1) * 8}
        api.release_bootstrap_key(S.bk)
        api.release_ks_key(S.ksk)
    except Exception as e:                                 # the headline numbers must survive a checker-side failure
        out["handle_path_error"] = repr(e)[:300]
    return out


def run_extras(args, api, bsk, ksk):
    extra = {}
    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    try:
        import bench_extras
        bsk.free(); ksk.free()                             # make room: Level-2 keys are 1.4 GB
        extra["level2_batch1"] = bench_extras.level2_batch1(api)
        extra["next"] = bench_extras.next_configs(api)
    except Exception as e:
        extra["error"] = repr(e)[:300]
    return extra


def strong_scaling_level2(args, api, sharding, syn, device, rank, world, dist, torch):
    """configs[2]: programmable bootstrap + key switch at the default Level-2 parameters, `--strong-batch` ciphertexts in
    total, sharded contiguously over the ranks (keys made on rank 0, broadcast once over NCCL)."""
    from mosfhet_b200.params import LEVEL2 as P2
    total = args.strong_batch
    lo, hi = sharding.my_shard(total, rank, world)
    Bs = hi - lo
    lwe_key, rlwe_key = syn.binary_key(P2.n, 2001), syn.binary_key(P2.k * P2.N, 2002)
    bsk = ksk = None
    if rank == 0:
        bsk = api.BootstrapKey.synthesize(P2, lwe_key, rlwe_key, seed=21)
        ksk = api.KeySwitchKey.synthesize(P2, rlwe_key, lwe_key, seed=22)
    bsk, ksk = sharding.broadcast_keys(P2, bsk, ksk, device)
    msgs = (syn.splitmix64_stream(5, total) & np.uint64(3)).astype(np.int64)[lo:hi]
    cts = syn.tlwe_encrypt(syn.encode(msgs, 4), lwe_key, P2.lwe_sigma, seed=300 + rank)
    lut = syn.encode((3 * np.arange(4) + 1) % 4, 4)
    d_in = torch.from_numpy(cts.view(np.int64)).to(device)
    d_tv = torch.from_numpy(syn.test_vector(lut, P2.N, P2.k).view(np.int64)).to(device)
    d_mid = torch.empty((Bs, P2.k * P2.N + 1), dtype=torch.int64, device=device)
    d_out = torch.empty((Bs, P2.n + 1), dtype=torch.int64, device=device)
    stream = torch.cuda.Stream(device)
    times = []
    for it in range(3):                                    # 1 warm-up + 2 timed passes
        dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record(stream)
            api.pbs_dev(bsk, d_mid, d_tv, 1, d_in, 4, Bs, stream.cuda_stream)
            api.ks_dev(ksk, d_out, d_mid, Bs, stream.cuda_stream)
            e1.record(stream)
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    out_np = d_out.cpu().numpy().view(np.uint64)
    dec = ((syn.tlwe_phase(out_np, lwe_key) + (np.uint64(1) << np.uint64(60))) >> np.uint64(61)).astype(np.int64) % 8
    wrong = int((dec != (3 * msgs + 1) % 4).sum())
    t = torch.tensor([min(times[1:]), float(wrong)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t[0])
    return {"config": "configs[2]: Level 2 (n=632, N=2048, l=4, Bg_bit=9, t=8, base_bit=4), global batch split over the ranks",
            "global_batch": total, "per_rank": Bs, "n_gpus": world, "scaling": "strong", "ms_per_step": ms,
            "value": total / ms * 1e3, "unit": UNIT, "wrong_results_max_over_ranks": int(t[1]),
            "timing": "CUDA events on the launching stream, device-resident inputs, best of 2 after a warm-up, max over ranks"}


# ------------------------------------------------------------------------------------------------
def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    from mosfhet_b200.params import NAMED
    P = NAMED[args.workload]

    if args.impl == "reference":
        reference_arm(args, P, rank, world)
        return

    import torch
    import torch.distributed as dist

    from mosfhet_b200 import api, sharding, synthetic as syn

    api.require_gpu()
    torch.cuda.set_device(local_rank)
    api.init(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)

    # ---- keys: made on rank 0's GPU from seeded binary secrets, broadcast once ------------------
    lwe_key = syn.binary_key(P.n, 1001)
    rlwe_key = syn.binary_key(P.k * P.N, 1002)
    t0 = time.perf_counter()
    bsk = ksk = None
    if rank == 0:
        bsk = api.BootstrapKey.synthesize(P, lwe_key, rlwe_key, seed=11)
        ksk = api.KeySwitchKey.synthesize(P, rlwe_key, lwe_key, seed=12)
    bsk, ksk = sharding.broadcast_keys(P, bsk, ksk, device)
    torch.cuda.synchronize()
    key_setup_s = time.perf_counter() - t0

    # ---- synthetic plaintexts (splitmix64 seed 1), encrypted on the host ------------
    B, torus_base = args.batch, 4
    msgs = (syn.splitmix64_stream(1 + rank, B) & np.uint64(3)).astype(np.int64)
    cts = syn.tlwe_encrypt(syn.encode(msgs, torus_base), lwe_key, P.lwe_sigma, seed=77 + rank)
    lut = syn.encode((3 * np.arange(torus_base) + 1) % torus_base, torus_base)
    tv = syn.test_vector(lut, P.N, P.k)
    h_in = torch.from_numpy(cts.view(np.int64)).pin_memory()
    h_tv = torch.from_numpy(tv.view(np.int64)).pin_memory()
    h_out = torch.empty((B, P.n + 1), dtype=torch.int64).pin_memory()
    d_in, d_tv = h_in.to(device), h_tv.to(device)
    d_mid = torch.empty((B, P.k * P.N + 1), dtype=torch.int64, device=device)
    d_out = torch.empty((B, P.n + 1), dtype=torch.int64, device=device)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=device)
    stream = torch.cuda.Stream(device)
    sptr = stream.cuda_stream

    def step_device(ev=None):
        """one pass, inputs resident in HBM; events recorded on the launching stream"""
        with torch.cuda.stream(stream):
            if ev:
                ev[0].record(stream)
            api.pbs_dev(bsk, d_mid, d_tv, 1, d_in, torus_base, B, sptr)
            if ev:
                ev[1].record(stream)
            api.ks_dev(ksk, d_out, d_mid, B, sptr)
            if ev:
                ev[2].record(stream)

    def barrier():
        if world > 1:
            dist.barrier()

    # ---- warm-up ------------------------------------------------------------------------------------
    for _ in range(args.warmup):
        flush.fill_(1)
        step_device()
    torch.cuda.synchronize()

    # correctness of what is being timed: every output of this rank decrypts to LUT[m]
    out_np = d_out.cpu().numpy().view(np.uint64)
    dec = ((syn.tlwe_phase(out_np, lwe_key) + (np.uint64(1) << np.uint64(60))) >> np.uint64(61)).astype(np.int64) % 8
    wrong = int((dec != (3 * msgs + 1) % torus_base).sum())

    # ---- timed region: K steps, device-resident --------------------------------------------------------
    sampler = ClockSampler(local_rank)
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    barrier(); torch.cuda.synchronize()
    sampler.start()
    api.reset_launch_count()
    for s in range(args.steps):
        flush.fill_(s & 0xFF)                      # L2 flush between timed iterations (not inside the events)
        step_device(evs[s])
    torch.cuda.synchronize(); barrier()
    launches = api.launch_count()
    pbs_ms = [evs[s][0].elapsed_time(evs[s][1]) for s in range(args.steps)]
    ks_ms = [evs[s][1].elapsed_time(evs[s][2]) for s in range(args.steps)]
    step_ms = [evs[s][0].elapsed_time(evs[s][2]) for s in range(args.steps)]
    total_ms = float(sum(step_ms))

    # ---- e2e: host buffers through the C-ABI batch call (H2D + kernels + D2H inside the timed region) ---
    for _ in range(2):
        api.pbs_ks_host(bsk, ksk, h_tv, h_in, torus_base, out=h_out)
    barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for s in range(args.steps):
        api.pbs_ks_host(bsk, ksk, h_tv, h_in, torus_base, out=h_out)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    barrier()
    clocks = sampler.stop()
    e2e_ok = bool(np.array_equal(h_out.numpy().view(np.uint64), out_np))

    # ---- max over ranks -----------------------------------------------------------------------------------
    if world > 1:
        t = torch.tensor([total_ms, e2e_s, float(wrong), float(sum(pbs_ms)), float(sum(ks_ms))], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, e2e_s, wrong = float(t[0]), float(t[1]), int(t[2])
        pbs_total, ks_total = float(t[3]), float(t[4])
    else:
        pbs_total, ks_total = float(sum(pbs_ms)), float(sum(ks_ms))

    if rank == 0:
        value = B * world * args.steps / (total_ms * 1e-3)
        e2e_v = B * world * args.steps / e2e_s
        # roofline of the dominant kernel (blind rotation): algorithmic FP64 flops per launch
        flops_launch = P.flops_per_pbs() * B
        pbs_avg_s = pbs_total / args.steps * 1e-3
        fp64_peak = api.measure_fp64_tflops()
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        achieved_tf = flops_launch / pbs_avg_s * 1e-12
        # DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture of this
        # workload (profiles/ncu_traffic.json, written by scripts/ncu_summary.py); null if not captured
        traffic = None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
            ent = tj.get(f"{args.workload}:{B}")
            if ent:
                traffic = ent["dram_bytes_per_launch"]
        except (OSError, ValueError, KeyError):
            pass
        roofline = {"bound": "fp64", "kernel": api.last_blind_rotate_kernel(), "achieved": achieved_tf, "peak": fp64_peak,
                    "unit": "TFLOP/s", "frac": achieved_tf / fp64_peak if fp64_peak else None, "traffic": traffic,
                    "peak_source": "measured live: mb200_measure_fp64_tflops (dependent-FMA microbenchmark, same clocks)",
                    "algorithmic_flops_per_unit": P.flops_per_pbs(), "units_per_launch": B,
                    "kernel_share_of_step": pbs_total / total_ms,
                    "hbm": {"achieved": P.bsk_bytes / pbs_avg_s * 1e-9, "peak": hbm_peak, "unit": "GB/s",
                            "note": "bootstrapping-key bytes streamed once per launch (ideal reuse W = batch) over the launch time",
                            "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback"},
                    "keyswitch": {"ms_per_launch": ks_total / args.steps,
                                  "achieved": P.ksk_bytes / (ks_total / args.steps * 1e-3) * 1e-9, "unit": "GB/s",
                                  "note": "KSK bytes swept once per launch"}}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64+u64", "data": "synthetic", "config": workload_config(args, P, world),
                "roofline": roofline,
                "e2e": {"value": e2e_v, "unit": UNIT, "h2d_bytes_per_step": int(h_in.numel() * 8 + h_tv.numel() * 8),
                        "d2h_bytes_per_step": int(h_out.numel() * 8), "matches_device_path": e2e_ok},
                "gpu_launches": int(launches), "clocks": clocks,
                "wrong_results": wrong, "key_setup_s": key_setup_s,
                "ms_per_bootstrap_batch1": None}
        # batch-1 latency (the metric's second half): one ciphertext through the same device path
        lat = []
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for i in range(6):
            with torch.cuda.stream(stream):
                e0.record(stream)
                api.pbs_dev(bsk, d_mid, d_tv, 1, d_in, torus_base, 1, sptr)
                api.ks_dev(ksk, d_out, d_mid, 1, sptr)
                e1.record(stream)
            torch.cuda.synchronize()
            lat.append(e0.elapsed_time(e1))
        line["ms_per_bootstrap_batch1"] = float(np.median(lat[2:]))
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            per_op_ms = {"level1": 25.0, "level2": 60.0, "set1": 13.0}[args.workload]
            ops = max(1, int(args.cpu_seconds * 1e3 / per_op_ms / 2))
            res = run_reference_cpu(P, cores, ops, 1, 1, both=True)
            if res and "error" not in res:
                line["cpu_baseline"] = {"value": res["pbs_ks_per_s"], "unit": UNIT, "cores": cores, "kind": "reference",
                                        "sample": f"{res['ops_per_step']} PBS+KS ({ops} per thread x {cores} threads), "
                                                  f"1 warm-up + 1 timed pass, same parameters, build variant {res['variant']}",
                                        "builds": res["all_variants"],
                                        "builds_note": "avx512_lto = the reference's own benchmark recipe (Makefile.def:2-6, one -flto "
                                                       "-fwhole-program binary); avx512 = shared library + driver; ISA level x86-64-v4+vaes "
                                                       "instead of -march=native",
                                        "cpu": cpu_model(), "wrong_results": res["wrong"]}
            else:
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": cores, "kind": "reference",
                                        "sample": f"unavailable: {res}"}
        if world == 1 and not args.no_extras:
            line.update(handle_path_and_parity(args, P, api, B, torus_base, value))
            line["extra"] = run_extras(args, api, bsk, ksk)
    # ---- configs[2]: Level-2 parameters, a fixed global batch split over the ranks (strong scaling) ------------------------
    strong = None
    if world > 1 and not args.no_extras:
        strong = strong_scaling_level2(args, api, sharding, syn, device, rank, world, dist, torch)
    if rank == 0:
        if strong:
            line.setdefault("extra", {})["strong_scaling_level2"] = strong
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
