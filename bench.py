#!/usr/bin/env python
"""bench.py -- chain-steps/second of the jump_mala + RealNVP hot path on N B200s (one process per GPU).

Workload (BASELINE.json metric, SURVEY.md section 8 config CT): jump_mala, standard Gaussian target
U = sum x^2, d = 100, RealNVP Lc = 2 with the default conditioner (M = 2, H = 5), frozen flow perturbed by
0.1*randn, K = 100 MALA steps + 1 NF jump per outer iteration, 2^20 chains PER GPU (weak scaling; the
chain state, 419 MB, is larger than L2), Philox noise keyed by global chain index, store_samples = False.
One bench "step" = one outer iteration = (K+1)*n chain-steps per GPU.

    python bench.py --gpus 1 --steps 20 --warmup 3              # our arm
    python bench.py --impl reference --steps 3 --warmup 1       # the reference algorithm on the host cores

Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=20)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="native", choices=["native", "reference"])
    p.add_argument("--workload", default="jump_mala", choices=["jump_mala", "jump_hmc"])
    p.add_argument("--dim", type=int, default=100)
    p.add_argument("--chains-per-gpu", type=int, default=1 << 20)
    p.add_argument("--inner", type=int, default=None, help="local steps per jump (default 100 for jump_mala, 5 for jump_hmc)")
    p.add_argument("--leapfrog", type=int, default=20)
    p.add_argument("--cpu-chains", type=int, default=32768, help="chains of the bounded CPU-baseline sample")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-e2e", action="store_true")
    return p.parse_args()


# ---------------------------------------------------------------------------------------------------------------
def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            parts = [s.strip() for s in line.split(",")]
            if len(parts) >= 6:
                self.rows.append(parts)

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        mx = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.rows)}


def workload_params(args):
    d = args.dim
    if args.workload == "jump_mala":
        K = args.inner or 100
        return dict(kind=0, K=K, step=d ** (-1 / 3), L=0, pot="g0")
    K = args.inner or 5
    return dict(kind=1, K=K, step=0.01, L=args.leapfrog, pot="g1")


def make_oracle_flow(d):
    from oracle.realnvp_ref import make_flow
    return make_flow((d,), n_layers=2, perturb=0.1, seed=0)


# ---------------------------------------------------------------------------------------------------------------
# CPU: the reference algorithm (oracle port of jump.py:156-246 + langevin.py / hmc.py) on the host cores
# ---------------------------------------------------------------------------------------------------------------
def cpu_run(args, n, n_outer):
    from oracle import samplers_ref as R
    from oracle.potentials_ref import make_potential_ref
    w = workload_params(args)
    d = args.dim
    torch.manual_seed(0)
    flow = make_oracle_flow(d)
    target = make_potential_ref(w["pot"], (d,))
    x0 = torch.randn(n, d)
    t0 = time.perf_counter()
    run = R.run_jump(x0, target, flow, "mala" if w["kind"] == 0 else "hmc", n_outer, w["K"], R.GlobalDraws(), w["step"],
                     torch.ones(d), n_leapfrog=max(w["L"], 1), store=True)
    dt = time.perf_counter() - t0
    steps = run.samples.shape[0] * n
    return steps / dt, dt


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count() or 1)
    w = workload_params(args)
    n = args.cpu_chains
    for _ in range(args.warmup):
        cpu_run(args, max(256, n // 8), 1)
    t_all, steps_all = 0.0, 0
    for _ in range(args.steps):
        rate, dt = cpu_run(args, n, 1)
        t_all += dt
        steps_all += (w["K"] + 1) * n
    value = steps_all / t_all
    line = {
        "impl": "reference", "metric": f"{args.workload} chain-steps/sec", "value": value, "unit": "chain-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_all / max(args.steps, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": bench_config(args, w, note="CPU arm: oracle port of the reference algorithm on the host cores"),
        "cpu_baseline": {"value": value, "unit": "chain-steps/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": f"{n} chains x 1 outer iteration ({w['K']} local steps + 1 jump) per step, "
                                   f"store_samples=True as the reference requires (jump.py:163-164)"},
        "e2e": {"value": value, "unit": "chain-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def bench_config(args, w, note=None):
    cfg = {"workload": f"{args.workload}+realnvp d={args.dim} Lc=2 M=2 H=default, K={w['K']} local steps + 1 jump per step, "
                       f"{args.chains_per_gpu} chains per GPU, potential {w['pot']}, frozen flow, Philox noise, store_samples=False",
           "chains_per_gpu": args.chains_per_gpu, "dim": args.dim, "inner_steps": w["K"],
           "l2": "inputs larger than L2 (chain state 4*n*d bytes per GPU)" if args.chains_per_gpu * args.dim * 4 > 126e6
           else "L2 flushed between timed steps (256 MiB write)"}
    if w["kind"] == 1:
        cfg["leapfrog"] = w["L"]
    if note:
        cfg["note"] = note
    return cfg


# ---------------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------------
def native_arm(args):
    import torch.distributed as dist
    from nfmc_b200 import _native as N
    from nfmc_b200.flow import Flow, RealNVP
    from nfmc_b200.potentials import make_potential

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = N.lib()
    w = workload_params(args)
    d, n, K = args.dim, args.chains_per_gpu, w["K"]
    chain0 = rank * n
    seed = 20261018

    oflow = make_oracle_flow(d)
    flow = Flow(RealNVP((d,), n_layers=2))
    flow.load_state_dict(oflow.state_dict())
    flow = flow.to(dev)
    pot = make_potential(w["pot"], (d,))
    pd, keep_p = pot.descriptor(dev)
    fd, keep_f = flow.bijection.descriptor(dev)

    gen = torch.Generator(device="cpu").manual_seed(rank)
    x_host = torch.randn(n, d, generator=gen).pin_memory()
    x = x_host.to(dev, non_blocking=True)
    moments = torch.zeros(2 * d, device=dev, dtype=torch.float64)
    counts = torch.zeros(8, device=dev, dtype=torch.int64)
    st_local = N.StatsDesc(moments.data_ptr(), moments.data_ptr() + 8 * d, counts.data_ptr())
    st_jump = N.StatsDesc(moments.data_ptr(), moments.data_ptr() + 8 * d, counts.data_ptr() + 32)
    stream = N.stream_ptr(dev)
    logq_scratch = torch.empty(n, device=dev, dtype=torch.float32)     # log q(x) between the two kernels of the NF jump
    flush = None
    if n * d * 4 <= 126e6:
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def local_launch(it):
        rng = N.rng_desc(seed, it * K, None, None)
        if w["kind"] == 0:
            N.check(lib.nfmc_mala_steps(C.byref(pd), N.ptr(x), n, K, float(w["step"]), None, 1, C.byref(rng), chain0,
                                        C.byref(st_local), None, stream))
        else:
            N.check(lib.nfmc_hmc_steps(C.byref(pd), N.ptr(x), n, K, float(w["step"]), w["L"], None, 1, C.byref(rng), chain0,
                                       C.byref(st_local), None, stream))

    def jump_launch(it):
        rng = N.rng_desc(seed, it, None, None)
        N.check(lib.nfmc_jump_step2(C.byref(pd), C.byref(fd), N.ptr(x), N.ptr(logq_scratch), n, 1, C.byref(rng), chain0,
                                    C.byref(st_jump), None, stream))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def step_launch(it):
        """One step of the hot path: K local steps + 1 NF jump for every chain, through the whole-run entry point (chains
        cut into slabs spread over several streams so that kernels of different slabs overlap)."""
        N.check(lib.nfmc_jump_sample_device(C.byref(pd), C.byref(fd), N.ptr(x), n, w["kind"], 1, K, float(w["step"]), w["L"],
                                            None, 1, 1, seed, it * K, it, chain0, C.byref(st_local), C.byref(st_jump), N.ptr(logq_scratch),
                                            stream))

    it = 0
    for _ in range(args.warmup):
        step_launch(it)
        it += 1
    barrier()

    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(args.steps)]
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    barrier()
    for s in range(args.steps):
        if flush is not None:
            flush.fill_(s & 0xFF)
        ev[s][0].record()
        step_launch(it)
        ev[s][1].record()
        it += 1
    barrier()
    clk = clocks.stop() if rank == 0 else None
    t_total = sum(e[0].elapsed_time(e[1]) for e in ev) * 1e-3

    # ---- roofline pass: inside the timed region the launches of different slabs overlap, so a per-launch duration is not
    #      defined there; the dominant kernel is timed here, right after it, alone on the stream, on the same buffers:
    #      one full-batch launch of K local steps per step, then the jump launch --------------------------------------
    r_steps = max(3, min(args.steps, 10))
    evr = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(r_steps)]
    for s in range(r_steps):
        if flush is not None:
            flush.fill_(s & 0xFF)
        evr[s][0].record()
        local_launch(it)
        evr[s][1].record()
        jump_launch(it)
        evr[s][2].record()
        it += 1
    barrier()
    t_local = sum(e[0].elapsed_time(e[1]) for e in evr) * 1e-3 / r_steps * args.steps     # per-launch time x steps
    t_serial = sum(e[0].elapsed_time(e[2]) for e in evr) * 1e-3 / r_steps * args.steps
    tt = torch.tensor([t_total, t_local, t_serial], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t_total, t_local, t_serial = float(tt[0]), float(tt[1]), float(tt[2])
    chain_steps = (K + 1) * n * args.steps * world
    value = chain_steps / t_total

    # pooled statistics: the only collective of the path (moments + counters, once per run)
    if world > 1:
        dist.all_reduce(moments)
        dist.all_reduce(counts)
    acc = counts.cpu().tolist()

    # ---- end to end through the C ABI with host buffers (H2D + run + D2H inside the timed region) -------------
    e2e = None
    if not args.no_e2e:
        blob_host = flow.bijection.blob(dev).cpu().pin_memory()
        ws_bytes = lib.nfmc_jump_workspace_bytes(d, n, blob_host.numel())
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        sx = torch.zeros(d, dtype=torch.float64).pin_memory()
        sx2 = torch.zeros(d, dtype=torch.float64).pin_memory()
        cnt = torch.zeros(8, dtype=torch.int64).pin_memory()
        hp = pot.host_params()
        hp_ptr, hp_n = (None, 0) if hp is None else (hp.contiguous().data_ptr(), hp.numel())
        pdesc_host = N.PotentialDesc(pot.kind, d, None, (C.c_float * 4)(*list(pot.scalars())[:4]))
        fdesc_host = N.RealNVPDesc(d, 2, fd.n_linear, fd.hidden, None, blob_host.numel())

        def e2e_call(outer):
            N.check(lib.nfmc_jump_sample_host(C.byref(pdesc_host), hp_ptr, hp_n, C.byref(fdesc_host), blob_host.data_ptr(),
                                              x_host.data_ptr(), n, w["kind"], outer, K, float(w["step"]), w["L"], seed, chain0,
                                              sx.data_ptr(), sx2.data_ptr(), cnt.data_ptr(), ws.data_ptr(), ws_bytes, stream))

        e2e_call(1)
        barrier()
        e_steps = max(3, min(args.steps, 10))
        t0 = time.perf_counter()
        for _ in range(e_steps):
            e2e_call(1)
        barrier()
        te = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e = {"value": (K + 1) * n * e_steps * world / float(te[0]), "unit": "chain-steps/s",
               "h2d_bytes_per_step": n * d * 4 + blob_host.numel() * 4 + hp_n * 4,
               "d2h_bytes_per_step": n * d * 4 + 2 * d * 8 + 64,
               "call": "nfmc_jump_sample_host (pinned host x0 in, final state + pooled moments + counters out)"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peaks()
    bytes_per_launch = 8.0 * d * n * K               # 8*d bytes per chain-step, K*n chain-steps per launch
    t_launch = t_local / args.steps
    achieved = bytes_per_launch / t_launch / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get(f"{args.workload}_d{d}")
        except Exception:
            traffic = None
    line = {
        "metric": f"{args.workload} chain-steps/sec", "value": value, "unit": "chain-steps/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_total / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": bench_config(args, w),
        "roofline": {"bound": "hbm", "kernel": "mala_kernel" if w["kind"] == 0 else "hmc_kernel",
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                     "peak_source": peak_src, "kernel_ms_per_launch": 1e3 * t_launch,
                     "kernel_share_of_step": t_local / t_serial,
                     "serial_ms_per_step": 1e3 * t_serial / args.steps,
                     "timing": "CUDA events around single-stream full-batch launches of the kernel, run right after the timed "
                               "region on the same buffers (inside the timed region launches of different slabs overlap)",
                     "note": "algorithmic bytes = 8*d per chain-step (state read+write, SURVEY 8d) x K*n chain-steps per launch; "
                             "the kernel keeps the state on chip for all K steps, so real DRAM traffic is ~8*d*n per launch and the "
                             "kernel is bound by fp32/integer issue (Philox + Box-Muller), see DESIGN.md"},
        "e2e": e2e,
        "gpu_launches": 3 * int(lib.nfmc_jump_sample_slabs(d, n, 0)) * args.steps,
        "clocks": clk,
        "acceptance": {"local": acc[0] / max(acc[1], 1), "jump": acc[4] / max(acc[5], 1)},
    }
    if not args.no_cpu_baseline and world == 1:
        torch.set_num_threads(os.cpu_count() or 1)
        cpu_run(args, 512, 1)
        reps, t_cpu = 4, 0.0                              # ~10 s of CPU work: 4 outer iterations of the bounded sample
        for _ in range(reps):
            t_cpu += cpu_run(args, args.cpu_chains, 1)[1]
        rate = reps * (K + 1) * args.cpu_chains / t_cpu
        line["cpu_baseline"] = {"value": rate, "unit": "chain-steps/s", "cores": torch.get_num_threads(), "kind": "port",
                                "sample": f"{args.cpu_chains} chains x {reps} outer iterations ({K} local steps + 1 jump each), "
                                          f"{t_cpu:.1f} s"}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        reference_arm(args)
    else:
        native_arm(args)


if __name__ == "__main__":
    main()
