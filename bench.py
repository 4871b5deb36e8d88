#!/usr/bin/env python
"""bench.py -- chain-steps/second of the jump_mala + RealNVP hot path on N B200s (one process per GPU).

Workload (BASELINE.json metric, SURVEY.md section 8 config CT): jump_mala, standard Gaussian target
U = sum x^2, d = 100, RealNVP Lc = 2 with the default conditioner (M = 2, H = 5), frozen flow perturbed by
0.1*randn, K = 100 MALA steps + 1 NF jump per outer iteration, 2^20 chains PER GPU (weak scaling; the
chain state, 419 MB, is larger than L2), Philox noise keyed by global chain index, store_samples = False.
One bench "step" = one outer iteration = (K+1)*n chain-steps per GPU.

    python bench.py --gpus 1 --steps 20 --warmup 3              # our arm
    python bench.py --impl reference --steps 3 --warmup 1       # the reference's own classes on the host cores

Prints ONE JSON line (rank 0).  Besides the contract keys the line carries:
  roofline   the dominant kernel (K local steps) against the bound that actually limits it -- fp32/int instruction issue:
             achieved = algorithmic fp32 operations / launch time against the FFMA peak measured in the same run
             (tools/microbench_fp32), frac_issue = warp-instructions issued / issue slots, and the 8*d-bytes figure of
             SURVEY 8d as frac_hbm_notional (the state stays on chip for K steps, so that one is NOT a bound)
  strong     the north star's own configuration: 2^20 chains IN TOTAL over the N ranks (value and e2e)
  shard_check (N > 1) 4096 global chains run sharded and on one rank: pooled moments and counters must agree
  e2e.link_gbs / frac_of_link: a host<->device copy ceiling measured with the same buffers
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
    p.add_argument("--no-strong", action="store_true")
    p.add_argument("--strong-chains", type=int, default=1 << 20, help="total chains of the strong-scaling block")
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


def make_bench_flow(d):
    """The benchmark's frozen RealNVP (Lc = 2, default conditioner), seed-0 initialisation perturbed by 0.1*randn so that it
    is not the identity.  Built from the product's own module classes, on the host; the CPU arms load the same
    state_dict into the reference-side flow."""
    from nfmc_b200.flow import Flow, RealNVP
    torch.manual_seed(0)
    flow = Flow(RealNVP((d,), n_layers=2))
    g = torch.Generator().manual_seed(0)
    with torch.no_grad():
        for p_ in flow.parameters():
            p_.add_(0.1 * torch.randn(p_.shape, generator=g))
    return flow.eval()


# ---------------------------------------------------------------------------------------------------------------
# CPU: the reference's own sampler classes (oracle/_ref, copied from /root/reference by oracle/build_ref.py) with the
# oracle flow standing in for the absent torchflows; falls back to the oracle port of the same algorithm
# ---------------------------------------------------------------------------------------------------------------
def cpu_kind():
    from oracle.build_ref import import_reference
    return "reference" if import_reference() else "port"


def cpu_run(args, n, n_outer, kind):
    from oracle.potentials_ref import make_potential_ref
    from oracle.realnvp_ref import FlowRef, RealNVPRef
    w = workload_params(args)
    d = args.dim
    torch.manual_seed(0)
    flow = FlowRef(RealNVPRef((d,), n_layers=2))
    flow.load_state_dict(make_bench_flow(d).state_dict())
    flow.eval()
    target = make_potential_ref(w["pot"], (d,))
    x0 = torch.randn(n, d)
    if kind == "reference":
        from nfmc.algorithms.sampling.base import NFMCKernel
        from nfmc.algorithms.sampling.mcmc.hmc import HMCKernel, HMCParameters
        from nfmc.algorithms.sampling.mcmc.langevin import LangevinKernel, LangevinParameters
        from nfmc.algorithms.sampling.nfmc.jump import JumpHMC, JumpMALA, JumpNFMCParameters
        if w["kind"] == 0:
            s = JumpMALA((d,), target, kernel=NFMCKernel((d,), flow=flow), params=JumpNFMCParameters(n_iterations=n_outer),
                         inner_kernel=LangevinKernel(event_size=d), inner_params=LangevinParameters(n_iterations=w["K"]))
        else:
            s = JumpHMC((d,), target, kernel=NFMCKernel((d,), flow=flow), params=JumpNFMCParameters(n_iterations=n_outer),
                        inner_kernel=HMCKernel(event_size=d, step_size=w["step"], n_leapfrog_steps=max(w["L"], 1)),
                        inner_params=HMCParameters(n_iterations=w["K"]))
        t0 = time.perf_counter()
        out = s.sample(x0, show_progress=False)
        dt = time.perf_counter() - t0
        steps = out.samples.shape[0] * n
    else:
        from oracle import samplers_ref as R
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
    kind = cpu_kind()
    for _ in range(args.warmup):
        cpu_run(args, max(256, n // 8), 1, kind)
    t_all, steps_all = 0.0, 0
    for _ in range(args.steps):
        rate, dt = cpu_run(args, n, 1, kind)
        t_all += dt
        steps_all += (w["K"] + 1) * n
    value = steps_all / t_all
    cfg = bench_config(args, w, note="CPU arm: " + ("the reference's own JumpMALA / JumpHMC classes (oracle/_ref) with the oracle RealNVP "
                                                      "standing in for torchflows" if kind == "reference" else
                                                      "oracle port of the reference algorithm") + " on the host cores")
    cfg["chains_per_gpu"] = n                      # what this arm actually runs per step ...
    cfg["extrapolated"] = True                     # ... its chain-steps/s is a per-chain rate, compared as such
    cfg["workload"] = cfg["workload"].replace(f"{args.chains_per_gpu} chains per GPU", f"{n} chains (bounded sample of the {args.chains_per_gpu}-chain workload)")
    line = {
        "impl": "reference", "metric": f"{args.workload} chain-steps/sec", "value": value, "unit": "chain-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_all / max(args.steps, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": cfg,
        "cpu_baseline": {"value": value, "unit": "chain-steps/s", "cores": torch.get_num_threads(), "kind": kind,
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
def bind_to_gpu_numa_node(local_rank):
    """Run this rank's host thread (and, by first touch, its pinned buffers) on the NUMA node the GPU hangs off, so that the
    host side of every H2D / D2H copy is node-local.  Returns a short description for the bench line."""
    try:
        bus = torch.cuda.get_device_properties(local_rank).pci_bus_id
        dom = torch.cuda.get_device_properties(local_rank).pci_domain_id
        dev_id = torch.cuda.get_device_properties(local_rank).pci_device_id
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev_id:02x}.0/numa_node"
        node = int(open(path).read().strip())
        nodes = sorted(int(x[4:]) for x in os.listdir("/sys/devices/system/node") if x.startswith("node") and x[4:].isdigit())
        if node < 0:
            node = nodes[local_rank % len(nodes)] if len(nodes) > 1 else 0
        cpus = []
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus += list(range(int(lo), int(hi or lo) + 1))
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if allowed:
            os.sched_setaffinity(0, allowed)
        return {"numa_node": node, "numa_nodes": len(nodes), "cpus_bound": len(allowed)}
    except Exception as e:  # noqa: BLE001
        return {"numa_node": None, "error": str(e)[:80]}


def fp32_peak_tflops():
    """FFMA peak of this GPU, measured now (tools/microbench_fp32, built by __graft_entry__.build)."""
    exe = os.path.join(ROOT, "tools", "microbench_fp32")
    if not os.path.exists(exe):
        return None
    try:
        out = subprocess.run([exe], capture_output=True, text=True, timeout=60).stdout
        best = 0.0
        for line in out.splitlines():
            if "TFLOP/s" in line and ("FFMA " in line or "FFMA2" in line) and "int" not in line:
                best = max(best, float(line.split("ms")[1].split("TFLOP/s")[0]))
        return best or None
    except Exception:  # noqa: BLE001
        return None


class Arm:
    """One rank's buffers and launch closures for `n` chains starting at global chain `chain0`."""

    def __init__(self, args, w, dev, flow, pot, n, chain0, seed):
        from nfmc_b200 import _native as N
        self.N, self.lib, self.args, self.w, self.dev, self.n, self.chain0, self.seed = N, N.lib(), args, w, dev, n, chain0, seed
        d = args.dim
        self.d, self.K = d, w["K"]
        self.flow, self.pot = flow, pot
        self.pd, self._kp = pot.descriptor(dev)
        self.fd, self._kf = flow.bijection.descriptor(dev)
        gen = torch.Generator(device="cpu").manual_seed(chain0 + 1)
        self.x_host = torch.randn(n, d, generator=gen).pin_memory()
        self.x = self.x_host.to(dev, non_blocking=True)
        self.moments = torch.zeros(2 * d, device=dev, dtype=torch.float64)
        self.counts = torch.zeros(8, device=dev, dtype=torch.int64)
        self.st_local = N.StatsDesc(self.moments.data_ptr(), self.moments.data_ptr() + 8 * d, self.counts.data_ptr())
        self.st_jump = N.StatsDesc(self.moments.data_ptr(), self.moments.data_ptr() + 8 * d, self.counts.data_ptr() + 32)
        self.stream = N.stream_ptr(dev)
        self.logq = torch.empty(n, device=dev, dtype=torch.float32)      # log q(x) between the two kernels of the NF jump
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev) if n * d * 4 <= 126e6 else None
        self.it = 0
        self._e2e = None

    def local_launch(self):
        N, w = self.N, self.w
        rng = N.rng_desc(self.seed, self.it * self.K, None, None)
        if w["kind"] == 0:
            N.check(self.lib.nfmc_mala_steps(C.byref(self.pd), N.ptr(self.x), self.n, self.K, float(w["step"]), None, 1, C.byref(rng),
                                             self.chain0, C.byref(self.st_local), None, self.stream))
        else:
            N.check(self.lib.nfmc_hmc_steps(C.byref(self.pd), N.ptr(self.x), self.n, self.K, float(w["step"]), w["L"], None, 1,
                                            C.byref(rng), self.chain0, C.byref(self.st_local), None, self.stream))

    def jump_launch(self):
        N = self.N
        rng = N.rng_desc(self.seed, self.it, None, None)
        N.check(self.lib.nfmc_jump_step2(C.byref(self.pd), C.byref(self.fd), N.ptr(self.x), N.ptr(self.logq), self.n, 1, C.byref(rng),
                                         self.chain0, C.byref(self.st_jump), None, self.stream))

    def step_launch(self):
        """One step of the hot path: K local steps + 1 NF jump for every chain, through the whole-run entry point (chains
        cut into slabs spread over several streams so that kernels of different slabs overlap)."""
        N, w = self.N, self.w
        N.check(self.lib.nfmc_jump_sample_device(C.byref(self.pd), C.byref(self.fd), N.ptr(self.x), self.n, w["kind"], 1, self.K,
                                                 float(w["step"]), w["L"], None, 1, 1, self.seed, self.it * self.K, self.it, self.chain0,
                                                 C.byref(self.st_local), C.byref(self.st_jump), N.ptr(self.logq), self.stream))
        self.it += 1

    def launches_per_step(self):
        return 3 * int(self.lib.nfmc_jump_sample_slabs(self.d, self.n, 0))

    # -- end to end through the C ABI with host buffers (H2D + run + D2H inside the call) -----------------------------------
    def e2e_setup(self):
        N, d, n = self.N, self.d, self.n
        blob_host = self.flow.bijection.blob(self.dev).cpu().pin_memory()
        ws_bytes = self.lib.nfmc_jump_workspace_bytes(d, n, blob_host.numel())
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=self.dev)
        sx = torch.zeros(d, dtype=torch.float64).pin_memory()
        sx2 = torch.zeros(d, dtype=torch.float64).pin_memory()
        cnt = torch.zeros(8, dtype=torch.int64).pin_memory()
        hp = self.pot.host_params()
        hp_ptr, hp_n = (None, 0) if hp is None else (hp.contiguous().data_ptr(), hp.numel())
        pdesc = N.PotentialDesc(self.pot.kind, d, None, (C.c_float * 4)(*list(self.pot.scalars())[:4]))
        fdesc = N.RealNVPDesc(d, 2, self.fd.n_linear, self.fd.hidden, None, blob_host.numel())
        self._e2e = (blob_host, ws_bytes, ws, sx, sx2, cnt, hp, hp_ptr, hp_n, pdesc, fdesc)
        self.h2d = n * d * 4 + blob_host.numel() * 4 + hp_n * 4
        self.d2h = n * d * 4 + 2 * d * 8 + 64

    def e2e_call(self):
        N, w = self.N, self.w
        blob_host, ws_bytes, ws, sx, sx2, cnt, hp, hp_ptr, hp_n, pdesc, fdesc = self._e2e
        N.check(self.lib.nfmc_jump_sample_host(C.byref(pdesc), hp_ptr, hp_n, C.byref(fdesc), blob_host.data_ptr(),
                                               self.x_host.data_ptr(), self.n, w["kind"], 1, self.K, float(w["step"]), w["L"], self.seed,
                                               self.chain0, sx.data_ptr(), sx2.data_ptr(), cnt.data_ptr(), ws.data_ptr(), ws_bytes,
                                               self.stream))

    def link_ceiling_gbs(self, reps=3):
        """H2D of the state on one stream and D2H on another, at the same time, with this rank's pinned buffers: what the
        host<->device path of nfmc_jump_sample_host could move at best (GB/s, both directions summed)."""
        s1, s2 = torch.cuda.Stream(self.dev), torch.cuda.Stream(self.dev)
        back = torch.empty_like(self.x_host).pin_memory()
        tmp = torch.empty_like(self.x)
        torch.cuda.synchronize(self.dev)
        t0 = time.perf_counter()
        for _ in range(reps):
            with torch.cuda.stream(s1):
                tmp.copy_(self.x_host, non_blocking=True)
            with torch.cuda.stream(s2):
                back.copy_(self.x, non_blocking=True)
        torch.cuda.synchronize(self.dev)
        return reps * 2 * self.x.numel() * 4 / (time.perf_counter() - t0) / 1e9


def native_arm(args):
    import torch.distributed as dist
    from nfmc_b200.potentials import make_potential

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa = bind_to_gpu_numa_node(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    w = workload_params(args)
    d, K = args.dim, w["K"]
    seed = 20261018
    flow = make_bench_flow(d).to(dev)
    pot = make_potential(w["pot"], (d,))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(vals):
        t = torch.tensor(vals, device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(v) for v in t]

    def timed_device(arm, steps, warmup, clocks=None):
        for _ in range(warmup):
            arm.step_launch()
        barrier()
        ev = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(steps)]
        if clocks is not None:
            clocks.start()
        barrier()
        for s_ in range(steps):
            if arm.flush is not None:
                arm.flush.fill_(s_ & 0xFF)
            ev[s_][0].record()
            arm.step_launch()
            ev[s_][1].record()
        barrier()
        clk = clocks.stop() if clocks is not None else None
        return sum(e[0].elapsed_time(e[1]) for e in ev) * 1e-3, clk

    def timed_e2e(arm, steps):
        arm.e2e_setup()
        arm.e2e_call()
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            arm.e2e_call()
        barrier()
        return time.perf_counter() - t0

    # ---- weak scaling: chains_per_gpu chains on every rank (the driver's headline configuration) ------------------------------
    n = args.chains_per_gpu
    arm = Arm(args, w, dev, flow, pot, n, rank * n, seed)
    t_total, clk = timed_device(arm, args.steps, args.warmup, ClockSampler(local_rank) if rank == 0 else None)

    # ---- roofline pass: inside the timed region the launches of different slabs overlap, so a per-launch duration is not
    #      defined there; the dominant kernel is timed here, right after it, alone on the stream, on the same buffers:
    #      one full-batch launch of K local steps per step, then the jump launch --------------------------------------
    r_steps = max(3, min(args.steps, 10))
    evr = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(r_steps)]
    for s_ in range(r_steps):
        if arm.flush is not None:
            arm.flush.fill_(s_ & 0xFF)
        evr[s_][0].record()
        arm.local_launch()
        evr[s_][1].record()
        arm.jump_launch()
        evr[s_][2].record()
        arm.it += 1
    barrier()
    t_local = sum(e[0].elapsed_time(e[1]) for e in evr) * 1e-3 / r_steps * args.steps     # per-launch time x steps
    t_serial = sum(e[0].elapsed_time(e[2]) for e in evr) * 1e-3 / r_steps * args.steps
    t_total, t_local, t_serial = max_over_ranks([t_total, t_local, t_serial])
    value = (K + 1) * n * args.steps * world / t_total

    # pooled statistics: the only collective of the path (moments + counters, once per run)
    if world > 1:
        dist.all_reduce(arm.moments)
        dist.all_reduce(arm.counts)
    acc = arm.counts.cpu().tolist()

    e2e = None
    if not args.no_e2e:
        e_steps = max(3, min(args.steps, 10))
        (te,) = max_over_ranks([timed_e2e(arm, e_steps)])
        (link,) = max_over_ranks([-arm.link_ceiling_gbs()])        # min over ranks of the per-GPU ceiling
        link = -link
        per_gpu_gbs = (arm.h2d + arm.d2h) * e_steps / te / 1e9
        e2e = {"value": (K + 1) * n * e_steps * world / te, "unit": "chain-steps/s",
               "h2d_bytes_per_step": arm.h2d, "d2h_bytes_per_step": arm.d2h,
               "call": "nfmc_jump_sample_host (pinned host x0 in, final state + pooled moments + counters out)",
               "link_gbs": link, "moved_gbs_per_gpu": per_gpu_gbs, "frac_of_link": per_gpu_gbs / link if link else None,
               "link_note": "link_gbs = H2D + D2H of the chain state running concurrently on two streams with this rank's pinned "
                            "buffers, all ranks at once, slowest rank (GB/s, both directions summed)", "numa": numa}
    launches = arm.launches_per_step() * args.steps
    del arm
    torch.cuda.empty_cache()

    # ---- strong scaling: the north star's configuration -- 2^20 chains IN TOTAL over the N ranks -----------------------------
    strong = None
    if not args.no_strong:
        n_s = args.strong_chains // world
        arm_s = Arm(args, w, dev, flow, pot, n_s, rank * n_s, seed)
        (ts,) = max_over_ranks([timed_device(arm_s, args.steps, args.warmup)[0]])
        strong = {"chains_total": n_s * world, "chains_per_gpu": n_s, "value": (K + 1) * n_s * world * args.steps / ts,
                  "unit": "chain-steps/s", "ms_per_step": 1e3 * ts / args.steps, "scaling": "strong"}
        if not args.no_e2e:
            e_steps = max(3, min(args.steps, 10))
            (tes,) = max_over_ranks([timed_e2e(arm_s, e_steps)])
            strong["e2e"] = (K + 1) * n_s * world * e_steps / tes
        del arm_s
        torch.cuda.empty_cache()

    # ---- shard check: the same 4096 global chains sharded over the ranks and alone on rank 0 ---------------------------------
    shard_check = None
    if world > 1:
        n_c = 4096
        part = n_c // world
        a1 = Arm(args, w, dev, flow, pot, part, rank * part, seed)
        gen = torch.Generator().manual_seed(99)
        x_all = torch.randn(n_c, d, generator=gen)
        a1.x.copy_(x_all[rank * part:(rank + 1) * part])
        for _ in range(2):
            a1.step_launch()
        torch.cuda.synchronize(dev)
        dist.all_reduce(a1.moments)
        dist.all_reduce(a1.counts)
        if rank == 0:
            a0 = Arm(args, w, dev, flow, pot, n_c, 0, seed)
            a0.x.copy_(x_all)
            for _ in range(2):
                a0.step_launch()
            torch.cuda.synchronize(dev)
            m_ok = bool(torch.allclose(a0.moments, a1.moments, rtol=1e-9, atol=1e-6))
            c_ok = bool(torch.equal(a0.counts, a1.counts))
            shard_check = "ok" if (m_ok and c_ok) else f"MISMATCH moments={m_ok} counters={c_ok}"
        barrier()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- the dominant kernel against its bound ------------------------------------------------------------------------------------
    hbm_peak, hbm_src = measured_peaks()
    t_launch = t_local / args.steps
    chain_steps_per_launch = float(n) * K
    prof = {}
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            prof = json.load(open(tp))
        except Exception:  # noqa: BLE001
            prof = {}
    ops_per_chain_step = (6 + 30) * d if w["kind"] == 0 else (w["L"] * (2 * 2 + 4) + 30) * d
    fp32_peak = fp32_peak_tflops()
    achieved_tf = ops_per_chain_step * chain_steps_per_launch / t_launch / 1e12
    inst = prof.get(f"{args.workload}_d{d}_inst_per_warp_step")          # ncu: smsp__inst_executed / warp-steps, profiles/
    gs = 4 if d > 64 else (2 if d > 32 else 1)
    frac_issue = None
    frac_fma_pipe = None
    if inst and clk and clk.get("sm_mhz"):
        warp_steps = chain_steps_per_launch * gs / 32.0
        frac_issue = inst * warp_steps / t_launch / (148 * 4 * clk["sm_mhz"] * 1e6)
        pipe = prof.get(f"{args.workload}_d{d}_fma_pipe_cycles_per_warp_step")   # SASS mix x measured pipe rates (profiles/pipes_r02.txt)
        if pipe:
            frac_fma_pipe = pipe * warp_steps / t_launch / (148 * 4 * clk["sm_mhz"] * 1e6)
    roofline = {"bound": "fp32-issue", "kernel": "mala_kernel" if w["kind"] == 0 else "hmc_kernel",
                "achieved": achieved_tf, "peak": fp32_peak, "unit": "TFLOP/s", "frac": (achieved_tf / fp32_peak) if fp32_peak else None,
                "frac_fp32": (achieved_tf / fp32_peak) if fp32_peak else None, "frac_issue": frac_issue, "frac_fma_pipe": frac_fma_pipe,
                "frac_hbm_notional": 8.0 * d * chain_steps_per_launch / t_launch / 1e9 / hbm_peak,
                "traffic": prof.get(f"{args.workload}_d{d}"),
                "peak_source": "FFMA peak measured in this run by tools/microbench_fp32 (fp32 FMA = 2 flop)",
                "hbm_peak": hbm_peak, "hbm_peak_source": hbm_src,
                "kernel_ms_per_launch": 1e3 * t_launch, "kernel_share_of_step": t_local / t_serial,
                "serial_ms_per_step": 1e3 * t_serial / args.steps,
                "timing": "CUDA events around single-stream full-batch launches of the kernel, run right after the timed "
                          "region on the same buffers (inside the timed region launches of different slabs overlap)",
                "note": "achieved = algorithmic fp32 operations per chain-step (SURVEY 8d: (c_U + 30) d) x K n chain-steps per launch / "
                        "launch time; frac_issue = warp-instructions per warp-step (ncu smsp__inst_executed, profiles/) x warp-steps / "
                        "(148 SMs x 4 schedulers x SM clock); frac_fma_pipe = fma-pipe cycles per warp-step (SASS mix x the pipe rates of "
                        "tools/probes/pipes.cu: IMAD.WIDE 4, FFMA2 2, FFMA 1 cycles) over the same denominator -- the busiest pipe; frac_hbm_notional = the 8 d bytes per chain-step of SURVEY 8d -- the kernel "
                        "keeps the state on chip for all K steps (traffic = ncu dram bytes per launch ~ 8 d n), so HBM is not its bound"}
    line = {
        "metric": f"{args.workload} chain-steps/sec", "value": value, "unit": "chain-steps/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_total / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": bench_config(args, w),
        "roofline": roofline,
        "e2e": e2e,
        "strong": strong,
        "gpu_launches": launches,
        "clocks": clk,
        "acceptance": {"local": acc[0] / max(acc[1], 1), "jump": acc[4] / max(acc[5], 1)},
    }
    if shard_check is not None:
        line["shard_check"] = shard_check
    if not args.no_cpu_baseline and world == 1:
        torch.set_num_threads(os.cpu_count() or 1)
        kind = cpu_kind()
        cpu_run(args, 512, 1, kind)
        reps, t_cpu = 4, 0.0                              # ~10 s of CPU work: 4 outer iterations of the bounded sample
        for _ in range(reps):
            t_cpu += cpu_run(args, args.cpu_chains, 1, kind)[1]
        rate = reps * (K + 1) * args.cpu_chains / t_cpu
        line["cpu_baseline"] = {"value": rate, "unit": "chain-steps/s", "cores": torch.get_num_threads(), "kind": kind,
                                "sample": f"{args.cpu_chains} chains x {reps} outer iterations ({K} local steps + 1 jump each), "
                                          f"{t_cpu:.1f} s; per-chain rate (the CPU holds a bounded sample of the workload)"}
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
