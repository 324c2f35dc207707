#!/usr/bin/env python
"""bench.py — headline benchmark of the hot path (BASELINE.json):

    Smolyak d=6 B=4 (InteriorGrid2) `linear_combination` with first derivatives
    (value + gradient, 7 outputs per point) over mixed coordinate transformations (cfg 4b),
    reported as Mpoints/s, with the FP64 roofline fraction of the dominant kernel.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--points P] [--scaling weak|strong] [--impl ours|reference]

One process per GPU (torchrun for N > 1).  A "step" is one pass of linear_combination over this
rank's batch of synthetic points.  Points are independent, so ranks share nothing but the
coefficient vector θ (one NCCL broadcast, outside the timed region): no data-path collective.
`--scaling weak` (default): --points per GPU; `--scaling strong`: --points in total, ceil(P/N) per rank.  `value` = device-resident throughput; `e2e` = the same call through HOST buffers
(pinned), host->device and device->host copies inside the timed region.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

MIXED6 = [(1, 0, 4), (1, -2, 2), (2, 1, 2), (3, 0, 4), (1, -1, 2), (2, -3, 3)]   # SURVEY.md §8(d) cfg 4b
D, B = 6, 4
WORKLOAD = "smolyak_d6_B4_InteriorGrid2_lincomb_value+gradient_cfg4b"


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons, pw = [], [], set(), []
        for ts, line in self.rows:
            f = [v.strip() for v in line.split(",")]
            if len(f) < 8 or not (t0 - 0.05 <= ts <= t1 + 0.15):
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------ reference arm
def oracle_setup():
    from oracle import oracle as orc
    ctor = {1: orc.BoundedLinear, 2: orc.SemiInfRational, 3: orc.InfRational}
    return orc, [ctor[k](a, b) for k, a, b in MIXED6], orc.gradient_spec(D)


def cpu_reference_rate(sample_points: int, seed: int, steps: int = 1, warmup: int = 0, threads: int = 0):
    """Times the CPU oracle (C restatement of the reference algorithm, reference operation order, index
    state machine re-run per point, OpenMP over points) on `sample_points` points of the workload."""
    import numpy as np
    orc, otr, spec = oracle_setup()
    rng = np.random.default_rng(seed)
    theta = np.random.default_rng(4).standard_normal(orc.smolyak_length(2, D, B, B))
    y = orc.transform_from_batch(otr, rng.uniform(-0.99, 0.99, (sample_points, D)))
    cores = threads if threads > 0 else orc.max_threads()
    for _ in range(warmup):
        orc.smolyak_lincomb(2, D, B, B, theta, y[: max(1, sample_points // 8)], trs=otr, spec=spec, nthreads=cores)
    t0 = time.perf_counter()
    for _ in range(steps):
        orc.smolyak_lincomb(2, D, B, B, theta, y, trs=otr, spec=spec, nthreads=cores)
    dt = (time.perf_counter() - t0) / steps
    return sample_points / dt / 1e6, cores, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as orc
    cores = orc.max_threads()
    sample = 240000 * cores
    rate, cores, dt = cpu_reference_rate(sample, seed=104, steps=args.steps, warmup=min(args.warmup, 1))
    rate1, _, _ = cpu_reference_rate(80000, seed=104, threads=1)
    line = {
        "impl": "reference", "metric": "smolyak_d6_B4_linear_combination_Mpoints_per_s", "value": rate, "unit": "Mpoints/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "points_per_step": sample, "note": "Julia is not installed: the reference's CPU path is timed as its C restatement (oracle/sk_oracle.c: per-point index state machine, direct products in the reference's order; gradient requests with the component structure known at compile time and the components in SIMD lanes, as Julia's generated code for SVector-valued expansions; gcc -O3), OpenMP over points"},
        "cpu_baseline": {"value": rate, "unit": "Mpoints/s", "cores": cores, "kind": "port", "value_1thread": rate1,
                         "sample": f"{sample} points of the cfg 4b workload per step, all host threads; value_1thread: 80000 points on one thread"},
        "e2e": {"value": rate, "unit": "Mpoints/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ our arm
def bind_to_gpu_numa_node(local: int):
    """Pins this rank to the CPUs of its GPU's NUMA node before any pinned buffer is allocated (first touch then places the
    staging memory next to the GPU's PCIe root).  Best effort: returns a short description for the JSON line."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(local)
        bus = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read().strip())
        if node < 0:
            return f"gpu {local} ({bus}): no NUMA information"
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return f"gpu {local}: node {node} has no allowed CPUs"
        os.sched_setaffinity(0, cpus)
        return f"gpu {local} ({bus}) -> NUMA node {node}, {len(cpus)} CPUs"
    except Exception as e:       # no sysfs, no permission, old torch: run unpinned
        return f"unpinned ({type(e).__name__})"


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    import __graft_entry__ as ge
    sk = ge.load_package()

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = bind_to_gpu_numa_node(local) if not args.no_numa else "disabled"
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")      # NCCL's version banner / debug lines stay off stdout (one JSON line there)
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    strong = args.scaling == "strong"
    n = (args.points + world - 1) // world if strong else args.points      # strong: ceil(P/N) points per rank
    ctor = {1: sk.BoundedLinear, 2: sk.SemiInfRational, 3: sk.InfRational}
    ct = sk.coordinate_transformations(*[ctor[k](a, b) for k, a, b in MIXED6])
    basis = sk.smolyak_basis(sk.Chebyshev, sk.InteriorGrid2(), sk.SmolyakParameters(B), D) @ ct
    spec = sk.gradient(D)
    plan = sk.get_plan(basis, 0, spec, local)
    info = plan.info
    K = int(info.K)

    # θ: generated on rank 0, broadcast once over NCCL (C1 of SURVEY.md §2a), outside the timed region
    theta = torch.empty(K, dtype=torch.float64, device=dev)
    if rank == 0:
        theta.copy_(torch.as_tensor(np.random.default_rng(4).standard_normal(K)))
    sk.broadcast_coefficients(theta, 0)

    # synthetic points: per-coordinate u ~ U(-0.99, 0.99) mapped by transform_from (SURVEY.md §8(d) cfg 4)
    gen = torch.Generator(device=dev).manual_seed(104 + rank)
    u = torch.rand((n, D), generator=gen, device=dev, dtype=torch.float64) * 1.98 - 0.99
    y = sk.transform_from(None, ct, u)
    del u
    out = torch.empty((n, D + 1), dtype=torch.float64, device=dev)

    import ctypes as C
    L = sk._lib
    stream = torch.cuda.current_stream(dev)

    def step_device():
        L.check(L.lib.sk_linear_combination(plan.handle, C.c_void_p(theta.data_ptr()), K, 1, C.c_void_p(y.data_ptr()), n,
                                            C.c_void_p(out.data_ptr()), L.SK_MEM_DEVICE | L.SK_MODE_FAST, C.c_void_p(stream.cuda_stream)))

    # FP64 denominators (MEASURED_PEAKS.json has no FP64 figure): measured live on this GPU
    # burst = best single launch timed alone; sustained = launches queued back to back, no host sync in between
    dfma_burst, dfma_sust = sk.fp64_peak(local, 0, 0.4)
    _, dfma3_sust = sk.fp64_peak(local, 3, 0.3)       # DFMA with three distinct register operands: register-file bound

    for _ in range(args.warmup):
        step_device()
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.3)
    barrier()
    l0 = sk.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    w0 = time.time()
    e0.record(stream)
    for a, b in evs:
        a.record(stream)
        step_device()
        b.record(stream)
    e1.record(stream)
    barrier()
    w1 = time.time()
    launches = sk.launch_count() - l0
    total_ms = e0.elapsed_time(e1)
    kernel_ms = sum(a.elapsed_time(b) for a, b in evs) / args.steps
    clocks = sampler.stop(w0, w1)
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    n_all = min(args.points, world * n) if strong else world * n
    value = n_all / (ms_per_step * 1e-3) / 1e6

    # ---- end to end through HOST buffers (pinned): H2D of the points and D2H of the results every step
    e2e_n = n
    try:
        import psutil
        need = e2e_n * (D + D + 1) * 8
        if psutil.virtual_memory().available < 3 * need * max(1, world if world <= 8 else 8):
            e2e_n = max(1 << 22, n // 8)
    except Exception:
        pass
    xh = torch.empty((e2e_n, D), dtype=torch.float64, pin_memory=True)
    oh = torch.empty((e2e_n, D + 1), dtype=torch.float64, pin_memory=True)
    xh.copy_(y[:e2e_n])
    th_host = theta.cpu().numpy()

    def step_host():
        L.check(L.lib.sk_linear_combination(plan.handle, C.c_void_p(th_host.ctypes.data), K, 1, C.c_void_p(xh.data_ptr()), e2e_n,
                                            C.c_void_p(oh.data_ptr()), L.SK_MEM_HOST | L.SK_MODE_FAST, C.c_void_p(stream.cuda_stream)))

    e2e_steps = max(1, min(args.steps, 3))
    step_host()
    barrier()
    h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    h0.record(stream)
    for _ in range(e2e_steps):
        step_host()
    h1.record(stream)
    barrier()
    te = torch.tensor([h0.elapsed_time(h1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_ms = float(te.item()) / e2e_steps
    e2e_value = world * e2e_n / (e2e_ms * 1e-3) / 1e6
    # with strong scaling every rank's host buffers hold its own ceil(P/N) slice: e2e_n == n per rank
    # the host path must give the same numbers as the device path
    chk = min(e2e_n, 4096)
    assert torch.equal(oh[:chk], out[:chk].cpu()), "host path and device path disagree"

    # ---- C2 on request (--gather): device-resident full output by one in-place NCCL all-gather, timed separately
    gather = None
    if args.gather and world > 1:
        per = n
        full = torch.empty((per * world, D + 1), dtype=torch.float64, device=dev)
        full[rank * per:(rank + 1) * per].copy_(out)
        mine = full[rank * per:(rank + 1) * per]
        for _ in range(2):
            dist.all_gather_into_tensor(full, mine)
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5
        g0.record(stream)
        for _ in range(reps):
            dist.all_gather_into_tensor(full, mine)
        g1.record(stream)
        barrier()
        tg = torch.tensor([g0.elapsed_time(g1) / reps], dtype=torch.float64, device=dev)
        dist.all_reduce(tg, op=dist.ReduceOp.MAX)
        total_bytes = per * world * (D + 1) * 8
        recv_bytes = per * (world - 1) * (D + 1) * 8
        ok = bool(torch.equal(full[rank * per:rank * per + 64], out[:64]))
        gather = {"collective": "ncclAllGather (torch.distributed.all_gather_into_tensor, in place)", "ms": float(tg.item()), "output_bytes_total": total_bytes,
                  "received_bytes_per_gpu": recv_bytes, "recv_gbs_per_gpu": recv_bytes / (float(tg.item()) * 1e-3) / 1e9,
                  "nvlink_measured_gbs_per_direction": 770.0, "own_slice_intact": ok}
        # the same full output with the gather FUSED into the kernel: every rank's buffer is a torch symmetric-memory allocation
        # (mapped into all ranks), and the kernel's threads store their results into all of them over NVLink
        # (sk_linear_combination_peers).  Timed as compute + gather together.
        try:
            import torch.distributed._symmetric_memory as symm
            fsym = symm.empty((per * world, D + 1), dtype=torch.float64, device=dev)
            hdl = symm.rendezvous(fsym, dist.group.WORLD)
            off = rank * per * (D + 1) * 8
            peers = [int(hdl.buffer_ptrs[r]) + off for r in range(world) if r != rank]
            mine_sym = fsym[rank * per:(rank + 1) * per]
            with torch.cuda.stream(stream):
                fused = False
                for _ in range(2):
                    fused = sk.linear_combination_peers(plan, theta, y, mine_sym, peers, stream)
                barrier()
                f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                f0.record(stream)
                for _ in range(reps):
                    sk.linear_combination_peers(plan, theta, y, mine_sym, peers, stream)
                f1.record(stream)
            barrier()
            tf = torch.tensor([f0.elapsed_time(f1) / reps], dtype=torch.float64, device=dev)
            dist.all_reduce(tf, op=dist.ReduceOp.MAX)
            same = torch.tensor([1.0 if torch.equal(fsym, full) else 0.0], dtype=torch.float64, device=dev)
            dist.all_reduce(same, op=dist.ReduceOp.MIN)
            gather["fused"] = {"what": "compute + gather in one kernel per GPU: results stored into every rank's buffer by the computing threads, peer to peer over NVLink (sk_linear_combination_peers, torch symmetric memory)",
                               "kernel_stored_to_peers": bool(fused), "ms_compute_plus_gather": float(tf.item()),
                               "ms_compute_then_nccl_gather": ms_per_step + float(tg.item()), "equals_nccl_result_on_every_rank": bool(same.item() == 1.0)}
            del fsym
        except Exception as ex:      # symmetric memory unavailable on this build / topology: the NCCL figure above stands alone
            gather["fused"] = {"unavailable": str(ex)[:200]}
        del full

    if rank == 0:
        flops_pt = float(info.flops_per_point_fast)
        hbm_gbs = n * float(info.bytes_per_point) / (kernel_ms * 1e-3) / 1e9
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        # ncu figures of the same kernel build, per point: DRAM bytes and executed FP64 instructions (profiles/traffic.json is
        # written from an `ncu --set full` capture; the run itself cannot measure them)
        traffic = traffic_src = None
        flops_exec = None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            traffic = tj["leaf_dram_bytes_per_point"] * n
            traffic_src = tj.get("source", "profiles/traffic.json (ncu --set full, extrapolated per point)")
            flops_exec = tj.get("leaf_executed_flops_per_point")
        except Exception:
            pass
        flops_used = float(flops_exec) if flops_exec else flops_pt
        achieved = n * flops_used / (kernel_ms * 1e-3) / 1e12
        line = {
            "metric": "smolyak_d6_B4_linear_combination_Mpoints_per_s", "value": value, "unit": "Mpoints/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "points_per_gpu": n, "points_total": n_all, "K": K, "outputs_per_point": D + 1,
                       "basis_term_evals_per_s": value * 1e6 * K, "l2_policy": "inputs (4.8 GB/GPU) and outputs (5.6 GB/GPU) far exceed the 126 MB L2",
                       "parallelism": f"points sharded over {world} GPU(s), θ broadcast once"},
            "roofline": {"bound": "fp64", "achieved": achieved, "peak": dfma_burst, "unit": "TFLOP/s", "frac": achieved / dfma_burst,
                         "traffic": traffic, "traffic_source": traffic_src,
                         "kernel": "smolyak_leaf_kernel<InteriorGrid2, LL=6, GRAD> (spectralkit.jl_b200/csrc/sk_leaf_impl.cuh)", "kernel_ms": kernel_ms,
                         "flops_per_point": flops_used,
                         "flops_per_point_source": "ncu executed FP64 instructions (DFMA = 2, DADD/DMUL = 1) of this kernel build, profiles/traffic.json" if flops_exec else "closed form of the implemented algorithm (smolyak_leaf_flops)",
                         "flops_per_point_closed_form": flops_pt, "flops_per_point_reference_direct": float(info.flops_per_point_direct),
                         "peak_source": "measured live by sk_fp64_peak(DFMA): best single launch (burst); MEASURED_PEAKS.json has no FP64 entry; nominal 37.2 TFLOP/s",
                         "peak_sustained_back_to_back": dfma_sust, "frac_of_sustained": achieved / dfma_sust,
                         "peak_three_register_operands": dfma3_sust,
                         "note": "the register file of a B200 scheduler delivers ONE 64-bit operand per cycle and a DFMA occupies the FP64 pipe for two: fma with three fresh register operands runs at 2/3 of peak (peak_three_register_operands also pays for its LOP3 filler); profiles/micro/dfma_regs.cu, profiles/r2_leaf_register_file.md",
                         "hbm_gbs": hbm_gbs,
                         "hbm_frac_of_measured": hbm_gbs / peaks["hbm_gbs"] if "hbm_gbs" in peaks else None},
            "e2e": {"value": e2e_value, "unit": "Mpoints/s", "h2d_bytes_per_step": e2e_n * D * 8 + K * 8, "d2h_bytes_per_step": e2e_n * (D + 1) * 8,
                    "points_per_gpu": e2e_n, "ms_per_step": e2e_ms, "host_affinity_rank0": numa},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        if gather is not None:
            line["gather"] = gather
        if world == 1 and not args.no_cpu_baseline:
            from oracle import oracle as orc
            cores = orc.max_threads()
            sample = 400000 * cores
            rate, cores, dt = cpu_reference_rate(sample, seed=104)
            rate1, _, dt1 = cpu_reference_rate(80000, seed=104, threads=1)
            line["cpu_baseline"] = {"value": rate, "unit": "Mpoints/s", "cores": cores, "kind": "port", "value_1thread": rate1,
                                    "sample": f"first-seed {sample} points of the same workload, {dt:.1f} s on all host threads (+ 80000 points on one thread, {dt1:.1f} s); C restatement of the reference (Julia not installed)"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--points", type=int, default=100_000_000, help="points per GPU per step")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"], help="weak: --points per GPU; strong: --points in total, ceil(P/N) per rank")
    ap.add_argument("--no-numa", action="store_true", help="do not pin ranks to the NUMA node of their GPU")
    ap.add_argument("--gather", action="store_true", help="also time C2: an in-place NCCL all-gather of the result slices (N > 1)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
