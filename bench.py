#!/usr/bin/env python
"""Benchmark of the hot path named by BASELINE.json: PauliwordOp multiply + cleanup (cross-terms/s).

    python bench.py --gpus N --steps K --warmup W            # this engine (one rank per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU algorithm (oracle port)

Workload (config C5 of BASELINE.json, weak-scaled): 1000-qubit product of a 12 500*N-term operator A
(each rank holds one 12 500-row block, seeds 100+rank) with a 10 000-term operator B (seed 7),
i.e. 1.25e8 cross terms per GPU; at N = 8 this is exactly C5 (1e5 x 1e4 terms, 1e9 cross terms).
Inputs are generated on the host with the reference's own generator (PauliwordOp.random).
A step = one full product + cleanup. Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_QUBITS = 1000
ROWS_A_PER_GPU = 12500
ROWS_B = 10000
ROW_BYTES = 16 * ((N_QUBITS + 63) // 64) + 16          # R of SURVEY.md §8d: packed row + coefficient = 272 B
CPU_SAMPLE = (250, 2000)                                # bounded CPU sample: 5e5 cross terms of the same workload


def make_operator(n_terms, seed):
    from oracle import pauli_oracle as po           # only the input generator (same draws as the reference)
    return po.random_operator(N_QUBITS, n_terms, seed=seed)


def cpu_sample_inputs():
    a_s, a_c = make_operator(ROWS_A_PER_GPU, 100)
    b_s, b_c = make_operator(ROWS_B, 7)
    ma, mb = CPU_SAMPLE
    return a_s[:ma], a_c[:ma], b_s[:mb], b_c[:mb]


def time_cpu_reference(steps, warmup):
    """The reference's NumPy algorithm (oracle port, single core like the reference) on the bounded sample."""
    from oracle import pauli_oracle as po
    a_s, a_c, b_s, b_c = cpu_sample_inputs()
    T = a_s.shape[0] * b_s.shape[0]
    for _ in range(warmup):
        po.multiply(a_s, a_c, b_s, b_c)
    t0 = time.perf_counter()
    for _ in range(steps):
        po.multiply(a_s, a_c, b_s, b_c)
    dt = (time.perf_counter() - t0) / max(1, steps)
    return T / dt, dt, T


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 10))
    value, dt, T = time_cpu_reference(steps, max(0, min(args.warmup, 1)))
    sample = (f"{CPU_SAMPLE[0]}x{CPU_SAMPLE[1]} terms of the same operators = {T} cross terms per step "
              f"(full size needs a 250 GB bool tensor on the reference's path)")
    line = {
        "impl": "reference", "metric": "cross-terms/s (multiply+cleanup)", "value": value, "unit": "cross-terms/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": min(args.warmup, 1), "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64+c128", "data": "synthetic",
        "config": {"workload": "C5/8 per GPU: 1000 q, 12500x10000-term product + cleanup (bounded CPU sample)",
                   "n_qubits": N_QUBITS, "sample": sample},
        "cpu_baseline": {"value": value, "unit": "cross-terms/s", "cores": 1, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "cross-terms/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


class ClockSampler:
    """nvidia-smi clocks/throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.path = tempfile.NamedTemporaryFile(prefix="clocks_", suffix=".csv", delete=False).name
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in open(self.path):
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for nm, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        os.unlink(self.path)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smax)), "reasons": sorted(reasons),
                "samples": len(sm)}


def run_ours(args):
    import torch
    import torch.distributed as dist

    from symmer_b200 import dist as sdist
    from symmer_b200 import ops

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local_rank)
    dev = ops.device()
    if os.environ.get("SYMMER_EMIT_VARIANT"):                      # A/B knob: 0 = two-kernel compaction + emission
        ops.set_tuning(1, int(os.environ["SYMMER_EMIT_VARIANT"]))
    tuning = {}
    for kv in filter(None, os.environ.get("SYMMER_TUNING", "").split(",")):   # A/B knobs, e.g. "6=0,7=32"
        k, v = kv.split("=")
        tuning[int(k)] = int(v)
        ops.set_tuning(int(k), int(v))
    tile_mode = tuning.get(6, 1) != 0                              # ordered-tile mode (default) vs sorted-hash order
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    # ---------------------------------------------------------------- inputs (host, reference generator)
    a_s, a_c = make_operator(ROWS_A_PER_GPU, 100 + rank)
    b_s, b_c = make_operator(ROWS_B, 7)
    host = [torch.from_numpy(x).pin_memory() for x in (a_s, a_c, b_s, b_c)]
    h2d_bytes = int(sum(t.numel() * t.element_size() for t in host))
    T_local = ROWS_A_PER_GPU * ROWS_B
    T_total = T_local * world

    def upload():
        a_bool, ac, b_bool, bc = [t.to(dev, non_blocking=True) for t in host]
        return ops.pack(a_bool, N_QUBITS), ac, ops.pack(b_bool, N_QUBITS), bc

    def product(a, ac, b, bc):
        if world == 1:
            return ops.mul_cleanup(a, ac, b, bc, 1e-15)
        xz, c, _ = sdist.sharded_product(a, ac, b, bc, 1e-15, method=os.environ.get("SYMMER_DIST_METHOD", "owner"))
        return xz, c

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def timed(fn, n):
        """n steps, each bracketed by CUDA events on the launching stream; L2 flushed between steps."""
        ms = []
        for _ in range(n):
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = fn()
            e1.record()
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
            del out
        return ms

    # ---------------------------------------------------------------- device-resident arm (`value`)
    a, ac, b, bc = upload()
    state = {}

    def step_resident():
        xz, c = product(a, ac, b, bc)
        state["U"] = xz.shape[0]
        return xz, c

    clocks = ClockSampler(local_rank)
    clocks.start()
    for _ in range(max(3, args.warmup)):
        timed(step_resident, 1)
    barrier()
    ops.emit_events = []
    ops.emit_kernel_events = []
    launches0 = ops.launch_count()
    t_wall0 = time.perf_counter()
    ms_steps = timed(step_resident, args.steps)
    barrier()
    wall = time.perf_counter() - t_wall0
    launches = ops.launch_count() - launches0
    emit_ms = [s.elapsed_time(e) for s, e in ops.emit_events]
    kern_ms = [s.elapsed_time(e) for s, e in ops.emit_kernel_events]
    ops.emit_events = None
    ops.emit_kernel_events = None
    step_ms = float(np.mean(ms_steps))

    # ---------------------------------------------------------------- end-to-end arm (`e2e`): host buffers in, result summary out
    def step_e2e():
        ua, uac, ub, ubc = upload()
        xz, c = product(ua, uac, ub, ubc)
        summary = torch.stack([torch.sum(c), torch.tensor(complex(xz.shape[0]), device=dev, dtype=torch.complex128)])
        return summary.cpu()                                        # device->host read of the step's result

    quick = os.environ.get("SYMMER_BENCH_QUICK") == "1"      # profiling runs (ncu): device-resident arm only
    timed(step_e2e, 1)
    barrier()
    e2e_ms = float(np.mean(timed(step_e2e, 1 if quick else max(1, min(args.steps, 10)))))
    barrier()
    clock_info = clocks.stop()

    # ---------------------------------------------------------------- secondary metric of BASELINE.json: commute-pair checks/s
    commute = None
    if world == 1 and not quick:
        gen = torch.Generator(device=dev)
        gen.manual_seed(3)
        big = torch.randint(-2 ** 63, 2 ** 63 - 1, (65536, a.shape[1]), dtype=torch.int64, device=dev, generator=gen)
        blk = big[:16384].contiguous()
        ops.commute(blk, big)
        torch.cuda.synchronize()
        cms = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            adj = ops.commute(blk, big)
            e1.record()
            torch.cuda.synchronize()
            cms.append(e0.elapsed_time(e1))
            del adj
        pairs = 16384 * 65536
        del big, blk
        commute = {"metric": "commute-pair checks/s", "value": pairs / (min(cms) * 1e-3), "unit": "pairs/s",
                   "workload": "16384 x 65536 block of a 1024-bit-wide (1000-qubit layout) adjacency matrix, random rows",
                   "kernel": "commute_mma_kernel (tcgen05 kind::i8, TMEM accumulators)",
                   "int8_tops": pairs * 2 * 2048 / (min(cms) * 1e-3) / 1e12,
                   "note": "K = 2048 unpacked bits per pair; nominal dense int8 peak 4500 TOP/s"}

    # ---------------------------------------------------------------- max over ranks
    if world > 1:
        t = torch.tensor([step_ms, e2e_ms, float(np.mean(emit_ms)) if emit_ms else 0.0,
                          float(np.mean(kern_ms)) if kern_ms else 0.0], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        step_ms, e2e_ms, emit_mean, kern_mean = [float(x) for x in t.cpu()]
        u = torch.tensor([state["U"]], dtype=torch.int64, device=dev)
        dist.all_reduce(u)
        U_total = int(u.item())
    else:
        emit_mean = float(np.mean(emit_ms)) if emit_ms else 0.0
        kern_mean = float(np.mean(kern_ms)) if kern_ms else 0.0
        U_total = state["U"]

    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        U_local = state["U"]
        row_only = 16 * ((N_QUBITS + 63) // 64)
        # tile_emit_kernel writes row + coefficient (272 B per survivor); in the sorted-hash path emit_kernel
        # writes the 256 B row and the coefficient leaves in compact_kernel
        emit_bytes = U_local * (ROW_BYTES if tile_mode else row_only)
        achieved = emit_bytes / (kern_mean * 1e-3) / 1e9 if kern_mean > 0 else None
        phase_gbs = U_local * ROW_BYTES / (emit_mean * 1e-3) / 1e9 if emit_mean > 0 else None
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "tile_emit_traffic.json" if tile_mode else "emit_traffic.json")
        if os.path.exists(tpath):
            traffic = float(json.load(open(tpath))["dram_bytes_per_row"]) * U_local
        path_bytes = T_local * (2 * ROW_BYTES + (U_local / T_local) * ROW_BYTES)   # SURVEY §8d model: 816 B/ct at U=T
        cpu_value, cpu_dt, cpu_T = (None, None, None)
        if world == 1 and not quick:
            cpu_value, cpu_dt, cpu_T = time_cpu_reference(steps=3, warmup=1)
        line = {
            "metric": "cross-terms/s (multiply+cleanup)", "value": T_total / (step_ms * 1e-3), "unit": "cross-terms/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": step_ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64+c128", "data": "synthetic",
            "config": {"workload": "C5/8 per GPU: 1000 q, (12500 x n_gpus)-term A times 10000-term B, product + cleanup; "
                                   "n_gpus=8 is BASELINE config C5 (1e9 cross terms)",
                       "n_qubits": N_QUBITS, "rows_a_per_gpu": ROWS_A_PER_GPU, "rows_b": ROWS_B,
                       "cross_terms_total": T_total, "unique_terms_total": U_total,
                       "parallelism": ("term-block sharded A all-gathered once; exchange-free hash partition (GF(2)-linear owner classes): "
                                       "every rank generates and dedups exactly the cross terms it owns") if world > 1 else "1 GPU",
                       "l2": "explicit 256 MB flush write between timed steps; per-step working set ~40 GB >> L2",
                       "output_materialised": True},
            "e2e": {"value": T_total / (e2e_ms * 1e-3), "unit": "cross-terms/s", "h2d_bytes_per_step": h2d_bytes,
                    "d2h_bytes_per_step": 32, "ms_per_step": e2e_ms,
                    "note": "host bool[M,2n]+complex128 operands (pinned) -> device, pack, product+cleanup; the result "
                            "operator stays device-resident (as in the reference-facing API), its term count and "
                            "coefficient checksum are read back"},
            "gpu_launches": int(launches),
            "clocks": clock_info,
            "roofline": {"bound": "hbm", "kernel": ("tile_emit_kernel (rows + coefficients of the survivors in cross-term order: "
                                                    "272 B per survivor, one launch per block)") if tile_mode else
                                                   "emit_kernel (row emission of the survivors: 256 B per row, one launch per step)",
                         "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": (achieved / peak) if achieved else None, "traffic": traffic, "peak_source": peak_src,
                         "kernel_ms": kern_mean, "kernel_share_of_step": kern_mean / step_ms if step_ms else None,
                         "emit_phase_ms": emit_mean, "emit_phase_gbs": phase_gbs,
                         "emit_phase_note": ("tile_emit_kernel + fix-up of group sums" if tile_mode else
                                             "compact_kernel + emit_kernel") + ": 272 B (row + coefficient) per survivor",
                         "tuning": tuning,
                         "path_model_bytes_per_step": path_bytes,
                         "path_achieved_gbs": path_bytes / (step_ms * 1e-3) / 1e9,
                         "path_frac": path_bytes / (step_ms * 1e-3) / 1e9 / peak},
            "cpu_baseline": ({"value": cpu_value, "unit": "cross-terms/s", "cores": 1, "kind": "port",
                              "sample": f"{CPU_SAMPLE[0]}x{CPU_SAMPLE[1]} terms of the same operators = {cpu_T} cross "
                                        f"terms, {cpu_dt:.2f} s per pass, NumPy single core like the reference"}
                             if cpu_value else None),
            "secondary": commute,
            "wall_s_timed_region": wall,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
