#!/usr/bin/env python
"""bench.py — constraint-rows/sec of the EVM-circuit hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--groups G]

Workload (config.workload): BASELINE cfg2 generator scaled to the metric's 2^20 rows —
per GPU 2^18 groups `PUSH32 b, PUSH32 a, {ADD,SUB,MUL,DIV,MOD}, POP` = 2^20 execution steps,
17.8 M bytecode-table rows, 1.57 M rw-table rows, fixed table 224,490 rows (synthetic, seeded).
One "step" of the bench = one pass of the whole hot path over that witness: build the lookup
indexes of the bytecode / rw tables on the device, then check every execution step.
  value : rows/s with all inputs resident in HBM (CUDA events on the launching stream)
  e2e   : rows/s through the C-ABI with HOST (pinned) buffers: H2D of tables + steps, index
          build, check, D2H of the result vector, all inside the timed region
  roofline : algorithmic bytes of the check kernel / its device time, vs MEASURED_PEAKS.json
  cpu_baseline : the CPU oracle (a C port of the reference algorithm; the reference itself is
          pure Python and absent on this box) on a bounded sample of the same workload
Multi-GPU (torchrun): rows are sharded (each rank checks its own 2^20-step shard against its
replicated tables), then ONE all-reduce(MIN) of the first-fail vector; scaling "weak".
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_CELLS_STEP, N_CELLS_RW, N_CELLS_BYTECODE = 13, 14, 6


class ClockSampler(threading.Thread):
    """samples SM clock + throttle reasons through NVML while the timed region runs (NVML is
    initialised in the constructor, before the region; one sample is taken at start and one at
    stop so that even a 10 ms region is covered)"""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self._stop_evt = index, [], set(), threading.Event()
        self.max_mhz, self._nv, self._h, self._names = None, None, None, {}
        try:
            import pynvml as nv

            nv.nvmlInit()
            self._nv, self._h = nv, nv.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(self._h, nv.NVML_CLOCK_SM)
            self._names = {
                nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
            }
        except Exception as e:  # noqa: BLE001
            self.reasons.add(f"nvml_unavailable:{type(e).__name__}")

    def _sample(self):
        nv = self._nv
        if nv is None:
            return
        try:
            self.samples.append(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM))
            r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
            for bit, nm in self._names.items():
                if r & bit:
                    self.reasons.add(nm)
        except Exception as e:  # noqa: BLE001
            self.reasons.add(f"nvml_unavailable:{type(e).__name__}")

    def run(self):
        while not self._stop_evt.is_set():
            self._sample()
            time.sleep(0.005)

    def stop(self):
        self._sample()  # still under load: the caller stops the sampler before synchronising
        self._stop_evt.set()
        self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons)}


def algorithmic_bytes(n_steps, n_rw, n_bytecode, n_constraints):
    """every witness/table cell once at canonical 32 B; fixed table excluded."""
    return 32 * (n_steps * N_CELLS_STEP + n_rw * N_CELLS_RW + n_bytecode * N_CELLS_BYTECODE) + 4 * n_constraints


def _cpu_port_worker(args):
    sample_groups, seed = args
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib
    from zkevm_specs_b200 import synth
    from zkevm_specs_b200.evm_circuit.table import fixed_table_matrix

    w = synth.evm_trace(sample_groups, seed=seed)
    fixed = fixed_table_matrix()
    oracle_lib.lib()
    t0 = time.perf_counter()
    ff, _ = oracle_lib.check_evm(w["steps"], w["bytecode"], w["rw"], fixed)
    dt = time.perf_counter() - t0
    assert (ff == 0xFFFFFFFF).all(), "oracle rejected the synthetic witness"
    return w["n_steps"], dt


def cpu_port(sample_groups: int, seed: int, threads: int):
    """time the CPU oracle (C port of the reference algorithm) on a bounded sample.  threads > 1:
    that many processes each check their own sample of the same size concurrently (the reference
    has no parallelism of its own; rows are independent, so this is how it would use the cores);
    returns (total rows, wall seconds of the slowest worker)"""
    if threads <= 1:
        return _cpu_port_worker((sample_groups, seed))
    import multiprocessing as mp

    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib

    oracle_lib.lib()  # build once before forking
    with mp.get_context("fork").Pool(threads) as pool:
        t0 = time.perf_counter()
        res = pool.map(_cpu_port_worker, [(sample_groups, seed + k) for k in range(threads)])
        wall = time.perf_counter() - t0
    # generation time is inside `wall`; use the max of the measured check times (they overlap)
    return sum(r[0] for r in res), max(r[1] for r in res)


def run_reference_arm(args, rank, world):
    """--impl reference: the reference's algorithm on the host cores.  The reference is pure
    Python and /root/reference is absent on the GPU box, so this times the C port (oracle/) with
    one process per core (rows are independent: that is how the reference would use the cores).
    A step = every core checks its own bounded sample of the cfg2 trace; the sample is sized from
    a calibration step so that the K timed steps end within ~2 minutes."""
    if rank != 0:
        return
    import multiprocessing as mp

    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib

    oracle_lib.lib()  # build once before forking
    cores = os.cpu_count() or 1
    with mp.get_context("fork").Pool(cores) as pool:
        def step(groups, seed0):
            res = pool.map(_cpu_port_worker, [(groups, seed0 + k) for k in range(cores)])
            return sum(r[0] for r in res), max(r[1] for r in res)

        n, dt = step(256, 2)  # calibration (also the first warm-up)
        rate = n / dt  # rows/s over all cores
        budget = min(3.0, 120.0 / max(1, args.steps))  # seconds per timed step
        sample_groups = int(min(args.ref_groups, max(64, rate * budget / cores / 4)))
        for _ in range(max(0, args.warmup - 1)):
            step(sample_groups, 2)
        rows = secs = 0.0
        for k in range(args.steps):
            n, dt = step(sample_groups, 2 + k)
            rows += n
            secs += dt
    v = rows / secs
    line = {
        "impl": "reference", "metric": "constraint-rows/sec", "value": v, "unit": "rows/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * secs / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64 limbs (BN254 Fr, 254-bit modular)",
        "data": "synthetic",
        "config": {"workload": f"evm_circuit ADD/SUB/MUL/DIV/MOD trace (cfg2 generator), bounded sample: {cores} x "
                               f"{4 * sample_groups} steps per bench step (same generator as the CUDA arm)", "seed": 2},
        "cpu_baseline": {"value": v, "unit": "rows/s", "cores": cores, "kind": "port",
                         "sample": f"{cores} processes x {4 * sample_groups} steps each per bench step, C oracle incl. "
                                   "sorted-index build of all tables"},
        "e2e": {"value": v, "unit": "rows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--groups", type=int, default=1 << 18, help="trace groups per GPU (4 steps each)")
    ap.add_argument("--ref-groups", type=int, default=1 << 12)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--storage", choices=["adaptive", "typed", "packed", "canonical"], default="adaptive",
                    help="adaptive: the packer's default — every column at its measured minimal width, constant "
                         "columns stored once (packing.pack_matrix); typed (= packed): data-independent widths by "
                         "column type (packing.TYPE_WIDTHS); both through zk_upload_*_packed.  canonical: 32-byte "
                         "cells through zk_upload_*")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    from zkevm_specs_b200 import native, synth
    from zkevm_specs_b200.evm_circuit.table import fixed_table_matrix

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = native.Context(local)
    stream = torch.cuda.current_stream().cuda_stream

    # ---- synthetic witness (per rank: its own row shard, own seed) -------------------------
    w = synth.evm_trace(args.groups, seed=2 + rank)
    n_steps = w["n_steps"]
    fixed = fixed_table_matrix()
    n_rw, n_bc = w["rw"].shape[1], w["bytecode"].shape[1]
    n_constraints = ctx.n_constraints(native.CIRCUIT_EVM)
    ctx.upload_table(native.TABLE_FIXED, fixed, stream=stream)  # circuit constant: uploaded once
    if args.storage == "packed":
        args.storage = "typed"
    if args.storage in ("adaptive", "typed"):
        from zkevm_specs_b200 import packing
        if args.storage == "typed":
            packed = {"steps": packing.pack_matrix(w["steps"], min_widths=packing.TYPE_WIDTHS["evm_steps"]),
                      "bytecode": packing.pack_matrix(w["bytecode"], min_widths=packing.TYPE_WIDTHS["bytecode_table"]),
                      "rw": packing.pack_matrix(w["rw"], min_widths=packing.TYPE_WIDTHS["rw_table"])}
            fmt = "packed columns, data-independent type widths (packing.TYPE_WIDTHS)"
        else:
            w["bytecode"] = None  # 3.4 GB of unrolled canonical rows: not needed on this path (8 ranks share the host)
            packed = {k: packing.pack_matrix(w[k]) for k in ("steps", "rw")}
            fmt = ("steps / rw table: packed columns, measured minimal width per column, constant columns stored "
                   "once (packing.pack_matrix default; the packing scan is host preparation, outside the timed "
                   "region).  bytecode table: the raw code bytes + 1 is_code bit per byte + 1 hash per contract, "
                   "unrolled on the device (zk_upload_bytecode_table_from_code = Bytecode.table_assignments, "
                   "typing.py:390-427), inside the timed region")
        pinned = {k: torch.from_numpy(pm.buf).pin_memory() for k, pm in packed.items()}
        h2d_bytes = sum(pm.nbytes for pm in packed.values())
        src = None
        if "bytecode" not in packed:
            src = w["bytecode_src"]
            pinned["code"] = torch.from_numpy(src["code"]).pin_memory()
            pinned["bits"] = torch.from_numpy(src["is_code_bits"]).pin_memory()
            h2d_bytes += src["code"].nbytes + src["is_code_bits"].nbytes + src["code_offsets"].nbytes + src["hashes"].nbytes
        stored = sum(pm.nbytes for pm in packed.values()) + (42 * n_bc if src is not None else 0)
        widths = {k: [int(x) for x in pm.widths] for k, pm in packed.items()}
        if src is not None:
            widths["bytecode"] = [16, 16, 1, 4, 1, 4]  # written by k_bytecode_table_expand
        storage = {"format": fmt, "widths": widths, "stored_bytes": stored}

        def upload_inputs():
            if src is not None:
                ctx.upload_bytecode_table_from_code(src["code"], src["is_code_bits"], src["code_offsets"], src["hashes"],
                                                    stream=stream, ptrs=(pinned["code"].data_ptr(), pinned["bits"].data_ptr()))
            else:
                ctx.upload_table_packed(native.TABLE_BYTECODE, packed["bytecode"], stream=stream,
                                        host_ptr=pinned["bytecode"].data_ptr())
            ctx.upload_table_packed(native.TABLE_RW, packed["rw"], stream=stream, host_ptr=pinned["rw"].data_ptr())
            ctx.upload_columns_packed(native.CIRCUIT_EVM, packed["steps"], stream=stream, host_ptr=pinned["steps"].data_ptr())
    else:
        pinned = {k: torch.from_numpy(w[k]).pin_memory() for k in ("steps", "bytecode", "rw")}
        h2d_bytes = 32 * ((n_steps + 1) * N_CELLS_STEP + n_rw * N_CELLS_RW + n_bc * N_CELLS_BYTECODE)
        storage = {"format": "canonical 32-byte cells", "stored_bytes": h2d_bytes}

        def upload_inputs():
            ctx.upload_table_ptr(native.TABLE_BYTECODE, n_bc, 6, pinned["bytecode"].data_ptr(), stream)
            ctx.upload_table_ptr(native.TABLE_RW, n_rw, 14, pinned["rw"].data_ptr(), stream)
            ctx.upload_columns_ptr(native.CIRCUIT_EVM, n_steps + 1, 13, pinned["steps"].data_ptr(), stream)

    # result buffer as a torch tensor (zero-copy) for the NCCL all-reduce
    class _Raw:
        def __init__(self, ptr, n):
            self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i4", "data": (ptr, False), "version": 2}

    def hot_path():
        """index builds + check of every step; + the one collective when sharded"""
        ctx.invalidate_indexes()
        ctx.check_async(native.CIRCUIT_EVM, 0, n_steps, rank * n_steps, 0, stream)
        if world > 1:
            ff_ptr, _ = ctx.result_device_ptrs(native.CIRCUIT_EVM)
            t = torch.as_tensor(_Raw(ff_ptr, n_constraints), device=f"cuda:{local}")
            # first_fail is uint32 with 0xFFFFFFFF = pass; as int32 that is -1, so MIN over uint32
            # == MIN over (x ^ 0x80000000) as int32
            t ^= -0x80000000
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            t ^= -0x80000000

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, k):
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        sync_all()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=f"cuda:{local}")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    upload_inputs()
    for _ in range(args.warmup):
        hot_path()
    ff, fc = ctx.fetch_result(native.CIRCUIT_EVM, stream)
    assert (ff == native.PASS).all(), f"witness rejected: {native.first_failure(ff, native.CIRCUIT_EVM)}"

    sampler = ClockSampler(local)
    sampler.start()
    launches0 = ctx.launch_count()
    ms = timed(hot_path, args.steps)
    launches = ctx.launch_count() - launches0
    clocks = sampler.stop()
    value = world * n_steps * args.steps / (ms / 1e3)

    # ---- roofline of the dominant kernel (k_check_evm), device-timed per launch -------------
    ctx.enable_timing(True)
    idx_ms, chk_ms = [], []
    for _ in range(args.steps):
        ctx.invalidate_indexes()
        ctx.check_async(native.CIRCUIT_EVM, 0, n_steps, rank * n_steps, 0, stream)
        a, b = ctx.last_timing()
        idx_ms.append(a)
        chk_ms.append(b)
    ctx.enable_timing(False)
    bytes_alg = algorithmic_bytes(n_steps, n_rw, n_bc, n_constraints)
    chk = float(np.mean(chk_ms))
    achieved = bytes_alg / (chk / 1e3) / 1e9
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:  # noqa: BLE001
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    traffic = None
    try:  # DRAM bytes of the same kernels from the committed `ncu --set full` capture of this command
        tj = json.load(open(os.path.join(ROOT, "profiles", "r01_final_traffic.json")))
        traffic = tj["dram_bytes_per_check"] if tj.get("storage", "canonical") == args.storage else None
    except Exception:  # noqa: BLE001
        pass
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic if args.groups == (1 << 18) else None,
                "kernel": "evm check phase: k_evm_classify + k_evm_push_pos (~33 %) + k_evm_gadget<ADD|MUL|POP, POS> (MUL ~35 %) + k_evm_misc",
                "kernel_ms": chk, "index_build_ms": float(np.mean(idx_ms)),
                "algorithmic_bytes": bytes_alg, "stored_bytes": storage["stored_bytes"],
                "achieved_stored_gbs": storage["stored_bytes"] / (chk / 1e3) / 1e9,
                "peak_source": "MEASURED_PEAKS.json (burst copy)" if peaks else "fallback 6.65 TB/s"}

    # ---- e2e: host buffers through the C-ABI, copies inside the timed region ----------------
    e2e = None
    if not args.no_e2e:
        def e2e_step():
            upload_inputs()
            hot_path()
            ctx.fetch_result(native.CIRCUIT_EVM, stream)

        e2e_step()
        k_e2e = max(3, min(args.steps, 20))
        ms_e2e = timed(e2e_step, k_e2e)
        e2e = {"value": world * n_steps * k_e2e / (ms_e2e / 1e3), "unit": "rows/s", "h2d_bytes_per_step": h2d_bytes,
               "d2h_bytes_per_step": n_constraints * 12, "steps": k_e2e, "ms_per_step": ms_e2e / k_e2e}

    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        n, dt = cpu_port(args.ref_groups, 2, cores)
        cpu = {"value": n / dt, "unit": "rows/s", "cores": cores, "kind": "port",
               "sample": f"{cores} processes x {n // cores} steps of the same generator, C oracle incl. sorted-index build"}

    if rank == 0:
        line = {
            "metric": "constraint-rows/sec", "value": value, "unit": "rows/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u64 limbs (BN254 Fr, 254-bit modular)", "data": "synthetic",
            "config": {"workload": f"evm_circuit ADD/SUB/MUL/DIV/MOD trace (cfg2 generator), {n_steps} steps per GPU",
                       "steps_per_gpu": n_steps, "rw_rows": n_rw, "bytecode_rows": n_bc, "fixed_rows": int(fixed.shape[1]),
                       "parallelism": f"row-shard x{world}, tables replicated, 1 all-reduce(min)",
                       "l2": "inputs larger than L2 (%.2f GB stored per GPU)" % (storage["stored_bytes"] / 1e9),
                       "storage": storage,
                       "timed_region": "lookup-index build of bytecode+rw tables, then step check"},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
