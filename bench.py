#!/usr/bin/env python3
"""bench.py -- the Plymouth-adjacency query at 10k synthetic universes (BASELINE.json configs[3]) on N B200s.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (libcolq.so)
    python bench.py --impl reference --gpus N --steps K ...  # the reference engine's CPU algorithm (oracle port)

One "step" = one execution of the query (app/src/main/java/dgroomes/app/Runner.java:230-236) over the whole
workload: 293,530,000 ZIP rows / 257,010,000 city rows, sharded by universe range over the N ranks (strong scaling),
state table replicated.  metric = root-table (ZIP) rows per second, whole job.

  value     K steps timed with CUDA events on the launching stream, tables resident in HBM, max over ranks.
            Every step enqueues all kernels, the state-mask all-gather and the final index gather (N>1).
  e2e       the same query through the C ABI with HOST buffers: per step all columns are copied from pinned host
            memory into HBM (colq_col_* / colq_associate_*), the query runs, the matched indices are read back.
  roofline  the dominant kernel (the TMA-staged city-name scan), timed per launch with CUDA events inside libcolq
            (COLQ_OPT_PROFILE) in separate profiled steps; algorithmic bytes = offsets + name bytes + mask out.
  cpu_baseline  the oracle port (oracle/colq_oracle.c), single thread like the serial reference, bounded sample.

Inputs are far larger than L2 (6.6 GB touched per step on one GPU, 0.82 GB per GPU on eight; L2 is 126 MB), so no
explicit flush between steps.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
# stdout carries exactly ONE JSON line.  Native libraries print there too (NCCL's "NCCL version ..." banner on rank 0), so
# file descriptor 1 is pointed at stderr for the whole run and the JSON line is written to the saved, real stdout.
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(line: dict) -> None:
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())

for p in (ROOT / "java-columnar-query-engine_b200", ROOT / "oracle"):
    sys.path.insert(0, str(p))

import numpy as np  # noqa: E402

N_ZIPS, N_CITIES = 29_353, 25_701
METRIC = "plymouth_query_zip_rows_per_sec"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="colq", choices=["colq", "reference"])
    ap.add_argument("--universes", type=int, default=10_000)
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--profile-steps", type=int, default=10)
    ap.add_argument("--cpu-universes", type=int, default=1000, help="bounded sample for the serial cpu_baseline leg of the GPU arm")
    ap.add_argument("--reference-universes", type=int, default=0,
                    help="--impl reference only: run the CPU arm on this many universes instead of the full --universes "
                         "(a bounded sample for hosts with little RAM: the full 10k-universe tables need ~30 GB)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ingest", action="store_true", help="skip the load-time (ingest) kernel measurements")
    ap.add_argument("--no-small-queries", action="store_true", help="skip the 1-universe latency measurements (keeps ncu launch lists clean)")
    ap.add_argument("--eager", action="store_true", help="disable the lazy FK chain (materialise every node)")
    ap.add_argument("--weak", action="store_true",
                    help="weak-scaling variant (not the driver's line): --universes per GPU instead of in total, i.e. every rank "
                         "holds the N=1 workload; reported with scaling=weak")
    ap.add_argument("--dict-names", action="store_true",
                    help="SURVEY 8f variant, not the headline: the city-name column is dictionary-encoded (int32 codes + 16.6k distinct "
                         "names); the name predicate runs over the dictionary and the row scan tests code bits. Implies --no-e2e.")
    ap.add_argument("--workload", default="plymouth", choices=["plymouth", "int_scan", "str_eq"],
                    help="plymouth = BASELINE configs[3] (the headline); int_scan = configs[1] (1B-row int range scan + "
                         "compaction); str_eq = configs[4] (city-name equality, one GPU's 62.5M-row shard). The last two are "
                         "single-GPU kernel benchmarks for DESIGN.md, not the driver's bench line.")
    ap.add_argument("--rows", type=int, default=0, help="row count override for int_scan / str_eq")
    ap.add_argument("--unfused-root", action="store_true", help="COLQ_OPT_ROOT_FUSED=0: scan_rows / csr_pull / compact_fused launches (A/B runs)")
    return ap.parse_args()


def workload_config(U):
    """The workload alone -- the SAME dict on the GPU arm and on the reference arm (how each arm executes it, sharding and
    strategy, are top-level keys of the GPU line)."""
    return {
        "workload": "plymouth_adjacency_query_10k_universes" if U == 10_000 else f"plymouth_adjacency_query_{U}_universes",
        "source": "BASELINE.json configs[3]; app/.../Runner.java:230-236",
        "universes": U, "zip_rows": U * N_ZIPS, "city_rows": U * N_CITIES, "state_rows": 51,
        "l2_policy": "inputs larger than L2 (no flush)",
    }


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index: int, period_s: float = 0.005):
        self.index, self.period = index, period_s
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as e:  # pragma: no cover
            self.nv, self.err = None, repr(e)

    def _run(self):
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
        }
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or getattr(
            nv, "nvmlDeviceGetCurrentClocksThrottleReasons")
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = get_reasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.period)

    def start(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()

    def stop(self):
        self._stop.set()
        if self._thread:
            self._thread.join()
        return {
            "sm_mhz": statistics.median(self.samples) if self.samples else None,
            "sm_max_mhz": self.max_mhz,
            "reasons": sorted(self.reasons),
            "samples": len(self.samples),
        }


def measured_peak():
    f = ROOT / "MEASURED_PEAKS.json"
    if f.exists():
        try:
            return float(json.loads(f.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(kernel_prefix: str):
    """dram bytes per launch of the dominant kernel from the committed ncu --set full capture, if any."""
    f = ROOT / "profiles" / "roofline_traffic.json"
    if f.exists():
        try:
            d = json.loads(f.read_text())
            for k, v in d.get("kernels", {}).items():
                if kernel_prefix.startswith(k):
                    return v.get("dram_bytes_per_launch")
        except Exception:
            pass
    return None


def bind_to_gpu_numa_node(local_rank: int):
    """Pin this rank's process to the CPUs next to its GPU before any pinned host buffer is allocated, so that the
    buffers the e2e leg streams over PCIe are first-touched on the GPU's own NUMA node (no cross-socket hop)."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(local_rank)
        path = f"/sys/bus/pci/devices/{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        node = int(open(path + "/numa_node").read())
        cpus = set()
        for part in open(path + "/local_cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if node < 0 or not cpus:
            return {"numa_node": node, "bound": False}
        os.sched_setaffinity(0, cpus)
        return {"numa_node": node, "bound": True, "cpus": len(cpus)}
    except Exception as e:  # best effort: a box without sysfs topology just keeps the default placement
        return {"bound": False, "error": repr(e)}


# ------------------------------------------------------------------------------------------------ CPU legs
def oracle_run(U, n_threads, repeats):
    """Times orc_execute (tables already resident in host memory, like the reference's execute) on U universes."""
    from colq import geography as G
    from oracle_system import OracleDataSystem
    geo = G.build_tables(U)
    ds = OracleDataSystem(n_threads=n_threads)
    G.register_geography(ds, geo)
    q = G.plymouth_query()
    times = []
    for _ in range(repeats):
        r = ds.execute(q)
        assert r.result_set.size() == 31 * U
        times.append(ds.last_execute_seconds)
    ds.close()
    return times


def run_reference(args, rank, world):
    """The reference arm: the reference engine's own algorithm on the host cores, on the SAME config as the GPU arm (all
    --universes universes; every step is one full execute of the Plymouth query over resident host tables).  The reference
    is Java and neither this image nor the GPU box has a JVM, so this times the C port of it
    (oracle/colq_oracle.c) with all host threads."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    U = args.reference_universes or args.universes
    if not args.reference_universes:
        # the full tables (7.75 GB of columns, the CSR forms and the transposed reverse columns of the literal port) peak
        # at ~3.4 MB of host RAM per universe; refuse to swap the box to death and fall back to the largest sample that fits
        try:
            import psutil
            fit = int(psutil.virtual_memory().available * 0.8 / 3.4e6)
            if fit < U:
                U = max(100, fit)
        except Exception:
            pass
    times = oracle_run(U, cores, args.warmup + args.steps)[args.warmup:]
    sec = statistics.mean(times)
    value = U * N_ZIPS / sec
    sample = (f"all {U} universes ({U * N_ZIPS} ZIP rows) per step" if U == args.universes else
              f"{U} of {args.universes} universes ({U * N_ZIPS} ZIP rows) per step (--reference-universes)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "rows/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "int32", "data": "synthetic",
        "config": workload_config(args.universes),
        "same_config_as_gpu_arm": U == args.universes,
        "cpu_baseline": {"value": value, "unit": "rows/s", "cores": cores, "kind": "port",
                         "sample": f"{sample}; C port of the serial-indices engine with its row loops split over {cores} OpenMP "
                                   f"threads (the Java engine itself is single-threaded); Java engine not timed: no JVM in image"},
        "e2e": {"value": value, "unit": "rows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ------------------------------------------------------------------------------------------------ GPU arm
def run_colq(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    from colq import _ffi
    from colq import geography as G
    from colq.device_data import build_geography_on_device, plymouth_colq_query
    from colq.engine import ColqContext

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa = bind_to_gpu_numa_node(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    stream = torch.cuda.Stream(device=dev)
    ctx = ColqContext(local_rank)
    ctx.set_stream(stream.cuda_stream)
    if world > 1:
        uid = torch.zeros(128, dtype=torch.uint8, device=dev)
        if rank == 0:
            uid.copy_(torch.frombuffer(bytearray(ctx.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(uid, 0)
        ctx.comm_init(bytes(uid.cpu().numpy().tobytes()), world, rank)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    U = args.universes * (world if args.weak else 1)
    base = G.load_base()
    gate = {"exact": False, "perturbed": None}
    if world > 1:
        gate["perturbed"] = perturbed_gate(ctx, rank, world, base)   # raises on any mismatch
        gate["perturbed_what"] = ("4*N+1 universes, every PLYMOUTH renamed except in the LAST rank's universes (the state mask exists on "
                                  "one rank only before the OR-exchange); sharded result == unsharded oracle, peer-memory and NCCL "
                                  "exchanges, raw C ABI and DataSystemColq.execute; then the same query with cities / zips split by plain row "
                                  "ranges and GLOBAL zip -> city keys (cross-shard bitmap all-gather)")
    geo = build_geography_on_device(ctx, U, world, rank, base=base, device=dev, sharded=world > 1, dict_names=args.dict_names)
    if args.dict_names:
        args.no_e2e = True
    q = plymouth_colq_query(ctx, lazy_fk=not args.eager)
    if args.unfused_root:
        q.set_option(_ffi.OPT_ROOT_FUSED, 0)

    # ---- correctness gate before any number: the result must be exactly {u * 29353 + r}
    from oracle_system import OracleDataSystem
    one = OracleDataSystem()
    G.register_geography(one, G.build_tables(1, base=base))
    one.execute(G.plymouth_query())
    rows1 = one.last_indices.astype(np.int64)
    one.close()
    want = (np.arange(U, dtype=np.int64)[:, None] * N_ZIPS + rows1[None, :]).reshape(-1)
    res = q.execute(want_indices=True, index_capacity=31 * U + 16)
    if res.count != 31 * U or not np.array_equal(res.indices.astype(np.int64), want):
        raise SystemExit(f"rank {rank}: GPU result differs from the oracle-derived expectation (count {res.count} vs {31 * U})")
    gate["exact"] = True
    r_plain = q.execute(want_indices=False)   # what one timed step launches (fetching the indices adds the gather's concatenation at N > 1)
    # (at N > 1 the fetch adds one launch of its own -- the concatenation of the gathered slots -- which is not part of a step)
    launches_per_step = int(r_plain.timing.kernel_launches) - (1 if world > 1 else 0)
    collectives_per_step = int(r_plain.timing.collectives)

    # ---- value: K steps, resident tables, CUDA events on the launching stream, max over ranks
    sampler = ClockSampler(local_rank)
    # The dominant launch is timed inside the timed region by a CUDA-event pair (COLQ_OPT_PROFILE=2) -- in every
    # HOT_EVERY-th step only: back-to-back executions are pipelined (COLQ_OPT_PIPELINE: the next step's string scan starts
    # while this step's root kernel drains), and an event between two steps keeps them apart, so a sampled step and its
    # successor run un-pipelined and the sample is the kernel's own duration.
    HOT_EVERY = 8

    def step(i):
        if i % HOT_EVERY == 0:
            q.set_option(_ffi.OPT_PROFILE, 2)
        elif i % HOT_EVERY == 1:
            q.set_option(_ffi.OPT_PROFILE, 0)
        q.execute_async()

    with torch.cuda.stream(stream):
        for i in range(max(args.warmup, 3)):
            step(i)
        q.profile_hot()  # drop the warm-up samples
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sampler.start()
        e0.record(stream)
        t_host = time.perf_counter()
        n_host = min(args.steps, 50)   # the first steps only: later ones may block on a full launch queue
        for i in range(args.steps):
            step(i)
            if i + 1 == n_host:
                host_us = (time.perf_counter() - t_host) * 1e6 / n_host   # verify + plan + launches; the GPU runs behind
        e1.record(stream)
        stream.synchronize()
        clocks = sampler.stop()
        barrier()
    hot_name, hot_ms, _hot_rows, hot_bytes, hot_samples = q.profile_hot()
    q.set_option(_ffi.OPT_PROFILE, 0)
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    ms_step = ms_total / args.steps
    res = q.fetch(want_indices=True, index_capacity=31 * U + 16)
    assert res.count == 31 * U and np.array_equal(res.indices.astype(np.int64), want)
    rows = U * N_ZIPS
    value = rows / (ms_step * 1e-3)

    # ---- roofline of the dominant kernel: per-launch CUDA events inside libcolq, separate profiled steps
    q.set_option(_ffi.OPT_PROFILE, 1)
    acc = {}
    for _ in range(args.profile_steps):
        q.execute(want_indices=False)
        for name, ms, r, b in q.profile():
            if ms >= 0:
                a = acc.setdefault(name, [0.0, 0, r, b])
                a[0] += ms
                a[1] += 1
    q.set_option(_ffi.OPT_PROFILE, 0)
    stages = {k: {"ms": v[0] / v[1], "rows": v[2], "algorithmic_bytes": v[3]} for k, v in acc.items() if v[1]}
    peak, peak_src = measured_peak()
    achieved = hot_bytes / (hot_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": hot_name, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": ncu_traffic(hot_name) if (U == 10_000 and world == 1 and not args.dict_names) else None,
                "peak_source": peak_src, "ms_per_launch": hot_ms,
                "algorithmic_bytes_per_launch": hot_bytes, "share_of_step": hot_ms / ms_step,
                "timed": f"CUDA events around this launch in every {HOT_EVERY}th timed step ({hot_samples} samples, COLQ_OPT_PROFILE=2), mean; "
                         "the other steps are pipelined behind their predecessor (COLQ_OPT_PIPELINE)"}

    # whole-query algorithmic bytes for the HBM GB/s half of BASELINE's metric
    algo_bytes = 4 * geo.n_zip_rows * 2 + 4 * (geo.n_city_rows + 1) + geo.name_bytes + 4 * geo.n_city_rows + 4 * 31 * geo.n_universes
    if args.dict_names:  # codes (4 B per city row) replace name offsets + bytes
        algo_bytes = 4 * geo.n_zip_rows * 2 + 4 * geo.n_city_rows * 2 + 4 * 31 * geo.n_universes
    algo_total = algo_bytes
    if world > 1:
        t = torch.tensor([float(algo_bytes)], dtype=torch.float64, device=dev)
        dist.all_reduce(t)
        algo_total = float(t.item())
    query_gbs = algo_total / (ms_step * 1e-3) / 1e9
    # bytes the launches really touch: the lazy chains skip all but a few sectors of the two FK columns (2.2 GB of the 6.58 GB
    # "every column read once" figure), so the physical rate is this one -- the algorithmic-equivalent above can exceed any peak
    touched = float(sum(v["algorithmic_bytes"] for v in stages.values()))
    if world > 1:
        t = torch.tensor([touched], dtype=torch.float64, device=dev)
        dist.all_reduce(t)
        touched = float(t.item())
    touched_gbs = touched / (ms_step * 1e-3) / 1e9

    # ---- e2e_resident: the public call with resident tables, matched indices read back every step
    q.execute(want_indices=True, index_capacity=31 * U + 16, pinned=True)   # untimed: allocates the pinned result buffer (cudaHostAlloc takes 5-300 ms on these hosts)
    barrier()
    with torch.cuda.stream(stream):
        e0.record(stream)
        n_res = max(5, min(args.steps, 50))
        for _ in range(n_res):
            r2 = q.execute(want_indices=True, index_capacity=31 * U + 16, pinned=True)
        e1.record(stream)
        stream.synchronize()
    ms_res = max_over_ranks(e0.elapsed_time(e1)) / n_res
    d2h_res = int(r2.timing.d2h_bytes)

    small = small_query_latency(base) if (world == 1 and not args.no_small_queries) else None

    # ---- e2e: HOST buffers in, matched indices out, every step (columns re-uploaded from pinned memory)
    e2e = e2e_upload = e2e_dict = None
    ingest = None
    if world == 1 and not args.no_ingest and not args.dict_names:
        ingest = measure_ingest(ctx, geo, base)
    if not args.no_e2e:
        host = {}
        for k, t in geo.tensors.items():
            h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
            h.copy_(t)
            host[k] = h.numpy()
        torch.cuda.synchronize(dev)
        nz, nc, nb = geo.n_zip_rows, geo.n_city_rows, geo.name_bytes
        sharded = world > 1
        place = _ffi.SHARDED if sharded else _ffi.REPLICATED
        h2d = 4 * nz * 3 + 4 * (nc + 1) + nb + 4 * nc + int(base["state_code_bytes"].size + base["state_name_bytes"].size) + 4 * 104 + 8 * 52 + 4 * 219
        # free the resident copy first so that both never coexist (keeps the e2e leg honest about allocation cost too)
        q.close()
        for tb in (geo.zips, geo.cities, geo.states):
            ctx.table_destroy(tb)
        geo.tensors.clear()
        ctx._keepalive.clear()
        torch.cuda.empty_cache()

        small_h2d = int(base["state_code_bytes"].size + base["state_name_bytes"].size) + 4 * 104 + 8 * 52 + 4 * 219

        # dictionary-encoded city names held by the host (what the shim keeps once the column has been encoded at load time)
        from colq.engine import encode_dictionary
        from colq.in_memory import StringColumn
        codes1, d_off, d_bytes, _vals = encode_dictionary(StringColumn(offsets=base["city_name_offsets"], data=base["city_name_bytes"]))
        hc = torch.empty(nc + 16, dtype=torch.int32, pin_memory=True)
        hc[:nc].copy_(torch.from_numpy(codes1).repeat(geo.n_universes))
        hc[nc:] = 0
        host["city_name_codes"] = hc.numpy()

        def e2e_step(upload: bool, dict_names: bool = False):
            """upload=False (the product's host path): the big columns stay in the pinned host buffers and are
            registered in place (colq_*_host); the query moves only what it touches over PCIe.
            upload=True: every column is copied to HBM first (colq_col_* / colq_associate_fk), then the query runs.
            dict_names: the host holds the city names dictionary-encoded (int32 codes + 16,584 distinct values)."""
            states = ctx.table_create(51, _ffi.REPLICATED, 0)
            cities = ctx.table_create(nc, place, geo.u0 * N_CITIES)
            zips = ctx.table_create(nz, place, geo.u0 * N_ZIPS)
            ctx.col_str(states, 0, base["state_code_offsets"], base["state_code_bytes"])
            ctx.col_str(states, 1, base["state_name_offsets"], base["state_name_bytes"])
            if upload:
                ctx.col_str(cities, 0, host["city_name_offsets"][: nc + 1].view(np.uint32), host["city_name_bytes"][:nb])
                ctx.associate_fk(cities, 1, states, 2, host["city_state"][:nc])
                ctx.col_i32(zips, 0, host["zip_code"][:nz])
                ctx.col_i32(zips, 1, host["zip_pop"][:nz])
                ctx.associate_fk(zips, 2, cities, 2, host["zip_city"][:nz])
            else:
                if dict_names:
                    ctx.col_str_dict_host(cities, 0, host["city_name_codes"], d_off, d_bytes, n=nc)
                else:
                    ctx.col_str_host(cities, 0, host["city_name_offsets"].view(np.uint32), host["city_name_bytes"], nc, nb)
                ctx.associate_fk_host(cities, 1, states, 2, host["city_state"], n=nc)
                ctx.col_i32_host(zips, 0, host["zip_code"], n=nz)
                ctx.col_i32_host(zips, 1, host["zip_pop"], n=nz)
                ctx.associate_fk_host(zips, 2, cities, 2, host["zip_city"], n=nz)
            ctx.associate_csr(states, 3, states, 4, base["adj_offsets"].astype(np.int64), base["adj_targets"])
            for name, tb in (("states", states), ("cities", cities), ("zips", zips)):
                ctx.register(name, tb)
            qq = plymouth_colq_query(ctx, lazy_fk=not args.eager)
            r = qq.execute(want_indices=True, index_capacity=31 * U + 16)
            qq.close()
            for tb in (zips, cities, states):
                ctx.table_destroy(tb)
            ctx._keepalive.clear()
            return r

        def time_e2e(upload: bool, dict_names: bool = False):
            for _ in range(2):  # warm-up: first touch of the pinned pages, and the device-buffer cache reaches its fixed point
                r3 = e2e_step(upload, dict_names)
            assert r3.count == 31 * U and np.array_equal(r3.indices.astype(np.int64), want)
            barrier()
            t0 = time.perf_counter()
            per_step = []
            with torch.cuda.stream(stream):
                e0.record(stream)
                for _ in range(args.e2e_steps):
                    t1 = time.perf_counter()
                    r3 = e2e_step(upload, dict_names)
                    per_step.append((time.perf_counter() - t1) * 1e3)
                e1.record(stream)
                stream.synchronize()
            wall_ms = (time.perf_counter() - t0) * 1e3 / args.e2e_steps
            ms = max_over_ranks(max(e0.elapsed_time(e1) / args.e2e_steps, wall_ms))
            assert r3.count == 31 * U and np.array_equal(r3.indices.astype(np.int64), want)
            return ms, r3, [round(x, 2) for x in per_step]

        ms_e2e, r3, steps_ms = time_e2e(upload=False)
        e2e = {"value": rows / (ms_e2e * 1e-3), "unit": "rows/s",
               "h2d_bytes_per_step": int(r3.timing.h2d_bytes) + small_h2d,
               "d2h_bytes_per_step": int(r3.timing.d2h_bytes), "ms_per_step": ms_e2e, "steps": args.e2e_steps,
               "ms_each_step_rank0": steps_ms,
               "what": "per step, per rank: colq_table_create, the big columns registered IN PLACE in pinned host memory "
                       "(colq_col_*_host / colq_associate_fk_host: nothing copied at registration), colq_execute -- the columns "
                       "the query scans in full (ZIP population, city-name offsets + bytes) are brought to HBM by the copy engine "
                       "ahead of their kernels, the lazily walked FK columns are read in place, single sectors over PCIe -- "
                       "matched indices read back, colq_table_destroy. h2d_bytes_per_step counts the fully scanned columns; the "
                       "never-touched ZIP-code column and all but ~0.5 M sectors of the FK columns (3.4 GB) do not cross PCIe"}
        ms_d, r5, steps_d = time_e2e(upload=False, dict_names=True)
        e2e_dict = {"value": rows / (ms_d * 1e-3), "unit": "rows/s", "h2d_bytes_per_step": int(r5.timing.h2d_bytes) + small_h2d + int(d_off.nbytes + d_bytes.nbytes),
                    "d2h_bytes_per_step": int(r5.timing.d2h_bytes), "ms_per_step": ms_d, "steps": args.e2e_steps, "ms_each_step_rank0": steps_d,
                    "what": "the same step when the host holds the city names dictionary-encoded (SURVEY 8f rank 2; int32 codes + 16,584 distinct "
                            "names, colq_col_str_dict_host): the query scans 4 B of code per city row instead of offsets + bytes"}
        ms_up, r4, steps_up = time_e2e(upload=True)
        e2e_upload = {"value": rows / (ms_up * 1e-3), "unit": "rows/s", "h2d_bytes_per_step": int(h2d),
                      "d2h_bytes_per_step": int(r4.timing.d2h_bytes), "ms_per_step": ms_up, "steps": args.e2e_steps,
                      "ms_each_step_rank0": steps_up,
                      "what": "same, but every column is first copied to HBM (colq_col_* / colq_associate_fk), touched or not"}

    # ---- CPU baseline beside it (rank 0, N=1 only): the serial port, like the serial reference engine
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        Uc = args.cpu_universes
        times = oracle_run(Uc, 1, 4)[1:]
        sec = statistics.mean(times)
        cpu = {"value": Uc * N_ZIPS / sec, "unit": "rows/s", "cores": 1, "kind": "port",
               "sample": f"{Uc} of {U} universes ({Uc * N_ZIPS} ZIP rows), 3 timed runs of orc_execute, single thread "
                         "(the reference engine is serial); Java engine not timed: no JVM in image",
               "ms_per_sample": sec * 1e3}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "rows/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak" if args.weak else "strong", "vs_baseline": None, "dtype": "int32",
            "data": "synthetic", "config": dict(workload_config(U), **({"city_names": "dictionary-encoded"} if args.dict_names else {})),
            "sharding": f"universe ranges over {world} rank(s); states replicated",
            "strategy": "lazy_fk_chain" if not args.eager else "materialise_all_nodes", "gate": gate,
            "hbm_gbs_touched": touched_gbs, "touched_bytes": touched,
            "hbm_gbs_query_algorithmic": query_gbs, "query_algorithmic_bytes": algo_total,
            "roofline": roofline, "stages_ms": {k: round(v["ms"], 5) for k, v in stages.items()},
            # every launch of the step against the same peak (separate profiled steps, one event pair per launch): algorithmic
            # bytes of the launch / its time.  root_fused additionally issues ~0.94 M random 32-byte key reads that its
            # algorithmic bytes do not show ("What bounds the root kernel")
            "stages_roofline": {k: {"algorithmic_bytes": int(v["algorithmic_bytes"]), "gbs": round(v["algorithmic_bytes"] / (v["ms"] * 1e-3) / 1e9, 1),
                                    "frac": round(v["algorithmic_bytes"] / (v["ms"] * 1e-3) / 1e9 / peak, 4)} for k, v in stages.items() if v["ms"] > 0},
            "cpu_baseline": cpu, "e2e": e2e, "e2e_dictionary": e2e_dict, "e2e_upload_all_columns": e2e_upload, "ingest": ingest,
            "e2e_resident": {"value": rows / (ms_res * 1e-3), "unit": "rows/s", "ms_per_step": ms_res, "h2d_bytes_per_step": 0,
                             "d2h_bytes_per_step": d2h_res, "what": "colq_execute with resident tables, matched indices read back into a pinned result buffer every step"},
            "small_query_latency": small,
            "host_enqueue_us_per_step": host_us, "host_numa": numa, "clocks": clocks, "gpu_launches": launches_per_step * args.steps, "gpu_launches_per_step": launches_per_step,
            "collectives_per_step": collectives_per_step,
        }
        emit(line)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def measure_ingest(ctx, geo, base):
    """Load-time work done by the GPU instead of host loops, timed on the full-size resident columns:
    dictionary encoding of the city-name column (colq_col_str_encode) and validation + None/One/Many classification of the
    zip -> city association shipped as a CSR (colq_associate_device).  Host wall clock around the (synchronous) calls."""
    import torch
    from colq import _ffi
    from colq.engine import encode_dictionary
    from colq.in_memory import StringColumn
    nc, nz = geo.n_city_rows, geo.n_zip_rows
    out = {}
    tt = geo.tensors["city_name_offsets"], geo.tensors["city_name_bytes"]
    best = None
    for _ in range(2):
        t = ctx.table_create(nc, _ffi.REPLICATED, 0)
        ctx.col_str_device(t, 0, tt[0].data_ptr(), tt[0].numel() * 4, tt[1].data_ptr(), tt[1].numel(), nc, geo.name_bytes, keepalive=tt)
        t0 = time.perf_counter()
        n_dict = ctx.col_str_encode(t, 0)
        ms = (time.perf_counter() - t0) * 1e3
        best = ms if best is None else min(best, ms)
        if _ == 0:   # parity with the host loop of round 1 (one universe's names have the same distinct values in the same order)
            _c, want_off, want_bytes, _v = encode_dictionary(StringColumn(offsets=base["city_name_offsets"], data=base["city_name_bytes"]))
            off, data = ctx.col_dict_str(t, 0)
            assert np.array_equal(off, want_off) and np.array_equal(data, want_bytes), "device dictionary differs from the host dictionary"
        ctx.table_destroy(t)
    read_bytes = 4 * (nc + 1) + geo.name_bytes
    out["dict_encode"] = {"ms": best, "rows": nc, "n_dict": int(n_dict), "bytes_read_per_pass": read_bytes,
                          "gbs_two_passes": 2 * read_bytes / (best * 1e-3) / 1e9,
                          "what": "colq_col_str_encode on the resident city-name column: hash insert, byte-exact verification, first-appearance "
                                  "codes, dictionary gather (reads the column twice, writes 4 B of code per row)"}
    dev = geo.tensors["zip_city"].device
    off = torch.arange(nz + 1, dtype=torch.int64, device=dev)          # every ZIP has exactly one city: Association.One
    torch.cuda.synchronize(dev)                                        # (torch filled it on its own stream)
    best = None
    for _ in range(2):
        x = ctx.table_create(nz, _ffi.REPLICATED, 0)
        y = ctx.table_create(nc, _ffi.REPLICATED, 0)
        t0 = time.perf_counter()
        is_fk = ctx.associate_device(x, 0, y, 0, off.data_ptr(), geo.tensors["zip_city"].data_ptr(), nz, nz, keepalive=(off,))
        ms = (time.perf_counter() - t0) * 1e3
        best = ms if best is None else min(best, ms)
        assert is_fk
        ctx.table_destroy(x)
        ctx.table_destroy(y)
    del off
    out["assoc_classify"] = {"ms": best, "rows": nz, "edges": nz, "stored_as": "dense to-one",
                             "what": "colq_associate_device on zip -> city as a CSR: offsets order, target range and max degree in one pass, "
                                     "then the dense to-one column"}
    return out


def perturbed_gate(ctx, rank, world, base):
    """Correctness gate of the multi-GPU exchanges, run by every bench invocation at N > 1 before anything is timed.
    The exact-copy workload cannot catch a broken mask exchange (every rank's local state mask already equals the global
    one), so this runs the PERTURBED workload of SURVEY.md 8d: PLYMOUTH is renamed in all universes except the last
    rank's, the sharded engine must still return exactly the unsharded oracle's rows (E/ExecutionContext.java:100-122
    semantics) -- once over NVLink peer memory, once over NCCL, through the raw C ABI and through the public
    DataSystemColq.execute."""
    from colq import _ffi, QueryResult
    from colq import geography as G
    from colq.engine import DataSystemColq
    from oracle_system import OracleDataSystem
    Ug = 4 * world + 1
    oracle = OracleDataSystem()
    G.register_geography(oracle, G.build_tables(Ug, base=base))
    full = oracle.execute(G.plymouth_query())
    want = oracle.last_indices.copy()
    want_codes = np.sort(full.result_set.columns()[0].ints())
    oracle.close()
    geo = G.build_tables(Ug, n_ranks=world, rank=rank, base=base, rename_plymouth_except_last_rank=True)
    lo, hi = geo.zip_row_base, geo.zip_row_base + geo.zips.size()
    for peer in (1, 0):
        ds = DataSystemColq(context=ctx, options={_ffi.OPT_PEER_EXCHANGE: peer})
        G.register_geography(ds, geo, sharded=True)
        ds._sync_tables()
        cq, why = ds._translate(G.plymouth_query())
        assert cq is not None, why
        res = cq.execute(want_indices=True, index_capacity=64)
        if res.count != want.shape[0] or not np.array_equal(res.indices, want):
            raise SystemExit(f"rank {rank}: perturbed multi-GPU gate FAILED (peer={peer}): {res.count} rows vs {want.shape[0]}")
        cq.close()
        # the public call: this rank's rows of the result table
        got = ds.execute(G.plymouth_query())
        if not isinstance(got, QueryResult.Success):
            raise SystemExit(f"rank {rank}: perturbed gate, DataSystemColq.execute failed: {got}")
        mine = want[(want >= lo) & (want < hi)]
        if got.result_set.size() != mine.shape[0]:
            raise SystemExit(f"rank {rank}: perturbed gate, DataSystemColq.execute returned {got.result_set.size()} local rows, want {mine.shape[0]}")
        ds.last_query.close()
        ds.last_query = None
        for h in set(ds._handles.values()):
            ctx.table_destroy(h)
    # cross-shard hops: cities / zips split by plain row ranges, zip -> city keys global
    from colq.device_data import plymouth_colq_query, register_cross_shard_geography
    handles = register_cross_shard_geography(ctx, G.build_tables(Ug, base=base), world, rank, base=base)
    cq = plymouth_colq_query(ctx)
    for _ in range(2):
        res = cq.execute(want_indices=True, index_capacity=want.shape[0] + 8)
        if res.count != want.shape[0] or not np.array_equal(res.indices, want):
            raise SystemExit(f"rank {rank}: cross-shard multi-GPU gate FAILED: {res.count} rows vs {want.shape[0]}")
    cq.close()
    for h in handles:
        ctx.table_destroy(h)
    return True


def small_query_latency(base):
    """BASELINE configs[0] and configs[2] on the GPU: 29k / 51 rows are launch-latency bound, so report microseconds
    per colq_execute call (host wall clock, result read back), not a roofline fraction."""
    from colq.device_data import build_geography_on_device, north_south_north_colq_query, plymouth_colq_query
    from colq.engine import ColqContext
    ctx = ColqContext(0)
    build_geography_on_device(ctx, 1, base=base)
    out = {}
    for name, make, want in (("plymouth_1_universe", plymouth_colq_query, 31), ("north_south_north", north_south_north_colq_query, 2)):
        q = make(ctx)
        for _ in range(20):
            r = q.execute(want_indices=True)
        assert r.count == want
        ts = []
        for _ in range(300):
            t0 = time.perf_counter()
            q.execute(want_indices=True)
            ts.append((time.perf_counter() - t0) * 1e6)
        out[name] = {"median_us": statistics.median(ts), "p10_us": sorted(ts)[30], "kernel_launches": int(r.timing.kernel_launches)}
        q.close()
    ctx.close()
    return out


def run_single_table(args, rank=0, local_rank=0, world=1):
    """int_scan / str_eq: one column, one predicate (BASELINE configs[1] and configs[4]).  str_eq also runs sharded by
    row range over N ranks (configs[4] as written: 500 M rows across 8 B200) with the final index gather."""
    import torch
    import torch.distributed as dist
    from colq import _ffi
    from colq import geography as G
    from colq.device_data import build_int_scan_on_device, build_name_scan_on_device
    from colq.engine import ColqContext
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    stream = torch.cuda.Stream(device=dev)
    ctx = ColqContext(local_rank)
    ctx.set_stream(stream.cuda_stream)
    if world > 1:
        uid = torch.zeros(128, dtype=torch.uint8, device=dev)
        if rank == 0:
            uid.copy_(torch.frombuffer(bytearray(ctx.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(uid, 0)
        ctx.comm_init(bytes(uid.cpu().numpy().tobytes()), world, rank)

    def allsum(x):
        if world == 1:
            return x
        t = torch.tensor([float(x)], dtype=torch.float64, device=dev)
        dist.all_reduce(t)
        return int(round(t.item()))

    def allmax(x):
        if world == 1:
            return x
        t = torch.tensor([float(x)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    base = G.load_base()
    if args.workload == "int_scan":
        n = args.rows or 1_000_000_000
        _t, col = build_int_scan_on_device(ctx, n, base=base)
        q = ctx.query("ints")
        q.criteria_i32_range(0, 0, 10_000, 10_099)
        v = col[:n]
        expect = int(((v >= 10_000) & (v <= 10_099)).sum().item())
        workload = {"workload": "int_range_scan_plus_compaction", "source": "BASELINE.json configs[1]", "rows": n,
                    "predicate": "[10000, 10099]", "generator": "pops[splitmix64(42, i) mod 29353]"}
        metric = "int_range_scan_rows_per_sec"
        algo = lambda m: 4 * n + 4 * m  # noqa: E731  
    else:
        n = args.rows or (62_500_000 if world == 1 else 500_000_000)   # total rows; each rank holds a contiguous range
        lo, hi = n * rank // world, n * (rank + 1) // world
        _t, off32, data, idx, total = build_name_scan_on_device(ctx, hi - lo, base=base, start=lo,
                                                                placement=_ffi.SHARDED if world > 1 else _ffi.REPLICATED)
        q = ctx.query("names")
        q.criteria_str(0, 0, 0, b"PLYMOUTH")
        names = [bytes(base["city_name_bytes"][base["city_name_offsets"][i]:base["city_name_offsets"][i + 1]]) for i in range(G.N_CITIES)]
        ply = torch.tensor([i for i, s_ in enumerate(names) if s_ == b"PLYMOUTH"], device=dev, dtype=torch.int32)
        expect = allsum(int(torch.isin(idx, ply).sum().item()))
        total = allsum(total)
        workload = {"workload": "city_name_equality_scan",
                    "source": "BASELINE.json configs[4]" + (" (one GPU's shard of 500M rows)" if world == 1 and n == 62_500_000 else ""),
                    "rows": n, "name_bytes": total, "generator": "cityNames[splitmix64(42, i) mod 25701]",
                    "sharding": f"contiguous row ranges over {world} rank(s), final index gather over NVLink peer memory" if world > 1 else "none"}
        metric = "string_equality_scan_rows_per_sec"
        algo = lambda m: 4 * (n + 1) + total + n // 8  # noqa: E731  
    if args.unfused_root:
        q.set_option(_ffi.OPT_ROOT_FUSED, 0)
    res = q.execute(want_indices=True, index_capacity=max(expect, 1) + 16)
    if res.count != expect or (res.indices.shape[0] > 1 and not np.all(np.diff(res.indices.astype(np.int64)) > 0)):
        raise SystemExit(f"GPU count {res.count} != independent expectation {expect} (or indices not ascending)")
    launches = int(res.timing.kernel_launches)
    sampler = ClockSampler(local_rank)
    q.set_option(_ffi.OPT_PROFILE, 2)
    with torch.cuda.stream(stream):
        for _ in range(max(args.warmup, 3)):
            q.execute_async()
        q.profile_hot()
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sampler.start()
        e0.record(stream)
        for _ in range(args.steps):
            q.execute_async()
        e1.record(stream)
        stream.synchronize()
        clocks = sampler.stop()
    ms_step = allmax(e0.elapsed_time(e1)) / args.steps
    hot_name, hot_ms, _hot_rows, hot_bytes, hot_samples = q.profile_hot()
    q.set_option(_ffi.OPT_PROFILE, 1)
    acc = {}
    for _ in range(args.profile_steps):
        q.execute(want_indices=False)
        for name, ms, r, b in q.profile():
            if ms >= 0:
                a = acc.setdefault(name, [0.0, 0, r, b])
                a[0] += ms
                a[1] += 1
    stages = {k: {"ms": v[0] / v[1], "algorithmic_bytes": v[3]} for k, v in acc.items() if v[1]}
    peak, peak_src = measured_peak()
    achieved = hot_bytes / (hot_ms * 1e-3) / 1e9
    line = {
        "metric": metric, "value": n / (ms_step * 1e-3), "unit": "rows/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong" if world > 1 else "weak", "vs_baseline": None,
        "dtype": "int32" if args.workload == "int_scan" else "u8", "data": "synthetic",
        "config": dict(workload, l2_policy="inputs larger than L2 (no flush)", matches=expect),
        "hbm_gbs_query_algorithmic": algo(expect) / (ms_step * 1e-3) / 1e9, "query_algorithmic_bytes": algo(expect),
        "roofline": {"bound": "hbm", "kernel": hot_name, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": None, "peak_source": peak_src, "ms_per_launch": hot_ms,
                     "algorithmic_bytes_per_launch": hot_bytes, "share_of_step": hot_ms / ms_step,
                     "timed": f"CUDA events around this launch in each of the {hot_samples} timed steps (COLQ_OPT_PROFILE=2), mean"},
        "stages_ms": {k: round(v["ms"], 5) for k, v in stages.items()},
        "cpu_baseline": None, "e2e": None, "clocks": clocks, "gpu_launches": launches * args.steps, "gpu_launches_per_step": launches,
    }
    if rank == 0:
        emit(line)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if args.workload != "plymouth":
        if world != 1 and args.workload != "str_eq":
            raise SystemExit("--workload int_scan is a single-GPU kernel benchmark")
        run_single_table(args, rank, local_rank, world)
        return
    if world != args.gpus:
        if args.gpus == 1 and world == 1:
            pass
        else:
            raise SystemExit(f"--gpus {args.gpus} needs WORLD_SIZE={args.gpus} (launch with torch.distributed.run); got {world}")
    run_colq(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
