#!/usr/bin/env python
"""bench.py — merged rows/s of the LSM merge hot path on B200 (BASELINE.json metric).

One "step" = one pass of the hot path (sampled partition -> plan -> scan -> emit) over one bucket of
synthetic sorted runs.  Workloads (BASELINE.json configs):
  c3 (default, the configuration the metric is quoted on): 16 runs x 6.25 M rows = 100 M rows,
      partial-update merge engine, 50-column wide row (pk + 20 BIGINT + 15 DOUBLE + 14 VARCHAR(8..24)),
      every non-pk cell NULL with p = 0.5
  c2: 8 runs x 12.5 M rows = 100 M rows, deduplicate, BIGINT pk + 10 BIGINT columns
  c1: 2 runs x 500 K rows, deduplicate, BIGINT pk + BIGINT value

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3|c2|c1] [--rows R]
    python bench.py --impl reference ...     # the reference algorithm on the host cores (CPU)

`value`     whole-job merged (= input) rows/s with the runs already resident in HBM.
`e2e`       same metric through the public reader API with HOST buffers: every step copies the runs
            host->device (pinned memory) and the merged batch device->host.
`roofline`  achieved HBM GB/s of the dominant kernel (emit) on the algorithmic bytes
            N_in*B + N_out*B, against MEASURED_PEAKS.json.
`cpu_baseline`  the oracle (C restatement of LoserTree + MergeFunction) timed on this box's host cores.
Multi-GPU: one process per GPU (torchrun); buckets are independent, so every rank merges its own
bucket and there is no data-path collective ("weak" scaling).
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

import numpy as np  # noqa: E402

WORKLOADS = {
    "c3": dict(n_runs=16, rows=100_000_000, engine="partial-update", null_prob=0.5,
               desc="16-run partial-update, 50-col wide row (pk+20 i64+15 f64+14 varchar), 100M rows"),
    # SURVEY §8d "C3-agg": same rows as C3, merge-engine aggregation: the 15 doubles and 20 bigints use `sum`
    # (ordered left fold, bit-exact), the strings last_non_null_value
    "c3agg": dict(n_runs=16, rows=100_000_000, engine="aggregate", null_prob=0.5,
                  desc="16-run aggregation (sum over 20 i64 + 15 f64, last_non_null_value over 14 varchar), 50-col wide "
                       "row, 100M rows"),
    "c2": dict(n_runs=8, rows=100_000_000, engine="deduplicate", null_prob=0.0,
               desc="8-run deduplicate, int64 pk + 10 int64 cols, 100M rows"),
    "c1": dict(n_runs=2, rows=1_000_000, engine="deduplicate", null_prob=0.0,
               desc="2-run deduplicate, int64 pk + int64 val, 1M rows"),
    # one bucket of the full-compaction config: string key, deletes, drop-delete, then the
    # merged batch is re-encoded to Parquet on the device (reported under "rewrite")
    "c4": dict(n_runs=32, rows=16_000_000, engine="deduplicate", null_prob=0.5, delete_prob=0.05, drop_delete=True,
               desc="one bucket of a full compaction rewrite: 32 runs x 500K rows, varchar(16) pk + 4 i64 + 2 f64 + "
                    "2 i32 + 3 varchar, 5% deletes, drop-delete, output re-encoded to Parquet"),
}


def schema_c4():
    from paimon_b200.types import DataField, KeyValueSchema, RowType
    fields = [DataField("pk", "VARCHAR(16)", False)]
    fields += [DataField(f"i{i}", "BIGINT", True) for i in range(4)]
    fields += [DataField(f"d{i}", "DOUBLE", True) for i in range(2)]
    fields += [DataField(f"n{i}", "INT", True) for i in range(2)]
    fields += [DataField(f"s{i}", "VARCHAR(64)", True) for i in range(3)]
    return KeyValueSchema.of(RowType(tuple(fields)), ["pk"])


def make_schema(workload):
    from paimon_b200 import datagen
    return {"c1": datagen.schema_c1, "c2": datagen.schema_c2, "c3": datagen.schema_c3, "c3agg": datagen.schema_c3,
            "c4": schema_c4}[workload]()


def make_spec(workload, schema):
    from paimon_b200.merge_function import (AggregateMergeFunction, DeduplicateMergeFunction,
                                            PartialUpdateMergeFunction)
    if WORKLOADS[workload]["engine"] == "aggregate":
        opts = {f"fields.{f.name}.aggregate-function": "sum" for f in schema.value_type.fields
                if f.name != "pk" and f.physical.name in ("INT64", "DOUBLE")}
        return AggregateMergeFunction.factory(opts, schema.value_type, ["pk"]).create()
    if WORKLOADS[workload]["engine"] == "partial-update":
        return PartialUpdateMergeFunction.factory({}, schema.value_type, ["pk"]).create()
    spec = DeduplicateMergeFunction.factory().create()
    return spec.with_drop_delete() if WORKLOADS[workload].get("drop_delete") else spec


# ------------------------------------------------------------------ device-side synthetic runs

def _splitmix64(x):
    import torch
    x = x + (-7046029254386353131)                       # 0x9E3779B97F4A7C15 as int64
    x = (x ^ ((x >> 30) & ((1 << 34) - 1))) * (-4658895280553007687)   # 0xBF58476D1CE4E5B9
    x = (x ^ ((x >> 27) & ((1 << 37) - 1))) * (-7723592293110705685)   # 0x94D049BB133111EB
    return x ^ ((x >> 31) & ((1 << 33) - 1))


def _hex_keys(keys, dev):
    """int64 keys -> 16-character lower-case hex strings (big endian: string order == integer order)."""
    import torch
    sh = torch.arange(60, -4, -4, device=dev, dtype=torch.int64)
    nib = ((keys[:, None] >> sh) & 15).to(torch.uint8)
    data = torch.where(nib < 10, nib + 48, nib + 87).flatten()
    data = torch.cat([data, torch.zeros(16, device=dev, dtype=torch.uint8)]).contiguous()
    offs = (torch.arange(keys.numel() + 1, device=dev, dtype=torch.int64) * 16).to(torch.int32).contiguous()
    return data, offs


def gen_device_run(schema, run_index, n, key_space, null_prob, seed, dev, delete_prob=0.0):
    """One sorted run generated directly in HBM.  Returns (columns, keepalive tensors, key tensor)."""
    import torch
    from paimon_b200.sort_merge_reader import DeviceColumn
    from paimon_b200.types import PhysicalType
    g = torch.Generator(device=dev)
    g.manual_seed(seed * 1000003 + run_index)
    keys = torch.randperm(key_space, device=dev, generator=g)[:n].sort().values.contiguous()
    keep = [keys]
    cols = []
    string_key = schema.key_type.fields[0].physical == PhysicalType.STRING
    if string_key:
        kdata, koffs = _hex_keys(keys, dev)
        keep += [kdata, koffs]
        key_col = DeviceColumn(kdata.data_ptr(), koffs.data_ptr())
        key_bytes = n * 16 + 4 * (n + 1)
    else:
        key_col = DeviceColumn(keys.data_ptr())
        key_bytes = n * 8
    for _ in schema.key_type.fields:
        cols.append(key_col)
    seq = (torch.arange(n, device=dev, dtype=torch.int64) + (run_index << 32)).contiguous()
    kind = torch.zeros(n, device=dev, dtype=torch.int8)
    if delete_prob > 0:
        kind[torch.rand(n, device=dev, generator=g) < delete_prob] = 3
    keep += [seq, kind]
    cols += [DeviceColumn(seq.data_ptr()), DeviceColumn(kind.data_ptr())]
    nbytes = key_bytes + seq.numel() * 8 + kind.numel()
    pk_names = {f.name[len("_KEY_"):] for f in schema.key_type.fields}
    for ci, f in enumerate(schema.value_type.fields):
        t = f.physical
        if f.name in pk_names:
            cols.append(key_col)
            nbytes += key_bytes
            continue
        h = _splitmix64(keys ^ ((run_index + 1) * 0x100 + ci << 40))
        valid_ptr = 0
        bits = None
        if f.nullable and null_prob > 0:
            assert null_prob == 0.5, "device generator draws validity bits with p = 0.5"
            vbytes = torch.randint(0, 256, ((n + 7) // 8 + 8,), device=dev, dtype=torch.uint8, generator=g)
            keep.append(vbytes)
            valid_ptr = vbytes.data_ptr()
            nbytes += (n + 7) // 8
            if t in (PhysicalType.STRING, PhysicalType.BINARY):
                sh = torch.arange(8, device=dev, dtype=torch.uint8)
                bits = ((vbytes[:, None] >> sh) & 1).flatten()[:n].to(torch.int64)
        if t == PhysicalType.INT64:
            keep.append(h)
            cols.append(DeviceColumn(h.data_ptr(), 0, valid_ptr))
            nbytes += n * 8
        elif t == PhysicalType.INT32:
            v = (h & 0x7fffffff).to(torch.int32).contiguous()
            keep.append(v)
            cols.append(DeviceColumn(v.data_ptr(), 0, valid_ptr))
            nbytes += n * 4
        elif t == PhysicalType.DOUBLE:
            d = ((h >> 11) & ((1 << 53) - 1)).to(torch.float64) * (2000.0 / (1 << 53)) - 1000.0
            keep.append(d)
            cols.append(DeviceColumn(d.data_ptr(), 0, valid_ptr))
            nbytes += n * 8
        elif t in (PhysicalType.STRING, PhysicalType.BINARY):
            lens = 8 + ((h >> 3) & 0xffff) % 17                       # U[8, 24]
            if bits is not None:
                lens = lens * bits                                     # NULL cells carry no payload
            offs = torch.zeros(n + 1, device=dev, dtype=torch.int64)
            torch.cumsum(lens, 0, out=offs[1:])
            total = int(offs[-1].item())
            offs32 = offs.to(torch.int32)
            data = torch.randint(48, 112, (max(total, 1) + 16,), device=dev, dtype=torch.uint8, generator=g)
            keep += [offs32, data]
            cols.append(DeviceColumn(data.data_ptr(), offs32.data_ptr(), valid_ptr))
            nbytes += total + 4 * (n + 1)
            del lens, offs, bits
        else:
            raise ValueError(f"bench generator: unsupported type {t}")
    return cols, keep, keys, nbytes, kind


def device_runs(workload, schema, rows, dev, seed):
    import torch
    from paimon_b200.sort_merge_reader import SortedRunReader
    w = WORKLOADS[workload]
    n_runs = w["n_runs"]
    per_run = rows // n_runs
    key_space = max(rows // 2, per_run)
    readers, all_keys, all_kinds, in_bytes = [], [], [], 0
    for r in range(n_runs):
        cols, keep, keys, nb, kind = gen_device_run(schema, r, per_run, key_space, w["null_prob"], seed, dev,
                                                    w.get("delete_prob", 0.0))
        readers.append(SortedRunReader.from_device(schema, per_run, cols, keepalive=keep))
        all_keys.append(keys)
        all_kinds.append(kind)
        in_bytes += nb
    torch.cuda.synchronize()
    return readers, all_keys, in_bytes, all_kinds


# ------------------------------------------------------------------ clocks sampling

class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.samples = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            p = [x.strip() for x in s.split(",")]
            if len(p) < 6:
                continue
            try:
                sm.append(float(p[0])); mx.append(float(p[1]))
            except ValueError:
                continue
            for nm, v in zip(names, p[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------ CPU baseline (oracle)

def cpu_baseline(workload, total_sample_rows, threads, steps=1, seed=7):
    """The reference algorithm (oracle port) on the host cores: `threads` independent buckets, one thread
    per bucket exactly like the reference's one-thread-per-split readers."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import pyoracle
    from paimon_b200 import datagen
    schema = make_schema(workload)
    spec = make_spec(workload, schema)
    w = WORKLOADS[workload]
    per_bucket = max(total_sample_rows // threads, w["n_runs"] * 64)
    n_distinct = min(threads, 8)        # distinct synthetic buckets; threads beyond that re-merge a copy's inputs
    buckets = [datagen.make_runs(schema, w["n_runs"], per_bucket, seed=seed + b, null_prob=w["null_prob"],
                                 delete_prob=w.get("delete_prob", 0.0))
               for b in range(n_distinct)]
    prepared = [pyoracle.prepare(schema, spec, buckets[b % n_distinct]) for b in range(threads)]

    def work(b):
        return pyoracle.run_prepared(prepared[b])          # C call only; the GIL is released

    times, outs = [], []
    with ThreadPoolExecutor(max_workers=threads) as ex:
        for _ in range(steps):
            t0 = time.perf_counter()
            outs = list(ex.map(work, range(threads)))
            times.append(time.perf_counter() - t0)
    rows = per_bucket // w["n_runs"] * w["n_runs"] * threads
    return rows, times, sum(outs)


# ------------------------------------------------------------------ main

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--rows", type=int, default=None, help="override total input rows per GPU")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-range-rows", type=int, default=8 << 20,
                    help="e2e: input rows per key range of the streaming reader (0 = one batch, no overlap)")
    ap.add_argument("--e2e-depth", type=int, default=3, help="e2e: key ranges in flight")
    ap.add_argument("--e2e-frac", type=float, default=0.0,
                    help="e2e: fraction of the key space to stream (0 = all of it if page-locked memory allows)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample-rows", type=int, default=None)
    ap.add_argument("--cpu-threads", type=int, default=None)
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    w = WORKLOADS[args.workload]
    rows = args.rows or w["rows"]
    metric = "merged rows/sec at 16 runs x 100M rows" if args.workload == "c3" else f"merged rows/sec ({args.workload})"
    config = {"workload": f"{args.workload}: {w['desc']}", "rows_per_gpu": rows, "n_runs": w["n_runs"],
              "merge_engine": w["engine"], "buckets_per_gpu": 1, "parallelism": f"bucket-per-gpu x{world}",
              "l2": "inputs (>50 GB) far exceed the 126 MB L2; no explicit flush" if rows >= 10_000_000
                    else "small input: L2-resident (not a headline configuration)"}

    # ---------------- reference arm: the reference's CPU algorithm on the host cores
    if args.impl == "reference":
        if rank != 0:
            return
        threads = args.cpu_threads or min(os.cpu_count() or 1, 64)
        sample = args.cpu_sample_rows or threads * (250_000 if args.workload == "c3" else 1_000_000)
        sample = min(sample, rows)
        cpu_baseline(args.workload, min(sample, 200_000), threads, steps=max(args.warmup, 1) if args.warmup else 0)
        nrows, times, _ = cpu_baseline(args.workload, sample, threads, steps=args.steps)
        total = sum(times)
        val = nrows * len(times) / total
        line = {"impl": "reference", "metric": metric, "value": val, "unit": "rows/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int64",
                "data": "synthetic", "config": config,
                "cpu_baseline": {"value": val, "unit": "rows/s", "cores": threads, "kind": "port",
                                 "sample": f"{threads} buckets x {nrows // threads} rows of the same shape, one "
                                           f"thread per bucket (C restatement of LoserTree+MergeFunction; no JVM in the image)"},
                "e2e": {"value": val, "unit": "rows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    # ---------------- B200 arm
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device: the merge path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from paimon_b200 import _native as N
    from paimon_b200.columnar import Column, KeyValueBatch
    from paimon_b200.sort_merge_reader import RangeStreamingMergeReader, SortedRunReader, SortMergeReader

    schema = make_schema(args.workload)
    spec = make_spec(args.workload, schema)
    N.init(local_rank)
    readers, all_keys, in_bytes, all_kinds = device_runs(args.workload, schema, rows, dev, seed=100 + rank)
    n_in = sum(r.n_rows for r in readers)
    rd = SortMergeReader.create_sort_merge_reader(readers, None, None, spec, device=local_rank)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        rd.execute()
    st = rd.stats()
    n_out = st.rows_out
    # sanity at full size (size-independent properties): row conservation + strictly increasing keys
    if w.get("drop_delete"):
        # the newest record of a key wins (sequence = run << 32 | row); keys whose winner is a DELETE drop out
        cat_k = torch.cat(all_keys)
        cat_r = torch.cat([torch.full_like(k, r) for r, k in enumerate(all_keys)])
        cat_d = torch.cat(all_kinds).to(torch.int64)
        order = torch.argsort(cat_k * 64 + cat_r)
        sk, sd = cat_k[order], cat_d[order]
        last = torch.ones_like(sk, dtype=torch.bool)
        last[:-1] = sk[1:] != sk[:-1]
        uniq = int((last & (sd == 0)).sum().item())
        del cat_k, cat_r, cat_d, order, sk, sd, last
    else:
        uniq = torch.unique(torch.cat(all_keys)).numel()
    assert n_out == uniq, f"merged rows {n_out} != expected rows {uniq}"
    out_bytes = st.bytes_out

    ext = torch.cuda.ExternalStream(rd.cuda_stream(), device=dev)
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ms_emit = ms_plan = ms_part = ms_tot = ms_alloc = 0.0
    launches = 0
    t0 = time.perf_counter()
    e0.record(ext)
    for _ in range(args.steps):
        rd.execute()
        s = rd.stats()
        ms_emit += s.ms_emit; ms_plan += s.ms_plan; ms_part += s.ms_partition; ms_tot += s.ms_total
        ms_alloc += s.ms_alloc
        launches += s.launches
    e1.record(ext)
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop()
    dev_ms = e0.elapsed_time(e1)
    t = torch.tensor([dev_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    step_ms = float(t.item()) / args.steps
    value = world * n_in / (step_ms * 1e-3)

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_kind = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s"
    alg_bytes = in_bytes + out_bytes
    emit_ms = ms_emit / args.steps
    # DRAM traffic of the dominant kernel from the committed ncu capture of this workload at its full size
    # (profiles/traffic.json; only meaningful for the default row count)
    traffic = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        if args.workload in tj and rows == w["rows"]:
            traffic = tj[args.workload]["traffic"]
    except Exception:
        pass
    achieved = alg_bytes / (emit_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": "k_emit", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "peak_source": peak_kind, "traffic": traffic,
                "algorithmic_bytes": alg_bytes, "kernel_ms": emit_ms,
                "step_frac": alg_bytes / (step_ms * 1e-3) / 1e9 / peak,
                "phase_ms": {"partition": ms_part / args.steps, "plan+scan": ms_plan / args.steps,
                             "size_readback+alloc": ms_alloc / args.steps, "emit": emit_ms,
                             "device_total": ms_tot / args.steps}}

    # ---------------- compaction rewrite: encode the merged batch to Parquet on the device (C4)
    rewrite = None
    if w.get("drop_delete"):
        import ctypes as C
        from paimon_b200.compact_rewriter import file_column_names
        names = file_column_names(schema)
        arr = (C.c_char_p * len(names))(*[nm.encode() for nm in names])
        lib = N.load()
        enc_ms, fbytes, pages = [], 0, 0
        for _ in range(3):
            fh = C.c_uint64(0)
            N.check(lib.pg_parquet_encode(rd._merge_h, arr, 0, -1, None, C.byref(fh)))
            fm = N.PgFileMeta()
            N.check(lib.pg_parquet_file_meta(fh.value, C.byref(fm)))
            enc_ms.append(float(fm.ms_encode)); fbytes = int(fm.file_bytes); pages = int(fm.n_pages)
            lib.pg_parquet_file_free(fh.value)
        em = min(enc_ms)
        rewrite = {"encode_ms": em, "file_bytes": fbytes, "pages": pages, "encode_GBps": fbytes / (em * 1e-3) / 1e9,
                   "merge_plus_encode_rows_per_s": n_in / ((step_ms + em) * 1e-3)}

    # ---------------- e2e: host buffers in, host batch out, through the public reader API
    e2e = None
    if not args.no_e2e:
        ftypes = schema.physical_types()
        # device -> pinned host copies of every input buffer (the step's inputs live in page-locked memory)
        # Page-locked host memory is finite and every rank of the box needs its own copy of the inputs and room for
        # the outputs: when that does not fit comfortably, the end-to-end leg streams a key-range PREFIX of the
        # bucket (the first e2e_frac of the key space; same runs, same shape) and reports rows/s on it.
        local_world = int(os.environ.get("LOCAL_WORLD_SIZE", world))
        e2e_frac = args.e2e_frac
        if e2e_frac <= 0:
            try:
                avail = next(int(l.split()[1]) * 1024 for l in open("/proc/meminfo") if l.startswith("MemAvailable"))
            except Exception:
                avail = 256 << 30
            need = (in_bytes + out_bytes * 1.05) * local_world
            e2e_frac = 1.0 if need <= 0.5 * avail else max(0.05, 0.5 * avail / need)
        hi_rows = [r.n_rows for r in readers]
        e2e_rows_in, e2e_rows_out = n_in, n_out
        if e2e_frac < 1.0 and schema.n_key == 1 and not w.get("drop_delete"):
            cut = all_keys[0][min(int(e2e_frac * all_keys[0].numel()), all_keys[0].numel() - 1)]
            hi_rows = [int(torch.searchsorted(k_, cut).item()) for k_ in all_keys]
            hi_rows = [max(h, 1) for h in hi_rows]
            e2e_rows_in = sum(hi_rows)
            e2e_rows_out = torch.unique(torch.cat([k_[:h] for k_, h in zip(all_keys, hi_rows)])).numel()
        else:
            e2e_frac = 1.0
        host_runs = []
        e2e_out_bytes = int(out_bytes * (e2e_rows_in / max(n_in, 1)) * 1.1) if e2e_frac < 1.0 else int(out_bytes)
        for r, hi in zip(readers, hi_rows):
            cols = []
            byptr = {tt.data_ptr(): tt for tt in r.keepalive}
            cache = {}

            def to_host(ptr, count=None, byptr=byptr, cache=cache):
                if not ptr:
                    return None
                if ptr not in cache:
                    tt = byptr[ptr]
                    if count is not None:
                        tt = tt[:count]
                    hb = torch.empty(tt.shape, dtype=tt.dtype, pin_memory=True)
                    hb.copy_(tt)
                    cache[ptr] = hb.numpy()
                return cache[ptr]
            for ci, dc in enumerate(r.device_columns):
                t_ = ftypes[ci]
                offs = to_host(dc.offsets, hi + 1)
                if offs is not None:
                    data = to_host(dc.data, int(offs[hi]) + 16).view(np.uint8)
                else:
                    data = to_host(dc.data, hi)
                val = to_host(dc.validity, (hi + 7) // 8 + 8)
                cols.append(Column(t_, data[:hi] if offs is None else data, offs, val))
            host_runs.append(KeyValueBatch(schema, cols))
        torch.cuda.synchronize()
        rd.close()
        del readers, all_keys
        torch.cuda.empty_cache()
        arena = torch.empty(int(e2e_out_bytes * 1.02) + (64 << 20), dtype=torch.uint8, pin_memory=True)
        arena_np = arena.numpy()
        e2e_times, h2d_b, d2h_b = [], 0, 0
        import threading
        single_key = schema.n_key == 1
        for it_ in range(args.e2e_steps + 1):
            top = [0]
            lock = threading.Lock()

            def alloc(nbytes):
                with lock:
                    a = (top[0] + 63) & ~63
                    top[0] = a + nbytes
                return arena_np[a:a + nbytes]
            barrier()
            t0 = time.perf_counter()
            rows_out = 0
            if single_key and args.e2e_range_rows > 0:
                # batches of key ranges: H2D of range i+1 | merge of range i | D2H of range i-1
                mr = RangeStreamingMergeReader(schema, host_runs, spec, target_rows=args.e2e_range_rows,
                                               depth=args.e2e_depth, device=local_rank,
                                               allocator_factory=lambda: alloc)
                while True:
                    out = mr.read_batch()
                    if out is None:
                        break
                    rows_out += out.n_rows
                torch.cuda.synchronize()
                dt = time.perf_counter() - t0
                h2d_b, d2h_b = mr.bytes_h2d, mr.bytes_d2h
                mr.close()
            else:
                hr = [SortedRunReader(schema, b) for b in host_runs]
                mr = SortMergeReader.create_sort_merge_reader(hr, None, None, spec, device=local_rank)   # H2D
                mr.execute()
                out = mr.fetch(allocator=alloc)                                                           # D2H
                torch.cuda.synchronize()
                dt = time.perf_counter() - t0
                s = mr.stats()
                h2d_b, d2h_b = s.bytes_h2d, s.bytes_d2h
                rows_out = out.n_rows
                mr.close()
            assert rows_out == e2e_rows_out, (rows_out, e2e_rows_out)
            if it_ > 0:
                e2e_times.append(dt)
        tt = torch.tensor([sum(e2e_times) / len(e2e_times)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e = {"value": world * e2e_rows_in / float(tt.item()), "unit": "rows/s", "h2d_bytes_per_step": int(h2d_b),
               "d2h_bytes_per_step": int(d2h_b), "ms_per_step": 1e3 * float(tt.item()), "steps": len(e2e_times),
               "rows_in_per_step": int(e2e_rows_in), "rows_out_per_step": int(e2e_rows_out),
               "sample": ("the whole bucket" if e2e_frac >= 1.0 else
                          f"key-range prefix of the bucket ({e2e_frac:.2f} of the key space): page-locked host memory "
                          f"for {local_world} ranks' full inputs + outputs was not available"),
               "api": ("RangeStreamingMergeReader(host runs).read_batch() loop over the C ABI: key ranges of "
                       f"~{args.e2e_range_rows} rows, {args.e2e_depth} in flight (H2D | merge | D2H overlap)")
               if single_key and args.e2e_range_rows > 0 else
               "SortMergeReader.create_sort_merge_reader(host runs).execute()+fetch() over the C ABI"}
    else:
        rd.close()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = args.cpu_threads or min(os.cpu_count() or 1, 64)
        sample = args.cpu_sample_rows or threads * (250_000 if args.workload == "c3" else 1_000_000)
        sample = min(sample, rows)
        nrows, times, _ = cpu_baseline(args.workload, sample, threads, steps=2)
        cpu = {"value": nrows / min(times), "unit": "rows/s", "cores": threads, "kind": "port",
               "sample": f"{threads} buckets x {nrows // threads} rows of the same shape, one thread per bucket; "
                         f"oracle = C restatement of LoserTree+MergeFunction (no JVM in the image)"}

    if rank == 0:
        line = {"metric": metric, "value": value, "unit": "rows/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "int64", "data": "synthetic", "config": config,
                "rows_in_per_gpu": int(n_in), "rows_out_per_gpu": int(n_out), "wall_ms_per_step": 1e3 * wall / args.steps,
                "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu}
        if rewrite is not None:
            line["rewrite"] = rewrite
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
