#!/usr/bin/env python
"""bench.py — BASELINE.json's headline metric on its configs[1] workload.

  metric   : LZ4 frame compress/decompress GB/s (uncompressed)
  workload : synthetic log text, independent 64 KiB blocks, level 1, block checksums on,
             no content checksum, device-resident; 8 GiB per GPU by default (weak scaling:
             blocks are independent, ranks share nothing and there is no collective on the data path)
  step     : one compress pass + one decompress pass over the whole workload
  value    : uncompressed bytes through both passes / device time  (2*U / (t_c + t_d)), summed over ranks
  e2e      : the same step through the host-buffer C ABI (plz4cu_*_batch_host) with pinned host
             memory, H2D and D2H inside the timed region
  roofline : the dominant kernel (lz4_compress_cta_kernel): algorithmic bytes U + C' per launch over
             its CUDA-event time, against MEASURED_PEAKS.json hbm_gbs
  cpu_baseline / --impl reference : the reference's own liblz4 (oracle/_ref, compiled from the
             reference's vendored C) driven by a pthread fan-out over blocks on all host cores
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
sys.path.insert(0, ROOT)

BSZ = 64 << 10
SEED = 0x504C5A34
METRIC = "LZ4 frame compress/decompress GB/s (uncompressed)"
UNIT = "GB/s"


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


# ------------------------------------------------------------------ clocks

class ClockSampler:
    """nvidia-smi sampled in the background DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.rows = []
        self.proc = None
        self.n0 = 0

    def mark(self):
        """The timed region starts here: only samples taken from now on are reported."""
        self.n0 = len(self.rows)

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.idx)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        rows = self.rows[self.n0:]
        nearest = not rows and bool(self.rows)
        if nearest:                      # region shorter than one sampling period: the rows right before it, under the same load
            rows = self.rows[-2:]
        sm = sorted(float(r[1]) for r in rows if len(r) > 8 and r[1].replace(".", "").isdigit())
        mx = [float(r[2]) for r in rows if len(r) > 8 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in rows:
            if len(r) < 9:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        out = {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
               "reasons": sorted(reasons), "samples": len(sm)}
        if nearest:
            out["note"] = "timed region shorter than the sampling period: samples taken during the warm-up just before it"
        return out


# ------------------------------------------------------------------ CPU reference arm

def cpu_reference(sample_bytes: int, steps: int, warmup: int, threads: int | None = None) -> dict:
    """The reference's CPU path (oracle/_ref liblz4 when built, else the pinned port) on all host cores."""
    import numpy as np
    from oracle import oracle as O
    O.build()
    port = O.Port()
    drv = C.CDLL(os.path.join(ROOT, "oracle", "cpu_driver.so"))
    drv.drv_compress_blocks.restype = C.c_int
    drv.drv_compress_blocks.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
    drv.drv_decompress_blocks.restype = C.c_int
    drv.drv_decompress_blocks.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
    drv.drv_gen_logtext.restype = C.c_int
    drv.drv_gen_logtext.argtypes = [C.c_uint32, C.c_uint64, C.c_void_p, C.c_uint64, C.c_int]
    if O.Ref.available():
        ref = O.Ref()
        kind = "reference"
        cfn = C.cast(ref.lib.LZ4_compress_fast, C.c_void_p)
        dfn = C.cast(ref.lib.LZ4_decompress_safe, C.c_void_p)
    else:
        kind = "port"
        # adapters with liblz4's argument order live in the port library
        cfn = C.cast(port.lib.orc_lz4_compress_fast, C.c_void_p)
        dfn = C.cast(port.lib.orc_lz4_decompress_safe, C.c_void_p)
    xfn = C.cast(port.lib.orc_xxh32, C.c_void_p)
    threads = threads or os.cpu_count() or 1
    # three buffers of the sample's size live on the host: shrink the sample to what the box has to spare, and say so
    try:
        with open("/proc/meminfo") as f:
            avail = next(int(l.split()[1]) for l in f if l.startswith("MemAvailable")) << 10
        sample_bytes = min(sample_bytes, max(64 << 20, avail // 5))
    except Exception:
        pass
    nblk = max(1, sample_bytes // BSZ)
    total = nblk * BSZ
    src = np.empty(total, dtype=np.uint8)
    drv.drv_gen_logtext(SEED, 0, C.c_void_p(src.ctypes.data), total, threads)   # the workload's generator, built into the oracle driver
    recs = np.empty(nblk * (BSZ + 8), dtype=np.uint8)
    rec_len = np.zeros(nblk, dtype=np.uint32)
    out = np.empty(total, dtype=np.uint8)
    out_len = np.zeros(nblk, dtype=np.uint32)
    rec_off = (np.arange(nblk, dtype=np.uint64) * (BSZ + 8))
    vp = lambda a: C.c_void_p(a.ctypes.data)
    tc = td = 0.0
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        drv.drv_compress_blocks(cfn, xfn, vp(src), total, BSZ, 1, vp(recs), vp(rec_len), threads)
        t1 = time.perf_counter()
        errs = drv.drv_decompress_blocks(dfn, xfn, vp(recs), vp(rec_off), vp(rec_len), nblk, BSZ, 1, vp(out), vp(out_len), threads)
        t2 = time.perf_counter()
        if errs:
            raise RuntimeError("cpu baseline: decode errors")
        if it >= warmup:
            tc += t1 - t0
            td += t2 - t1
    assert bytes(out[:4096]) == bytes(src[:4096])
    csize = int(rec_len.sum())
    return {
        "kind": kind, "cores": threads, "sample": f"{total >> 20} MiB logtext, 64 KiB blocks, block checksums, {steps} passes",
        "value": 2 * total * steps / (tc + td) / 1e9, "unit": UNIT,
        "compress_gbs": total * steps / tc / 1e9, "decompress_gbs": total * steps / td / 1e9,
        "ratio": csize / total, "ms_per_step": (tc + td) / steps * 1e3, "bytes": total,
    }


def run_reference(args) -> None:
    rank = env_int("RANK", 0)
    if rank != 0:
        return
    r = cpu_reference(args.cpu_sample_mib << 20, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": round(r["value"], 4), "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(r["ms_per_step"], 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": workload_config(args, r["bytes"]),
        "compress_gbs": round(r["compress_gbs"], 4), "decompress_gbs": round(r["decompress_gbs"], 4),
        "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": round(r["value"], 4), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, nbytes_per_gpu: int) -> dict:
    return {
        "workload": "BASELINE configs[1]: synthetic log text, independent 64 KiB blocks, level 1, "
                    "block checksums, no content checksum, device-resident",
        "bytes_per_gpu": int(nbytes_per_gpu), "block_size": BSZ, "blocks_per_gpu": int(nbytes_per_gpu // BSZ),
        "level": 1, "block_checksum": True, "content_checksum": False, "seed": hex(SEED),
        "step": "compress pass + decompress pass over all blocks",
        "l2": "inputs (GiBs) far exceed the 126 MB L2; no flush needed",
        "parallelism": f"{args.gpus} x independent block shards, no collective",
    }


# ------------------------------------------------------------------ GPU arm

def run_gpu(args) -> None:
    import numpy as np
    import torch
    import torch.distributed as dist
    from plz4_b200 import _lib
    from plz4_b200._lib import check

    world = env_int("WORLD_SIZE", 1)
    rank = env_int("RANK", 0)
    local = env_int("LOCAL_RANK", 0)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the engine has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    # The engine's host threads wait for the GPU on blocking events by default: a waiting call costs no core, which is what
    # lets eight ranks share one host.  A lone process has cores to spare and spins instead (saves the wake-up latency of
    # every chunk: ~5 % of the end-to-end figure).  Stated in the line as e2e.host_wait.
    if world == 1:
        os.environ.setdefault("PLZ4CU_SPIN", "1")
    L = _lib.lib()
    check(L.plz4cu_init(local), "plz4cu_init")

    nbytes = int(args.gib * (1 << 30)) // BSZ * BSZ
    nblk = nbytes // BSZ
    stride = BSZ + 16
    stream = torch.cuda.current_stream()
    sh = C.c_void_p(stream.cuda_stream)
    u8 = lambda n: torch.empty(n, dtype=torch.uint8, device=dev)
    src, recs, out = u8(nbytes), u8(nblk * stride), u8(nbytes)
    src_off = torch.arange(nblk, dtype=torch.int64, device=dev) * BSZ
    src_len = torch.full((nblk,), BSZ, dtype=torch.int32, device=dev)
    rec_off = torch.arange(nblk, dtype=torch.int64, device=dev) * stride
    rec_len = torch.zeros(nblk, dtype=torch.int32, device=dev)
    out_len = torch.zeros(nblk, dtype=torch.int32, device=dev)
    p = lambda t: C.c_void_p(t.data_ptr())
    # every rank gets its own slice of the stream: segments [rank*nblk, (rank+1)*nblk)
    check(L.plz4cu_gen_logtext_device(sh, SEED, rank * nblk, p(src), nbytes), "gen_logtext")
    torch.cuda.synchronize()

    def compress():
        check(L.plz4cu_compress_batch_device(sh, p(src), p(src_off), p(src_len), nblk, BSZ, 1, 0, None,
                                             p(recs), stride, p(rec_len)), "compress_batch_device")

    def decompress():
        check(L.plz4cu_decompress_batch_device(sh, p(recs), p(rec_off), None, nblk, BSZ, 1, 0, None,
                                               p(out), BSZ, p(out_len)), "decompress_batch_device")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # nvidia-smi needs up to a second to deliver its first row on a multi-GPU box: start it before the warm-up and
    # count only the rows that arrive inside the timed region (sampler.mark() below)
    sampler = ClockSampler(local)
    sampler.start()
    for _ in range(args.warmup):
        compress()
        decompress()
    barrier()
    # correctness gate inside the bench: the round trip must be exact and every block decoded
    assert bool((out_len == BSZ).all()), "decode failed"
    assert torch.equal(out, src), "round trip mismatch"
    csize = int(rec_len.to(torch.int64).sum())      # C' = framed compressed bytes (size word + payload + xxh32)

    launches0 = L.plz4cu_launch_count()
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    barrier()
    sampler.mark()
    t_wall0 = time.perf_counter()
    for k in range(args.steps):
        ev[k][0].record(stream)
        compress()
        ev[k][1].record(stream)
        decompress()
        ev[k][2].record(stream)
    barrier()
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop()
    launches = L.plz4cu_launch_count() - launches0
    tc = sum(e[0].elapsed_time(e[1]) for e in ev) / 1e3
    td = sum(e[1].elapsed_time(e[2]) for e in ev) / 1e3
    tt = ev[0][0].elapsed_time(ev[-1][2]) / 1e3

    # ---- e2e through the host-buffer C ABI (pinned host memory; H2D + D2H inside the timed region)
    e2e = None
    if not args.no_e2e:
        # the same bytes per GPU at every N (three pinned buffers per rank), so that the 1 -> 8 curve is scaling and nothing else
        e_bytes = min(nbytes, int(args.e2e_gib * (1 << 30)) // BSZ * BSZ)
        e_blk = e_bytes // BSZ
        h_src = torch.empty(e_bytes, dtype=torch.uint8).pin_memory()
        h_src.copy_(src[:e_bytes])
        h_packed = torch.empty(e_blk * (BSZ + 8), dtype=torch.uint8).pin_memory()
        h_out = torch.empty(e_bytes, dtype=torch.uint8).pin_memory()
        h_len = np.full(e_blk, BSZ, dtype=np.uint32)
        h_res = np.zeros(e_blk, dtype=np.int32)
        vp = lambda a: C.c_void_p(a.ctypes.data)
        # The step is cut into parts: part k is decompressed (D2H-heavy) while part k+1 is still being compressed
        # (H2D-heavy), two caller threads making the same two public calls — PCIe is full duplex and the engine
        # gives every concurrent call its own pipeline.  --e2e-parts 1 runs the two passes strictly one after the other.
        n_parts = max(1, min(args.e2e_parts, e_blk))
        cuts = [e_blk * k // n_parts for k in range(n_parts + 1)]
        parts = []
        for k in range(n_parts):
            b0, b1 = cuts[k], cuts[k + 1]
            parts.append({"b0": b0, "nb": b1 - b0, "off": np.arange(b1 - b0, dtype=np.uint64) * BSZ,
                          "poff": np.zeros(b1 - b0 + 1, dtype=np.uint64), "ready": threading.Event(),
                          "src": C.c_void_p(h_src.data_ptr() + b0 * BSZ), "out": C.c_void_p(h_out.data_ptr() + b0 * BSZ),
                          "packed": C.c_void_p(h_packed.data_ptr() + b0 * (BSZ + 8)), "cap": (b1 - b0) * (BSZ + 8)})
        errors = []

        def compress_parts():
            try:
                check(L.plz4cu_init(local), "plz4cu_init")          # a new thread starts on device 0
                for q in parts:
                    check(L.plz4cu_compress_batch_host(q["src"], vp(q["off"]), vp(h_len[q["b0"]:]), q["nb"], BSZ, 1, 0, None,
                                                       q["packed"], q["cap"], vp(q["poff"])), "compress_batch_host")
                    q["ready"].set()
            except BaseException as ex:          # noqa: BLE001 - re-raised on the main thread
                errors.append(ex)
                for q in parts:
                    q["ready"].set()

        def decompress_parts():
            try:
                check(L.plz4cu_init(local), "plz4cu_init")
                for q in parts:
                    q["ready"].wait()
                    if errors:
                        return
                    check(L.plz4cu_decompress_batch_host(q["packed"], int(q["poff"][q["nb"]]), vp(q["poff"]), None, q["nb"], BSZ, 1, 0,
                                                         None, q["out"], BSZ, vp(h_res[q["b0"]:])), "decompress_batch_host")
            except BaseException as ex:          # noqa: BLE001
                errors.append(ex)

        def e2e_step():
            for q in parts:
                q["ready"].clear()
            ts = [threading.Thread(target=compress_parts), threading.Thread(target=decompress_parts)]
            for t in ts:
                t.start()
            for t in ts:
                t.join()
            if errors:
                raise errors[0]
        for _ in range(max(1, min(args.warmup, 3))):
            e2e_step()
        assert (h_res == BSZ).all() and torch.equal(h_out, h_src), "e2e round trip mismatch"
        barrier()
        t0 = time.perf_counter()
        e_steps = max(1, args.steps)
        for _ in range(e_steps):
            e2e_step()
        barrier()
        te = time.perf_counter() - t0
        c_e = sum(int(q["poff"][q["nb"]]) for q in parts)
        e2e = {"t": te, "steps": e_steps, "bytes": e_bytes, "parts": n_parts,
               "h2d": e_bytes + c_e + e_blk * 24, "d2h": c_e + e_bytes + e_blk * 12 + 8}
        # copy ceiling: the step's H2D and D2H bytes over the same pinned buffers, both directions at once, no kernels
        d_a, d_b = src[:e_bytes], out[:e_bytes]
        s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()
        up_n, dn_n = e2e["h2d"], e2e["d2h"]

        def copy_step():
            with torch.cuda.stream(s_up):
                d_a.copy_(h_src, non_blocking=True)                                       # U up (compress input)
                recs[:c_e].copy_(h_packed[:c_e], non_blocking=True)                      # C' up (decompress input)
            with torch.cuda.stream(s_dn):
                h_packed[:c_e].copy_(recs[:c_e], non_blocking=True)                      # C' down (compress output)
                h_out.copy_(d_b, non_blocking=True)                                       # U down (decompress output)
            s_up.synchronize(); s_dn.synchronize()
        copy_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e_steps):
            copy_step()
        barrier()
        e2e["t_copy"] = time.perf_counter() - t0
        del h_src, h_packed, h_out

    # ---- max over ranks, sum of work
    def allmax(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    tt_max, tc_max, td_max = allmax(tt), allmax(tc), allmax(td)
    te_max = allmax(e2e["t"]) if e2e else None
    tcopy_max = allmax(e2e["t_copy"]) if e2e else None
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks, peak_src = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = float(json.load(f)["hbm_gbs"])
            peak_src = "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)"
    except Exception:
        pass
    # DRAM traffic per launch comes from an ncu --set full capture of this same workload made in this round
    # (tools/ncu_traffic.py writes the file and records the commit it measured); never measured in-run
    traffic, traffic_src = {}, None
    try:
        with open(os.path.join(ROOT, "profiles", "r02_traffic.json")) as f:
            t = json.load(f)
        if int(t.get("bytes_per_gpu", 0)) == nbytes:
            traffic = {k: v["traffic"] for k, v in t.items() if isinstance(v, dict) and "traffic" in v}
            traffic_src = "profiles/r02_traffic.json (ncu --set full, commit %s)" % t.get("commit", "?")
    except Exception:
        pass
    K = args.steps
    value = world * 2 * nbytes * K / tt_max / 1e9
    algo = nbytes + csize                       # per launch: U read + C' written (compress); C' read + U written (decompress)
    c_ach = algo / (tc / K) / 1e9
    d_ach = algo / (td / K) / 1e9
    line = {
        "metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": world, "steps": K, "warmup": args.warmup,
        "ms_per_step": round(tt_max / K * 1e3, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic", "config": workload_config(args, nbytes),
        "compress_gbs": round(world * nbytes * K / tc_max / 1e9, 3),
        "decompress_gbs": round(world * nbytes * K / td_max / 1e9, 3),
        "compressed_ratio": round(csize / nbytes, 5),
        "roofline": {"bound": "hbm", "kernel": "lz4_compress_cta_kernel", "achieved": round(c_ach, 2), "peak": peaks, "unit": "GB/s",
                     "frac": round(c_ach / peaks, 5), "traffic": traffic.get("lz4_compress_cta_kernel"), "traffic_source": traffic_src, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": algo, "launch_ms": round(tc / K * 1e3, 3)},
        "roofline_decompress": {"bound": "hbm", "kernel": "lz4_decompress_kernel", "achieved": round(d_ach, 2), "peak": peaks,
                                "unit": "GB/s", "frac": round(d_ach / peaks, 5), "traffic": traffic.get("lz4_decompress_kernel"),
                                "algorithmic_bytes_per_launch": algo, "launch_ms": round(td / K * 1e3, 3)},
        "gpu_launches": int(launches), "clocks": clocks, "wall_s": round(t_wall, 3),
    }
    if e2e:
        line["e2e"] = {"value": round(world * 2 * e2e["bytes"] * e2e["steps"] / te_max / 1e9, 3), "unit": UNIT,
                       "h2d_bytes_per_step": int(e2e["h2d"]), "d2h_bytes_per_step": int(e2e["d2h"]),
                       "bytes_per_gpu": int(e2e["bytes"]), "steps": e2e["steps"],
                       "copy_ceiling": round(world * 2 * e2e["bytes"] * e2e["steps"] / tcopy_max / 1e9, 3),
                       "copy_ceiling_note": "same H2D + D2H byte counts per step over the same pinned buffers, both directions at once, no kernels; same unit as value",
                       "host_wait": "spin" if os.environ.get("PLZ4CU_SPIN", "0") not in ("", "0") else "blocking",
                       "api": "plz4cu_compress_batch_host + plz4cu_decompress_batch_host, pinned host buffers, %d parts: part k decompresses while part k+1 compresses" % e2e["parts"]}
    if not args.no_cpu:                      # rank 0; the full sample at N=1, a quarter of it at N>1 (the other ranks wait)
        try:
            r = cpu_reference((args.cpu_sample_mib << 20) // (1 if world == 1 else 4), 3, 1)
            line["cpu_baseline"] = {k: (round(r[k], 4) if isinstance(r[k], float) else r[k])
                                    for k in ("value", "unit", "cores", "kind", "sample", "compress_gbs", "decompress_gbs", "ratio")}
        except Exception as e:        # the baseline is a reported number, never a reason to lose the GPU line
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "unavailable", "sample": str(e)}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def _keep_stdout_for_the_line():
    """Native libraries write to file descriptor 1 too (NCCL prints its version banner there): point fd 1 at stderr for
    the duration of the run and give `print` the real stdout, so that stdout carries the one JSON line and nothing else."""
    import sys
    sys.stdout.flush()
    real = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real, "w", buffering=1)


def main():
    _keep_stdout_for_the_line()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="plz4_b200", choices=["plz4_b200", "reference"])
    ap.add_argument("--gib", type=float, default=8.0, help="uncompressed GiB per GPU (configs[1] = 8)")
    ap.add_argument("--e2e-parts", type=int, default=8, help="parts the e2e step is cut into (decompress of part k overlaps compress of part k+1); 1 = strictly sequential passes")
    ap.add_argument("--e2e-gib", type=float, default=4.0, help="GiB per GPU pushed through the host-buffer API per step (the same at every N)")
    ap.add_argument("--cpu-sample-mib", type=int, default=8192, help="sample for the CPU baseline / reference arm (8192 = the whole configs[1] workload)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3 if args.impl != "reference" else max(args.warmup, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
