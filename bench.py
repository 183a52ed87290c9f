#!/usr/bin/env python
"""bench.py -- throughput of the SPRING reorder + encode hot path on B200 (BASELINE.json metric).

A "step" is one pass of the hot path (dictionaries -> chain kernel -> contig encoder) over one
batch of synthetic reads.  At N = 1 the workload is BASELINE.json configs[1]: 10 M single-end
150 bp reads, 50 Mbp uniform genome (30x), 0.5 % substitutions, no qualities.  For N > 1 every
rank holds 10 M reads of one N x 10 M read set (weak scaling); reads are bucketed by a
strand-canonical minimizer, regrouped with one NCCL all-to-all, and each rank then runs the same
single-GPU path on the reads it owns.

  value : whole-job Mreads/s with the packed reads already resident in HBM (CUDA events)
  e2e   : the same through the C ABI with HOST (pinned) buffers: H2D + kernels + D2H per step
  roofline : the chain kernel (k_chains), algorithmic bytes / live CUDA-event duration
  cpu_baseline : the reference's own call_reorder + call_encoder (oracle/_ref) on the host cores,
                 on a bounded sample of the same workload

`--impl reference` times the reference CPU implementation instead (same JSON line).
"""
from __future__ import annotations

import argparse
import json
import os
import shutil
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "Mreads/s, spring -c -r hot path (reorder+encode), 150bp synthetic"
READS_PER_GPU = 10_000_000
READ_LEN = 150
COVERAGE = 30
SUB_RATE = 0.005
SEED = 3
CPU_SAMPLE_READS = 2_000_000


def workload_config(n_gpus: int, reads_per_gpu: int) -> dict:
    total = reads_per_gpu * n_gpus
    return {"workload": f"{total // 1_000_000}M synthetic single-end {READ_LEN}bp reads, -r --no-quality"
                        + (f", sharded over {n_gpus} GPUs by minimizer bucket + all-to-all" if n_gpus > 1 else ", 1xB200"),
            "reads": total, "read_len": READ_LEN, "genome_bp": total * READ_LEN // COVERAGE, "sub_rate": SUB_RATE,
            "seed": SEED, "cache": "inputs (40 B/read packed) larger than L2; no flush needed"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            self.proc.terminate()
            self.t.join(timeout=2)

    def summary(self) -> dict:
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak_gbs() -> tuple[float, str]:
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def chain_kernel_bytes(stats: dict, n_reads: int, words: int) -> float:
    """Algorithmic bytes of one k_chains launch: per sequential-equivalent
    dictionary probe one 32 B sector, per Hamming compare one candidate row + its id + length,
    per claimed read its row once (staging) + 17 B of records + 4 B winner + claim bit."""
    row = 8 * words
    return (32.0 * stats["probes_seq"] + (row + 4 + 2) * stats["compares"] + n_reads * (row + 2 + 17 + 4 + 0.125))


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the reference's own code on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_sample_input(n_reads: int):
    from spring_b200 import synth
    rs = synth.generate(n_reads, READ_LEN, genome_len=n_reads * READ_LEN // COVERAGE, seed=SEED, sub_rate=SUB_RATE)
    return synth.to_hotpath_input(rs)


def time_reference(hp, threads: int) -> tuple[float, str]:
    """seconds for call_reorder + call_encoder on hp; (secs, kind)."""
    from oracle import pyoracle as po
    from spring_b200 import dnaio
    if po.have_reference():
        base = "/dev/shm" if os.path.isdir("/dev/shm") else None
        d = tempfile.mkdtemp(prefix="spring_ref_", dir=base)
        try:
            dnaio.write_hotpath_inputs(d, hp.packed, hp.lengths, max_readlen=hp.max_readlen, n_seqs=hp.n_seqs,
                                       order_n=hp.order_n, num_reads=hp.num_reads)
            tr, te, _ = po.run_reference_hotpath(d, threads, unbsc=False)
            return tr + te, "reference"
        except (RuntimeError, OSError, IndexError) as e:
            # the prebuilt reference binary (built with -march=native in the dev container) does not run on
            # this host: time the single-thread oracle port rather than lose the baseline
            print(f"bench.py: oracle/_ref/spring_ref failed here ({str(e)[:200]!r}); timing the oracle port instead", file=sys.stderr)
        finally:
            shutil.rmtree(d, ignore_errors=True)
    t0 = time.perf_counter()
    po.reorder_encode(hp.packed, hp.lengths, hp.max_readlen, hp.n_records, hp.order_n, hp.num_reads, 1)
    return time.perf_counter() - t0, "port"


def run_reference_arm(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    hp = cpu_sample_input(args.cpu_sample)
    times = []
    kind = "reference"
    for i in range(args.warmup + args.steps):
        secs, kind = time_reference(hp, threads)
        if i >= args.warmup:
            times.append(secs)
    cores = threads if kind == "reference" else 1
    ms = 1e3 * sum(times) / len(times)
    val = hp.num_reads / (ms * 1e-3) / 1e6
    sample = (f"{args.cpu_sample} reads of the same generator (150bp, 30x, 0.5% subs); reference call_reorder+call_encoder "
              f"-t {cores}, temp files on /dev/shm" if kind == "reference" else f"{args.cpu_sample} reads, oracle port, 1 thread")
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "Mreads/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u64", "data": "synthetic", "config": workload_config(args.gpus, args.reads),
            "cpu_baseline": {"value": val, "unit": "Mreads/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": val, "unit": "Mreads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--reads", type=int, default=READS_PER_GPU, help="reads per GPU")
    ap.add_argument("--chains", type=int, default=0, help="0 = as many chains as co-reside")
    ap.add_argument("--cpu-sample", type=int, default=CPU_SAMPLE_READS)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    from spring_b200 import capi, dnaio, multigpu, synth

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the hot path has no CPU fallback")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    # ---- synthetic input, resident in HBM -----------------------------------------------------
    n_local = args.reads
    W = dnaio.words_per_read(READ_LEN)
    total = n_local * world
    # one genome for the whole job; rank r generates the r-th block of reads (same seed => same genome)
    rs = synth.generate(n_local, READ_LEN, genome_len=total * READ_LEN // COVERAGE, seed=SEED, sub_rate=SUB_RATE,
                        device=dev, read_seed=SEED * 1000 + rank)
    d_reads = synth.pack_reads(rs.codes, rs.lengths, READ_LEN).contiguous()
    d_lens = rs.lengths.to(torch.int16).contiguous()
    del rs
    torch.cuda.synchronize()
    stream = torch.cuda.current_stream()
    ctx = capi.Context(local, stream.cuda_stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: inputs already in HBM ---------------------------------------------------------------
    xe0, xe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    exchange_ms = []

    def device_step():
        if world > 1:
            xe0.record(stream)
            r, l = multigpu.exchange_by_bucket(d_reads, d_lens, READ_LEN, world)
            xe1.record(stream)
            xe1.synchronize()
            exchange_ms.append(xe0.elapsed_time(xe1))
        else:
            r, l = d_reads, d_lens
        inp = ctx.make_input(r.data_ptr(), l.data_ptr(), r.shape[0], READ_LEN)
        ctx.reorder_encode_raw(inp, args.chains, device=True)
        return r.shape[0]

    for _ in range(args.warmup):
        device_step()
    barrier()
    launches, stats_acc, chain_ms = 0, None, []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
        e0.record(stream)
        for _ in range(args.steps):
            n_owned = device_step()
            st = ctx.stats()
            launches += st["gpu_launches"]
            chain_ms.append(st["ms_chain_kernel"])
            stats_acc = st
        e1.record(stream)
        barrier()
    ms_dev = e0.elapsed_time(e1) / args.steps
    if world > 1:
        t = torch.tensor([ms_dev], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_dev = float(t.item())
    value = total / (ms_dev * 1e-3) / 1e6

    # ---- e2e: host (pinned) buffers through the C ABI ---------------------------------------------------
    h_reads = torch.empty(d_reads.shape, dtype=d_reads.dtype, pin_memory=True).copy_(d_reads)
    h_lens = torch.empty(d_lens.shape, dtype=d_lens.dtype, pin_memory=True).copy_(d_lens)
    torch.cuda.synchronize()

    def host_step():
        if world > 1:  # H2D, then the same exchange + device path, then D2H of the streams
            r = d_reads.copy_(h_reads, non_blocking=True)
            l = d_lens.copy_(h_lens, non_blocking=True)
            r, l = multigpu.exchange_by_bucket(r, l, READ_LEN, world)
            inp = ctx.make_input(r.data_ptr(), l.data_ptr(), r.shape[0], READ_LEN)
            ctx.reorder_encode_raw(inp, args.chains, device=True)
            s = ctx.fetch_streams_raw()
        else:
            inp = ctx.make_input(h_reads.data_ptr(), h_lens.data_ptr(), n_local, READ_LEN)
            s = ctx.reorder_encode_raw(inp, args.chains, device=False)
        d2h = ((s.seq_len + 3) // 4 + s.num_aligned * 9 + s.noise_bytes + s.num_noise * 2 + s.num_reads * 6 + s.unaligned_bytes)
        return d2h

    for _ in range(max(1, args.warmup - 1)):
        host_step()
    barrier()
    t0 = time.perf_counter()
    d2h_bytes = 0
    for _ in range(args.steps):
        d2h_bytes = host_step()
    barrier()
    ms_e2e = 1e3 * (time.perf_counter() - t0) / args.steps
    if world > 1:
        t = torch.tensor([ms_e2e], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_e2e = float(t.item())
    e2e_val = total / (ms_e2e * 1e-3) / 1e6
    h2d_bytes = int(h_reads.numel() * 8 + h_lens.numel() * 2)

    # ---- the stage after the encoder: pe_encode + reorder_compress_streams re-blocking, streams
    # still resident in HBM; reported next to the headline, not part of it ------------------------------------
    after = None
    if world == 1:
        cpd = dnaio.CompressionParams(paired_end=False, preserve_order=False, num_reads=n_local, max_readlen=READ_LEN)
        cpc = capi.CP.from_buffer_copy(cpd.pack())
        device_step()
        rb_dev, rb_wall = [], []
        for i in range(args.warmup + args.steps):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            blk = ctx.reblock_streams_raw(cpc)
            t1 = time.perf_counter()
            if i >= args.warmup:
                rb_wall.append(1e3 * (t1 - t0)); rb_dev.append(ctx.stats()["ms_reblock"])
        after = {"stage": "reorder_compress_streams re-blocking (src/reorder_compress_streams.cpp:83-361), streams resident in HBM",
                 "gpu_ms": sum(rb_dev) / len(rb_dev), "ms_with_d2h_of_blocks": sum(rb_wall) / len(rb_wall),
                 "blocks": int(blk.num_blocks), "d2h_bytes": int(sum(blk.size[i] for i in range(capi.NUM_BLOCK_STREAMS)))}

    # ---- roofline of the dominant kernel (k_chains) ---------------------------------------------------
    peak, peak_src = measured_peak_gbs()
    ck_ms = sum(chain_ms) / len(chain_ms)
    ck_bytes = chain_kernel_bytes(stats_acc, n_owned, W)
    achieved = ck_bytes / (ck_ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "r01_chain_kernel_traffic.json")
    if os.path.exists(tp):
        tj = json.load(open(tp))
        if tj.get("reads") == n_local and world == 1:
            traffic = tj.get("dram_bytes_per_launch")
    roofline = {"kernel": "k_chains", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src, "kernel_ms": ck_ms,
                "kernel_share_of_step": ck_ms / ms_dev, "algorithmic_bytes_per_launch": ck_bytes,
                "bytes_per_read": ck_bytes / max(n_owned, 1)}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    line = {"metric": METRIC, "value": value, "unit": "Mreads/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_dev, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64",
            "data": "synthetic", "config": workload_config(world, n_local),
            "e2e": {"value": e2e_val, "unit": "Mreads/s", "ms_per_step": ms_e2e, "h2d_bytes_per_step": h2d_bytes,
                    "d2h_bytes_per_step": int(d2h_bytes)},
            "gpu_launches": int(launches), "roofline": roofline, "clocks": clocks.summary(),
            "stages_ms": {k: stats_acc[k] for k in stats_acc if k.startswith("ms_")},
            "chains": stats_acc["num_chains"], "rounds": stats_acc["rounds"], "unmatched": stats_acc["unmatched"],
            "mb_per_s_fastq": value * (2 * READ_LEN + 12)}
    if exchange_ms:  # rank 0's bucket kernel + owner sort + gathers + all-to-alls, per step (inside ms_per_step)
        line["exchange_ms"] = sum(exchange_ms[-args.steps:]) / args.steps
    if after is not None:
        line["after_encoder"] = after
    if world == 1 and not args.no_cpu_baseline:
        if after is not None:  # the same stage on one host core (oracle port of the reference's loops), same streams
            from oracle import pyoracle as po
            st = ctx.fetch_streams()
            t0 = time.perf_counter()
            po.reblock(st, False, False, 256000)
            after["cpu_port_ms"] = 1e3 * (time.perf_counter() - t0)
        hp = cpu_sample_input(args.cpu_sample)
        threads = os.cpu_count() or 1
        secs, kind = time_reference(hp, threads)
        cores = threads if kind == "reference" else 1
        line["cpu_baseline"] = {"value": hp.num_reads / secs / 1e6, "unit": "Mreads/s", "cores": cores, "kind": kind,
                                "sample": f"{args.cpu_sample} reads of the same generator (150bp, 30x coverage, 0.5% subs), "
                                          f"{secs:.1f} s of call_reorder+call_encoder at -t {cores}"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
