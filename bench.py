#!/usr/bin/env python
"""bench.py -- throughput of the SPRING reorder + encode hot path on B200 (BASELINE.json metric).

A "step" is one pass of the hot path (dictionaries -> chain kernel -> contig encoder) over one batch of
synthetic reads.  The workload is one of BASELINE.json's configs:

  --config 3 (default, the metric's own): 100 M paired-end 150 bp reads (= 50 M pairs), 500 Mbp uniform
             genome (30x), insert ~U[200,500], Illumina-like error model (substitution rate 0.1 % -> 1 %
             along the read), 0.2 % of the reads carry 1-3 N, `-r`
  --config 2: 10 M single-end 150 bp reads, 50 Mbp genome, 0.5 % uniform substitutions, `-r --no-quality`
  --config 5: 50 M single-end reads of 35-250 bp, 250 Mbp genome, 0.5 % substitutions, `-r`

For N > 1 every rank holds `reads` reads of one N x `reads` read set over an N x larger genome (weak scaling,
N = 8 at config 3 is 800 M reads: config 4's scale); reads are bucketed by a strand-canonical minimizer,
regrouped with one all-to-all together with their global ids, and each rank then runs the single-GPU path on
the reads it owns.

  value : whole-job Mreads/s with the packed reads already resident in HBM (CUDA events)
  e2e   : the same through the C ABI with HOST (pinned) buffers: H2D + kernels + D2H per step
  verify: after the timed region, the streams of the last step are re-blocked, block-decoded and compared with
          the input on the GPU (spring_b200_verify_roundtrip): every read must decode to its original
  e2e_files : the same through the file-level plugin call (temp_dir on /dev/shm in, stream files out)
  roofline : the chain kernel (k_chains), algorithmic bytes / live CUDA-event duration
  cpu_baseline : the reference's own call_reorder + call_encoder (oracle/_ref) on the host cores, on a
                 bounded sample of the same generator

`--impl reference` times the reference CPU implementation on the same reads instead (same JSON line); at
100 M reads one step takes minutes, so the number of steps is capped and the line says what really ran.
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

METRIC = "Mreads/s, spring -c -r hot path (reorder+encode), 150bp PE synthetic"
COVERAGE = 30

CONFIGS = {
    2: dict(label="synthetic single-end 150bp reads, -r --no-quality", reads=10_000_000, read_len=150, paired=False,
            error_model="uniform", sub_rate=0.005, n_frac=0.0, var_len=None, seed=3, mean_len=150),
    3: dict(label="synthetic paired-end 150bp reads (pairs = reads / 2, Illumina error model 0.1%->1%, 0.2% reads with N), -r",
            reads=100_000_000, read_len=150, paired=True, error_model="illumina", sub_rate=0.0, n_frac=0.002, var_len=None,
            seed=4, mean_len=150),
    5: dict(label="synthetic variable-length 35-250bp single-end reads, -r", reads=50_000_000, read_len=250, paired=False,
            error_model="uniform", sub_rate=0.005, n_frac=0.0, var_len=(35, 250), seed=6, mean_len=142.5),
}


def genome_len(cfg: dict, total_reads: int) -> int:
    return max(int(total_reads * cfg["mean_len"] / COVERAGE), 4 * cfg["read_len"] + 600)


def workload_config(cfg_id: int, cfg: dict, n_gpus: int, reads_per_gpu: int) -> dict:
    total = reads_per_gpu * n_gpus
    return {"workload": f"config {cfg_id}: {total / 1e6:g}M {cfg['label']}"
                        + (f", sharded over {n_gpus} GPUs by minimizer bucket + all-to-all" if n_gpus > 1 else ", 1xB200"),
            "config_id": cfg_id, "reads": total, "read_len": cfg["read_len"], "paired": cfg["paired"],
            "genome_bp": genome_len(cfg, total), "error_model": cfg["error_model"], "sub_rate": cfg["sub_rate"],
            "n_frac": cfg["n_frac"], "var_len": cfg["var_len"], "seed": cfg["seed"],
            "cache": "inputs (40+ B/read packed, >= 400 MB) larger than L2; no flush needed"}


def generate(cfg: dict, n_reads: int, total_reads: int, device, rank: int = 0):
    from spring_b200 import synth
    return synth.generate(n_reads, cfg["read_len"], genome_len=genome_len(cfg, total_reads), seed=cfg["seed"],
                          sub_rate=cfg["sub_rate"], error_model=cfg["error_model"], var_len=cfg["var_len"],
                          paired=cfg["paired"], n_frac=cfg["n_frac"], device=device, read_seed=cfg["seed"] * 1000 + rank)


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
def host_input(cfg: dict, n_reads: int):
    """Reads of the config's generator as host arrays (generated on the GPU when there is one: the
    generator is harness plumbing, only the timed call must be the reference's CPU code)."""
    import torch
    from spring_b200 import synth
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    rs = generate(cfg, n_reads, n_reads, dev)
    di = synth.to_device_input(rs)
    del rs
    packed = di.reads.cpu().numpy().view("uint64")
    lengths = di.lengths.cpu().numpy().view("uint16")
    return packed, lengths, di


def time_reference(packed, lengths, di, threads: int) -> tuple[float, str]:
    """seconds for call_reorder + call_encoder; (secs, kind)."""
    from oracle import pyoracle as po
    from spring_b200 import dnaio
    if po.have_reference():
        base = "/dev/shm" if os.path.isdir("/dev/shm") else None
        d = tempfile.mkdtemp(prefix="spring_ref_", dir=base)
        try:
            dnaio.write_hotpath_inputs(d, packed, lengths, max_readlen=di.max_readlen, n_seqs=di.n_seqs, order_n=di.order_n,
                                       num_reads=di.num_reads, paired_split=di.num_clean[0] if di.paired else None)
            tr, te, _ = po.run_reference_hotpath(d, threads, unbsc=False)
            return tr + te, "reference"
        except (RuntimeError, OSError, IndexError) as e:
            # the prebuilt reference binary (built with -march=native in the dev container) does not run on
            # this host: time the single-thread oracle port rather than lose the baseline
            print(f"bench.py: oracle/_ref/spring_ref failed here ({str(e)[:200]!r}); timing the oracle port instead", file=sys.stderr)
        finally:
            shutil.rmtree(d, ignore_errors=True)
    t0 = time.perf_counter()
    po.reorder_encode(packed, lengths, di.max_readlen, di.n_records, di.order_n, di.num_reads, 1)
    return time.perf_counter() - t0, "port"


def run_reference_arm(args, cfg_id: int, cfg: dict) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import pyoracle as po
    threads = os.cpu_count() or 1
    # the reads of the N = 1 workload, all of them (for N > 1 the arm stays on the N = 1 workload: a CPU run of
    # N x 100 M reads does not fit "a few minutes", and Mreads/s of the CPU path does not grow with N)
    n_reads = args.ref_reads or args.reads
    if not po.have_reference():
        n_reads = min(n_reads, 2_000_000)  # single-thread port: bounded sample
    packed, lengths, di = host_input(cfg, n_reads)
    # The first pass always runs.  If the budget leaves room for more, it counts as a warm-up and further passes
    # are timed; if one pass already eats the budget (100 M reads: minutes), it IS the timed step.
    t_begin = time.perf_counter()
    secs, kind = time_reference(packed, lengths, di, threads)
    per_pass = time.perf_counter() - t_begin          # includes writing the temp files
    passes = max(1, min(args.warmup + args.steps, int(args.ref_budget_s // per_pass)))
    if passes == 1:
        times, steps_run, warm_run = [secs], 1, 0
    else:
        warm_run = max(1, min(args.warmup, passes - args.steps))
        steps_run = passes - warm_run
        for _ in range(warm_run - 1):
            time_reference(packed, lengths, di, threads)
        times = [time_reference(packed, lengths, di, threads)[0] for _ in range(steps_run)]
    cores = threads if kind == "reference" else 1
    ms = 1e3 * sum(times) / len(times)
    val = di.num_reads / (ms * 1e-3) / 1e6
    what = (f"reference call_reorder+call_encoder -t {cores}, temp files on /dev/shm" if kind == "reference"
            else "oracle port, 1 thread")
    sample = (f"all {di.num_reads} reads of config {cfg_id} at N = 1 (same generator, same seed as the GPU arm); {what}; "
              f"{steps_run} timed step(s) after {warm_run} warm-up(s) (requested {args.steps}/{args.warmup}, capped by a "
              f"{args.ref_budget_s:.0f} s budget: one pass takes {ms / 1e3:.0f} s)")
    conf = workload_config(cfg_id, cfg, 1, di.num_reads)
    if args.gpus > 1:
        conf["note"] = f"GPU arm at N = {args.gpus} runs {args.gpus} x {args.reads} reads; this arm times the N = 1 workload"
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "Mreads/s", "n_gpus": args.gpus, "steps": steps_run,
            "warmup": warm_run, "steps_requested": args.steps, "warmup_requested": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u64", "data": "synthetic", "config": conf,
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
    ap.add_argument("--config", type=int, default=3, choices=sorted(CONFIGS))
    ap.add_argument("--reads", type=int, default=0, help="reads per GPU (default: the config's size)")
    ap.add_argument("--chains", type=int, default=0, help="0 = as many chains as co-reside")
    ap.add_argument("--cpu-sample", type=int, default=10_000_000, help="reads of the cpu_baseline sample")
    ap.add_argument("--ref-reads", type=int, default=0, help="--impl reference: reads to time (default: the whole N = 1 workload)")
    ap.add_argument("--ref-budget-s", type=float, default=300.0, help="--impl reference: wall-clock budget for warm-ups + steps")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-verify", action="store_true")
    ap.add_argument("--no-files-leg", action="store_true")
    ap.add_argument("--profile-after-setup", action="store_true",
                    help="cudaProfilerStart() once the synthetic input exists (ncu --profile-from-start off: launch lists without the generator's kernels)")
    args = ap.parse_args()
    cfg_id, cfg = args.config, CONFIGS[args.config]
    if not args.reads:
        args.reads = cfg["reads"]
    if cfg["paired"]:
        args.reads -= args.reads & 1
    if args.impl == "reference":
        run_reference_arm(args, cfg_id, cfg)
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
    L = cfg["read_len"]
    W = dnaio.words_per_read(L)
    total = n_local * world
    # one genome for the whole job; rank r generates the r-th block of reads / pairs (same seed => same genome)
    rs = generate(cfg, n_local, total, dev, rank)
    di = synth.to_device_input(rs)
    del rs
    torch.cuda.empty_cache()
    d_reads, d_lens = di.reads, di.lengths
    n_clean = int(d_reads.shape[0])
    n_records, order_n = di.n_records, di.order_n
    torch.cuda.synchronize()
    stream = torch.cuda.current_stream()
    ctx = capi.Context(local, stream.cuda_stream)
    if args.profile_after_setup:
        torch.cuda.cudart().cudaProfilerStart()
    # global ids of this rank's reads (paired: file-2 mates are numbered total / 2 + pair index); they travel with the
    # clean reads through the exchange, the rank's own N reads keep theirs for the shard finalisation
    d_ids, n_ids = None, None
    if world > 1:
        gid = multigpu.global_ids(n_local, rank, world, cfg["paired"], dev)
        isn = torch.zeros(n_local, dtype=torch.bool, device=dev)
        if len(order_n):
            isn[torch.from_numpy(order_n.astype(np.int64)).to(dev)] = True
        d_ids = gid[~isn].contiguous()
        n_ids = gid[isn].cpu().numpy().astype(np.uint32)
        multigpu.init_comm(ctx, rank, world, dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: inputs already in HBM ---------------------------------------------------------------
    exchange_ms = []
    keep = {}

    def device_step(finalize=True):
        if world > 1:
            # the library's exchange: bucket + histogram kernel, stable scatter, one NCCL group of send / recv
            x = ctx.exchange_reads(d_reads.data_ptr(), d_lens.data_ptr(), d_ids.data_ptr(), n_clean, L)
            n_own = int(x.num_reads)
            # the rank's own N reads stay where they are, numbered after the clean reads it owns
            on = (n_own + np.arange(len(order_n))).astype(np.uint32)
            inp = ctx.make_input(x.reads, x.lengths, n_own, L, n_records, on, n_own + len(on))
            ctx.reorder_encode_raw(inp, args.chains, device=True)
            if finalize:  # absolute positions over all ranks' consensus shards + global ids (the reference's thread-file merge)
                keep["layout"] = ctx.finalize_shard(x.ids, n_own, n_ids)
                exchange_ms.append(ctx.stats()["ms_exchange"])
        else:
            n_own = n_clean
            inp = ctx.make_input(d_reads.data_ptr(), d_lens.data_ptr(), n_clean, L, n_records, order_n, n_local)
            ctx.reorder_encode_raw(inp, args.chains, device=True)
        return n_own

    # the lookups / compares the roofline's algorithmic bytes are made of: ONE pass of the counting chain kernel, outside the
    # timed region (the timed passes run the production instantiation, which leaves the counting out of its hot loops)
    ctx.set_chain_stats(True)
    device_step()
    counted = ctx.stats()
    ctx.set_chain_stats(False)
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
    exchange_avg = sum(exchange_ms[-args.steps:]) / args.steps if exchange_ms else None
    if world > 1:
        t = torch.tensor([ms_dev], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_dev = float(t.item())
    value = total / (ms_dev * 1e-3) / 1e6

    # ---- verify: the streams of the last step, re-blocked, block-decoded and compared with the input in HBM ----
    verify = None
    if not args.no_verify:
        if world == 1:
            cpd = dnaio.CompressionParams(paired_end=cfg["paired"], preserve_order=False, num_reads=n_local, max_readlen=L)
        else:  # a rank's shard on its own, as a single-end job over the reads it owns (pairs are split across ranks)
            cpd = dnaio.CompressionParams(paired_end=False, preserve_order=False, num_reads=n_owned + len(order_n), max_readlen=L)
        if world > 1:
            n_owned = device_step(finalize=False)  # the check compares with the rank's own numbering
            cpd.num_reads = n_owned + len(order_n)
        t0 = time.perf_counter()
        v = ctx.verify_roundtrip(capi.CP.from_buffer_copy(cpd.pack()))
        verify = {"ok": bool(v["ok"]), "reads_checked": int(v["reads_checked"]), "mismatching_reads": int(v["base_mismatch_reads"] + v["length_mismatch_reads"]),
                  "bad_order": int(v["bad_order"]), "blocks": int(v["num_blocks"]), "block_stream_bytes": int(v["block_stream_bytes"]),
                  "seconds": time.perf_counter() - t0,
                  "how": "spring_b200_verify_roundtrip: re-block -> block decode -> exact compare with the input, on the GPU"}
        if world > 1:
            okt = torch.tensor([1 if verify["ok"] else 0], device=dev)
            dist.all_reduce(okt, op=dist.ReduceOp.MIN)
            verify["ok"] = bool(okt.item())
            verify["scope"] = "every rank verifies its own shard; ok = all ranks ok"

    # ---- e2e: host (pinned) buffers through the C ABI ---------------------------------------------------
    h_reads = torch.empty(d_reads.shape, dtype=d_reads.dtype, pin_memory=True).copy_(d_reads)
    h_lens = torch.empty(d_lens.shape, dtype=d_lens.dtype, pin_memory=True).copy_(d_lens)
    torch.cuda.synchronize()

    def host_step():
        if world > 1:  # H2D, then the same exchange + device path + finalisation, then D2H of the streams
            d_reads.copy_(h_reads, non_blocking=True)
            d_lens.copy_(h_lens, non_blocking=True)
            device_step()
            s = ctx.fetch_streams_raw()
        else:
            inp = ctx.make_input(h_reads.data_ptr(), h_lens.data_ptr(), n_clean, L, n_records, order_n, n_local)
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
    h2d_bytes = int(h_reads.numel() * 8 + h_lens.numel() * 2 + len(n_records) + 4 * len(order_n))
    del h_reads, h_lens

    # ---- the stage after the encoder: pe_encode + reorder_compress_streams re-blocking, streams
    # still resident in HBM; reported next to the headline, not part of it ------------------------------------
    after = None
    if world == 1:
        cpd = dnaio.CompressionParams(paired_end=cfg["paired"], preserve_order=False, num_reads=n_local, max_readlen=L)
        cpc = capi.CP.from_buffer_copy(cpd.pack())
        device_step()
        rb_dev, rb_wall = [], []
        for i in range(2 + 3):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            blk = ctx.reblock_streams_raw(cpc)
            t1 = time.perf_counter()
            if i >= 2:
                rb_wall.append(1e3 * (t1 - t0)); rb_dev.append(ctx.stats()["ms_reblock"])
        after = {"stage": ("pe_encode (src/pe_encode.cpp:24-84) + " if cfg["paired"] else "") +
                          "reorder_compress_streams re-blocking (src/reorder_compress_streams.cpp:83-361), streams resident in HBM",
                 "gpu_ms": sum(rb_dev) / len(rb_dev), "ms_with_d2h_of_blocks": sum(rb_wall) / len(rb_wall),
                 "blocks": int(blk.num_blocks), "d2h_bytes": int(sum(blk.size[i] for i in range(capi.NUM_BLOCK_STREAMS)))}

    # ---- e2e_files: the plugin call itself -- what the reference's own call_reorder / call_encoder boundary is: the
    # files preprocess leaves in temp_dir in, the encoder's stream files out (spring_b200_reorder_encode_files, persistent
    # context, temp_dir on /dev/shm like the reference arm) -------------------------------------------------------------
    e2e_files = None
    if world == 1 and not args.no_files_leg:
        base = "/dev/shm" if os.path.isdir("/dev/shm") else None
        d = tempfile.mkdtemp(prefix="spring_b200_files_", dir=base)
        try:
            h_packed = d_reads.cpu().numpy().view(np.uint64)
            h_lengths = d_lens.cpu().numpy().view(np.uint16)
            t_files, in_bytes, out_bytes = [], 0, 0
            for i in range(1 + min(args.steps, 3)):
                for f in os.listdir(d):
                    os.remove(os.path.join(d, f))
                cpy = dnaio.write_hotpath_inputs(d, h_packed, h_lengths, max_readlen=L, n_seqs=di.n_seqs, order_n=order_n,
                                                 num_reads=n_local, paired_split=di.num_clean[0] if cfg["paired"] else None, num_thr=8)
                os.remove(os.path.join(d, "cp_in.bin"))
                in_bytes = sum(os.path.getsize(os.path.join(d, f)) for f in os.listdir(d))
                t0 = time.perf_counter()
                ctx.reorder_encode_files(d, capi.CP.from_buffer_copy(cpy.pack()), args.chains)
                if i:
                    t_files.append(time.perf_counter() - t0)
                out_bytes = sum(os.path.getsize(os.path.join(d, f)) for f in os.listdir(d))
            ms_files = 1e3 * sum(t_files) / len(t_files)
            e2e_files = {"value": total / (ms_files * 1e-3) / 1e6, "unit": "Mreads/s", "ms_per_step": ms_files, "steps": len(t_files),
                         "input_file_bytes": int(in_bytes), "output_file_bytes": int(out_bytes),
                         "what": "spring_b200_reorder_encode_files: input_clean_*.dna / input_N.dna / read_order_N.bin on /dev/shm -> "
                                 "read_seq.bin.<t>, read_pos.bin, ... on /dev/shm (mmap + threaded record copy, H2D, kernels, D2H, "
                                 "threaded file writes); the boundary the reference arm crosses"}
            del h_packed, h_lengths
        finally:
            shutil.rmtree(d, ignore_errors=True)

    # ---- roofline of the dominant kernel (k_chains) ---------------------------------------------------
    peak, peak_src = measured_peak_gbs()
    ck_ms = sum(chain_ms) / len(chain_ms)
    ck_bytes = chain_kernel_bytes(counted, n_owned, W)
    achieved = ck_bytes / (ck_ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "r02_chain_kernel_traffic.json")
    if os.path.exists(tp):
        tj = json.load(open(tp)).get(str(cfg_id), {})  # one ncu capture per config: dram__bytes_read.sum + dram__bytes_write.sum
        if tj.get("reads") == n_local and world == 1:
            traffic = tj.get("dram_bytes_per_launch")
    roofline = {"kernel": "k_chains", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src, "kernel_ms": ck_ms,
                "kernel_share_of_step": ck_ms / ms_dev, "algorithmic_bytes_per_launch": ck_bytes,
                "bytes_per_read": ck_bytes / max(n_owned, 1),
                "counted": {"how": "one untimed pass of the counting instantiation of k_chains on the same input",
                            "probes_seq": int(counted["probes_seq"]), "compares": int(counted["compares"]),
                            "ms_chain_kernel_counting": counted["ms_chain_kernel"]}}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    fastq_bytes_per_read = 2 * cfg["mean_len"] + 12
    line = {"metric": METRIC, "value": value, "unit": "Mreads/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_dev, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64",
            "data": "synthetic", "config": workload_config(cfg_id, cfg, world, n_local),
            "e2e": {"value": e2e_val, "unit": "Mreads/s", "ms_per_step": ms_e2e, "h2d_bytes_per_step": h2d_bytes,
                    "d2h_bytes_per_step": int(d2h_bytes)},
            "gpu_launches": int(launches), "roofline": roofline, "clocks": clocks.summary(),
            "stages_ms": {k: stats_acc[k] for k in stats_acc if k.startswith("ms_")},
            "chains": stats_acc["num_chains"], "rounds": stats_acc["rounds"], "unmatched": stats_acc["unmatched"],
            "contigs": stats_acc["contigs"], "contigs_stitched": stats_acc["contigs_stitched"],
            "mb_per_s_fastq": value * fastq_bytes_per_read, "verify": verify}
    if exchange_ms:  # rank 0's bucket + scatter kernels, count all-gather and the NCCL send / recv group, per step (inside ms_per_step)
        line["exchange_ms"] = exchange_avg
        line["shard_layout_rank0"] = keep.get("layout")
    if after is not None:
        line["after_encoder"] = after
    if e2e_files is not None:
        line["e2e_files"] = e2e_files
    if world == 1 and not args.no_cpu_baseline:
        ctx.close()
        del d_reads, d_lens, di
        torch.cuda.empty_cache()
        n_s = min(args.cpu_sample, n_local)
        if cfg["paired"]:
            n_s -= n_s & 1
        packed, lengths, dis = host_input(cfg, n_s)
        threads = os.cpu_count() or 1
        secs, kind = time_reference(packed, lengths, dis, threads)
        cores = threads if kind == "reference" else 1
        line["cpu_baseline"] = {"value": dis.num_reads / secs / 1e6, "unit": "Mreads/s", "cores": cores, "kind": kind,
                                "sample": f"{n_s} reads of config {cfg_id}'s generator on a {genome_len(cfg, n_s)} bp genome "
                                          f"(same read model and coverage as the workload), "
                                          f"{secs:.1f} s of call_reorder+call_encoder at -t {cores}"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
