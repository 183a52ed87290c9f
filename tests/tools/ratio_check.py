"""GPU box: compressed size + stage times of the reference vs the reference host with libspring_b200
spliced in (oracle/_ref/spring_b200_ref), same FASTQ, `-c -r --no-quality`."""
import os, re, subprocess, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import pyoracle as po
from spring_b200 import synth

n = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 2_000_000
threads = os.cpu_count() or 8
rs = synth.generate(n, 150, genome_len=n * 150 // 30, seed=3, sub_rate=0.005, device="cuda")
d = tempfile.mkdtemp(dir="/dev/shm")
fq = os.path.join(d, "in.fastq")
t0 = time.time(); synth.write_fastq(rs, fq); print(f"wrote {n} reads in {time.time()-t0:.1f}s")
runs = [("reference", po.REF_BIN, {}), ("b200 auto chains", po.SPLICE_BIN, {}), ("b200 + reblock", po.SPLICE2_BIN, {}),
        ("b200 1 chain", po.SPLICE_BIN, {"SPRING_B200_CHAINS": "1"})]
if "--quick" in sys.argv:
    runs = runs[:3]
if "--policy" in sys.argv:  # chain-count policies: reads per chain of the auto rule (reorder.cu:run_reorder)
    runs = [("reference", po.REF_BIN, {})] + [(f"b200 {r} reads/chain", po.SPLICE2_BIN, {"SPRING_B200_READS_PER_CHAIN": str(r)})
                                              for r in (256, 2048, 6400)]
if "--stitch" in sys.argv:  # contig stitching in the encoder (spring_b200_set_stitch) off / on at the default chain count
    runs = [("reference", po.REF_BIN, {})] + [(f"b200 stitch {m}", po.SPLICE2_BIN, {"SPRING_B200_STITCH": str(m)}) for m in (0, 1)]
    if "--all-chains" in sys.argv:  # every co-resident chain whatever the input size (256 reads per chain: round 1's rule), stitched or not
        runs = [("reference", po.REF_BIN, {})] + [(f"b200 256 reads/chain stitch {m}", po.SPLICE2_BIN, {"SPRING_B200_STITCH": str(m), "SPRING_B200_READS_PER_CHAIN": "256"}) for m in (0, 1)]
if "--gpus" in sys.argv:  # ratio drift of the multi-GPU partitioning: the spliced binary on 1, 2, ... GPUs of this box
    import torch
    counts = [g for g in (1, 2, 4, 8) if g <= torch.cuda.device_count()]
    runs = [("reference", po.REF_BIN, {})] + [(f"b200 {g} GPU", po.SPLICE2_BIN, {"SPRING_B200_GPUS": str(g)}) for g in counts]
for name, binary, env in runs:
    out = os.path.join(d, name.replace(" ", "_").replace("/", "_per_") + ".spring")
    t0 = time.time()
    r = subprocess.run([binary, "-c", "-r", "--no-quality", "-i", fq, "-o", out, "-t", str(threads), "-w", d], capture_output=True, text=True, env={**os.environ, **env})
    wall = time.time() - t0
    if r.returncode != 0:
        print(name, "FAILED", r.stdout[-500:], r.stderr[-500:]); continue
    steps = re.findall(r"^(\w[\w /]+?) (?:done!|\.\.\.)\nTime for this step: (\d+) s", r.stdout, re.M)
    reads = re.search(r"Reads:\s+(\d+) bytes", r.stdout).group(1)
    unm = re.search(r"(\d+) were unmatched", r.stdout)
    al = re.findall(r"(\d+) (?:singleton reads|reads with N) were aligned", r.stdout)
    print(f"{name:18s} wall {wall:6.1f}s  Reads: {int(reads):>10d} bytes  ({8*int(reads)/n/150:.4f} bits/base)  unmatched {unm.group(1) if unm else '?'}  archive {os.path.getsize(out)}")
    print("    stage seconds:", [(a.strip(), int(b)) for a, b in steps])
