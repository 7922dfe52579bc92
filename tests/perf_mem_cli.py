"""Developer probe: the stand-alone driver bwaflow_mem end to end -- FASTQ files and bwa index files on disk in, SAM on disk out --
on a 64 Mbp reference with N pairs of 2x150 bp; prints the driver's own summary line (reads/s from the first record read to the last
SAM byte written, index load and device set-up reported separately).  usage: perf_mem_cli.py [pairs] [workers] [ref_mbp]"""
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bwaflow_b200 import capi, synth  # noqa: E402
from oracle import api  # noqa: E402  (index files in bwa's on-disk format)


def main():
    pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 500_000
    workers = sys.argv[2] if len(sys.argv) > 2 else "4"
    mbp = float(sys.argv[3]) if len(sys.argv) > 3 else 64
    td = tempfile.mkdtemp(dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    ref = synth.make_reference(int(mbp * 1e6), n_contigs=4, seed=64)
    B = capi.BuiltIndex(ref.concat(), ref.lengths, device=0)
    v = B.view
    api.write_index_files(os.path.join(td, "r"), v.primary, list(v.L2), v.seq_len, B.bwt_array(), B.sa_array(), v.sa_intv, B.pac_array(),
                          v.l_pac, ref.names, [int(x) for x in ref.lengths])
    B.close()
    reads = synth.simulate_pairs(ref, pairs, 150, 0.01, seed=3)
    t0 = time.time()
    for k, path in ((0, "r1.fq"), (1, "r2.fq")):
        with open(os.path.join(td, path), "wb") as f:
            for i in range(k, reads.n, 2):
                o0, o1 = int(reads.offsets[i]), int(reads.offsets[i + 1])
                f.write(b"@" + reads.names[i] + b"\n" + bytes(reads.bases[o0:o1]) + b"\n+\n" + bytes(reads.quals[o0:o1]) + b"\n")
    print("FASTQ written in %.1f s (%d MB)" % (time.time() - t0, sum(os.path.getsize(os.path.join(td, p)) for p in ("r1.fq", "r2.fq")) >> 20), flush=True)
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "bwaflow_b200", "host")], check=True)
    for w in workers.split(","):
        r = subprocess.run([os.path.join(ROOT, "bwaflow_b200", "bwaflow_mem"), "--use_gpu=1", "-t", w, "-o", os.path.join(td, "out.sam"),
                            os.path.join(td, "r"), os.path.join(td, "r1.fq"), os.path.join(td, "r2.fq")], capture_output=True, text=True)
        print("workers %s: rc %d; %s" % (w, r.returncode, r.stderr.strip().splitlines()[-1] if r.stderr.strip() else ""), flush=True)
    print("SAM: %d MB" % (os.path.getsize(os.path.join(td, "out.sam")) >> 20))


if __name__ == "__main__":
    main()
