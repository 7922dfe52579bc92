"""Developer probe: stage 5 (bwaflow_sam_process) on full-size kflow records -- 66 668 reads of 2x150 bp (or 40 000 of 2x250 bp
indel-heavy) per record, regions from the fused hot path; prints the phase breakdown of a few warm calls.
usage: perf_sam.py [ref_mbp] [read_len] [records]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bwaflow_b200 import capi, stages, synth  # noqa: E402


def main():
    mbp = float(sys.argv[1]) if len(sys.argv) > 1 else 64
    rl = int(sys.argv[2]) if len(sys.argv) > 2 else 150
    n_rec = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    ref = synth.make_reference(int(mbp * 1e6), n_contigs=4, seed=64, repeat_copies=200, repeat_len=500)
    pairs = -(-10_000_000 // (2 * rl))
    B = capi.BuiltIndex(ref.concat(), ref.lengths, device=0)
    ctx = capi.Context(0, B.index(), keepalive=B)
    names = ["chr%d" % (i + 1) for i in range(len(ref.lengths))]
    fp = sb = None
    for k in range(n_rec):
        reads = synth.simulate_pairs_indel(ref, pairs, rl, 0.05, seed=5 + k) if rl >= 250 else synth.simulate_pairs(ref, pairs, rl, 0.01, seed=3 + k)
        nt4 = reads.nt4()
        if fp is None:
            fp = stages.FusedHotPath(ctx, reads.n, len(nt4) + 1024)
            sb = capi.SamBatch(ctx, reads.n, len(nt4) + 1024, names)
        out = fp.run(stages.seqs_record(0, nt4, reads.offsets))
        t0 = time.time()
        sam, pes, t = sb.process(nt4, reads.quals, reads.offsets, reads.names, out["alnreg"], out["alnreg_off"], k * reads.n)
        wall = time.time() - t0
        print("record %d: %d reads, %d regions -> %.1f MB SAM; hot path %.1f ms on the device; stage 5 %.1f ms (dedup %.1f rescue %.1f [device %.2f] pair %.1f cigar %.1f [device %.2f] text %.1f), "
              "python wall %.0f ms; %d patch, %d rescue windows (%d used), %d CIGAR jobs; %.2f M pairs/s through stage 5"
              % (k, reads.n, len(out["alnreg"]), t.sam_bytes / 1e6, fp.last.total_ms, t.total_ms, t.dedup_ms, t.matesw_ms, t.matesw_kernel_ms, t.pair_ms, t.cigar_ms, t.cigar_kernel_ms,
                 t.text_ms, wall * 1e3, t.n_patch_jobs, t.n_matesw_jobs, t.n_matesw_used, t.n_cigar_jobs, reads.n / 2 / t.total_ms / 1e3))
    fp.close(); sb.close(); ctx.close(); B.close()


if __name__ == "__main__":
    main()
