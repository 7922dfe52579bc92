"""Generates the committed golden fixtures from the COMPILED REFERENCE (oracle/_ref, i.e. the reference's own
bwa/*.c + src/bwa_wrapper.cpp built in place).  Run here (where /root/reference exists):

    python tests/golden/make_golden.py

Outputs tests/golden/<name>.npz with, for a seeded synthetic reference + read set:
  ref_codes, ctg_len            the reference (codes 0..4) -> FASTA -> `bwa index` is reproducible from it
  bases, quals, offsets, names  the reads
  intv, intv_off                mem_collect_intv_new output per read          (K1 contract)
  chains, seeds, chain_off      mem_seq2chain_wrapper output per read         (stage-2 output)
  regs, reg_off                 mem_chain2aln output per read                 (stage-4 output)
  jobs, qpool, tpool            every ksw_extend2 call made by mem_chain2aln with its six outputs
  job_cells, cells_total        DP cells each of those calls executes (bwa/ksw.c:424-448 iterations), counted by the
                                restatement (oracle.c) after checking that it returns the reference's six outputs
  blocks_total                  64-byte occ-block touches of mem_collect_intv_new over all reads (oracle.c counter)
  sam, sam_off                  per-read SAM text out of RegionsToSam
  sa_k, sa_v                    bwt_sa known answers
"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from bwaflow_b200 import synth  # noqa: E402
from oracle import api  # noqa: E402

CASES = {
    # name: (reference kwargs, read generator, kwargs)
    "c1_150bp": (dict(n_bp=400_000, n_contigs=3, seed=5, n_run_frac=0.001, repeat_copies=40), "sub",
                 dict(n_pairs=600, read_len=150, sub_rate=0.01, seed=1, n_rate=0.0005)),
    "c5_250bp_indel": (dict(n_bp=300_000, n_contigs=2, seed=55, repeat_copies=20, repeat_len=600), "indel",
                       dict(n_pairs=200, read_len=250, err=0.05, seed=5)),
}


def make(name, ref_kw, kind, read_kw, outdir):
    ref = synth.make_reference(**ref_kw)
    reads = synth.simulate_pairs(ref, **read_kw) if kind == "sub" else synth.simulate_pairs_indel(ref, **read_kw)
    with tempfile.TemporaryDirectory() as td:
        fa = os.path.join(td, "ref.fa")
        ref.write_fasta(fa)
        api.build_index(fa, os.path.join(td, "ref"))
        R = api.Ref(os.path.join(td, "ref"), paired=True)
        b = R.batch(reads, start_idx=0)
        intv, intv_off = b.intervals()
        b.seq2chains()
        chains, seeds, chain_off = b.chains()
        R.hook_begin()
        b.chains2regions()
        jobs, qpool, tpool = R.hook_end()
        regs, reg_off = b.regions()
        sam, sam_off = b.regions2sam()
        # work-unit counters of SURVEY 8(d): the reference does not count, the pinned restatement does
        O = api.Oracle.from_ref(R)
        mat = R.mat()
        job_cells = np.zeros(len(jobs), dtype=np.int64)
        for i, j in enumerate(jobs):
            O.reset()
            got = api.orc_ksw_extend2(qpool[j["q_off"]:j["q_off"] + j["qlen"]], tpool[j["t_off"]:j["t_off"] + j["tlen"]], mat,
                                      R.opt("o_del"), R.opt("e_del"), R.opt("o_ins"), R.opt("e_ins"), int(j["w"]),
                                      int(j["end_bonus"]), int(j["zdrop"]), int(j["h0"]))
            assert list(got) == [j[k] for k in ("score", "qle", "tle", "gtle", "gscore", "max_off")], i
            job_cells[i] = O.cells()
        O.reset()
        nt4 = reads.nt4()
        for i in range(reads.n):
            got = O.collect_intv(nt4[reads.offsets[i]:reads.offsets[i + 1]])
            assert got.tobytes() == intv[intv_off[i]:intv_off[i + 1]].tobytes(), i
        blocks_total = O.blocks()
        rng = np.random.default_rng(7)
        sa_k = rng.integers(0, R.view.seq_len + 1, size=512).astype(np.uint64)
        sa_v = np.array([R.bwt_sa(int(k)) for k in sa_k], dtype=np.uint64)
        names, name_off = reads.name_blob()
        np.savez_compressed(
            os.path.join(outdir, name + ".npz"),
            ref_codes=ref.concat(), ctg_len=ref.lengths,
            bases=reads.bases, quals=reads.quals, offsets=reads.offsets,
            names=np.frombuffer(names, dtype=np.uint8), name_off=name_off,
            intv=intv, intv_off=intv_off, chains=chains, seeds=seeds, chain_off=chain_off,
            regs=regs, reg_off=reg_off, jobs=jobs, qpool=qpool, tpool=tpool,
            job_cells=job_cells, cells_total=np.int64(job_cells.sum()), blocks_total=np.int64(blocks_total),
            sam=np.frombuffer(sam, dtype=np.uint8), sam_off=sam_off, sa_k=sa_k, sa_v=sa_v,
            opt=np.array([R.opt(k) for k in ("a", "b", "o_del", "e_del", "o_ins", "e_ins", "pen_clip5", "pen_clip3", "w", "zdrop")], dtype=np.int32),
        )
        print(name, "reads", reads.n, "intv", len(intv), "chains", len(chains), "regs", len(regs), "jobs", len(jobs),
              "zdrop-capable jobs", int((jobs["tlen"] > 0).sum()))
        b.free()
        R.close()


if __name__ == "__main__":
    out = os.path.dirname(os.path.abspath(__file__))
    for nm, (rk, kind, kw) in CASES.items():
        make(nm, rk, kind, kw, out)
