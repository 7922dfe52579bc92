#!/usr/bin/env python
"""bench.py -- hot-path throughput of the B200 backend (and, with --impl reference, of the reference's own CPU path).

Metric: read pairs/s of the BWA-MEM per-read compute path (SMEM seeding -> chaining -> banded extension, i.e. kflow
stages 2..4: reads in, mem_alnreg_t lists out).  The default workload is the configuration BASELINE.json's metric is quoted
on (configs[3]): a synthetic human-scale 3.1 Gbp reference in 24 contigs, FM-index replicated per GPU, 2x150 bp pairs at
1% substitutions.  `--config` selects the other BASELINE shapes (c1 = configs[0], c3 = configs[2], c5 = configs[4]) and
`c4r`, the headline shape on a reference with a diverged repeat family and N runs; each prints the same line.
The index is built in HBM by the library's CUDA builder (seconds; bit-identical to `bwa index`,
tests/test_index_build_gpu.py) before the timed region.

One step = one pass of the hot path over one device batch of kflow records (10 Mbp each, src/Pipeline.cpp:98-146;
16 records = 533,344 pairs for 2x150 bp).  The record stream is global: record g holds pairs [g*P, (g+1)*P) and is generated
from seed(g); with N ranks the records are DEALT round-robin (rank r takes g = r, r+N, ..: dealer.deal) and every rank
restores record order among its in-flight batches with dealer.Reorder (the role of SamsReorder, src/Pipeline.cpp:651-737)
inside the timed end-to-end region.  Per rank the work is fixed (weak scaling); the index is replicated; there is no
data-path collective.

  value  = pairs/s with the reads already resident in HBM when the timed region starts
  e2e    = pairs/s through the public stage API from pinned HOST buffers (H2D of the reads and D2H of the regions inside)
  roofline      for the dominant kernels of the step (seeding: the reference's 64-byte occ-block touches vs measured HBM GB/s)
  roofline_ext  stage 4: executed DP cells (GCUPS) vs the measured DPX issue peak / 5 lane-instr per cell
  cpu_baseline  the compiled reference (oracle/_ref) on this box's host cores over a bounded sample of the same reads

`--impl reference` times the compiled reference (oracle/_ref: the reference's own bwa/*.c + src/bwa_wrapper.cpp) on the same
workload, same pairs per step, same config dict, one thread per host core.  Its index comes from a child process (the CUDA
builder writes bwa's on-disk files; `bwa index` itself needs hours at 3.1 Gbp), so the timed process maps nothing of the product.

Launch: python bench.py [--gpus N --steps K --warmup W]; for N>1 under torchrun (one rank per GPU).
"""
import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "read pairs/s, BWA-MEM per-read compute path (SMEM seeding + chaining + ksw_extend2 extension; reads -> mem_alnreg_t)"
KFLOW_BATCH_BP = 10_000_000            # src/Pipeline.cpp:146

CONFIGS = {
    # name: reference kwargs, read model, records per step, description
    "c4": dict(ref=dict(n_bp=3_100_000_000, n_contigs=24, seed=3100), read_len=150, model="sub", rate=0.01, records=16,
               desc="BASELINE config[3] shape: synthetic 3100 Mbp reference (24 contigs, FM-index replicated per GPU), 2x150bp pairs @1% subst"),
    "c4r": dict(ref=dict(n_bp=3_100_000_000, n_contigs=24, seed=3100, n_run_frac=0.001, repeat_copies=20000, repeat_len=500, repeat_div=0.03),
                read_len=150, model="sub", rate=0.01, records=16,
                desc="BASELINE config[3] shape on a harder reference: 3100 Mbp, 24 contigs, 20000 copies of a 500 bp repeat at 3% divergence, "
                     "0.1% N runs; 2x150bp pairs @1% subst"),
    "c1": dict(ref=dict(n_bp=5_000_000, n_contigs=1, seed=5), read_len=150, model="sub", rate=0.01, records=3,
               desc="BASELINE config[0] shape: synthetic 5 Mbp reference, 100k 2x150bp pairs @1% subst"),
    "c3": dict(ref=dict(n_bp=64_000_000, n_contigs=1, seed=64), read_len=150, model="sub", rate=0.01, records=16,
               desc="BASELINE config[2] shape: synthetic 64 Mbp (chr20-scale) reference, 2x150bp pairs @1% subst"),
    "c5": dict(ref=dict(n_bp=64_000_000, n_contigs=4, seed=55, repeat_copies=1000, repeat_len=600), read_len=250, model="indel", rate=0.05,
               records=4,
               desc="BASELINE config[4] shape: 2x250bp indel-heavy 5% error (40/30/30 subst/ins/del, chimeric reads, long indels), 64 Mbp reference"),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c4", choices=sorted(CONFIGS))
    ap.add_argument("--ref-mbp", type=float, default=0.0, help="override the reference size (developer runs)")
    ap.add_argument("--contigs", type=int, default=0)
    ap.add_argument("--host-index", action="store_true", help="build the index with the host builder (small references only)")
    ap.add_argument("--kflow-batches", type=int, default=0, help="10 Mbp records per device batch (= one step); 0 = the config's")
    ap.add_argument("--cpu-sample-pairs", type=int, default=0, help="pairs in the CPU baseline sample (0 = one step's worth)")
    ap.add_argument("--parity-reads", type=int, default=100_000, help="reads compared with the compiled reference in the CPU leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ext-microbench", action="store_true")
    ap.add_argument("--no-sam-stage", action="store_true", help="skip the stage-5 (RegionsToSam) measurement on one record")
    ap.add_argument("--ext-jobs", type=int, default=10_000_000, help="extension jobs dumped for the config[1] micro-benchmark")
    ap.add_argument("--in-flight", type=int, default=2,
                    help="device batches in flight per GPU, each driven by its own host thread and stream (the FPGA stages give every "
                         "worker two Task buffers for the same reason, src/fpga/FPGAPipeline.cpp:374-386) in the end-to-end leg; the "
                         "resident leg (value, per-stage breakdown, rooflines) always has one batch in flight")
    ap.add_argument("--write-index-files", default="", help=argparse.SUPPRESS)   # child mode of the reference arm
    return ap.parse_args()


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


class ClockSampler:
    """nvidia-smi clocks and throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu):
        self.gpu = gpu
        self.rows = []
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                       "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for nm, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


_POOL_REF = None


def _indel_chunk(job):
    from bwaflow_b200 import synth
    n, read_len, rate, seed, first_id = job
    return synth.simulate_pairs_indel(_POOL_REF, n, read_len, rate, seed=seed, first_id=first_id)


def make_record(cfg, ref, g, pairs_per_record, pool=None):
    """kflow record g of the global stream: pairs [g*P, (g+1)*P), generated from seed(g)."""
    from bwaflow_b200 import synth
    if cfg["model"] == "sub":
        return synth.simulate_pairs(ref, pairs_per_record, cfg["read_len"], cfg["rate"], seed=1000 + g, first_id=g * pairs_per_record)
    chunk = 1000
    jobs = [(min(chunk, pairs_per_record - lo), cfg["read_len"], cfg["rate"], 100000 + g * 1000 + lo // chunk, g * pairs_per_record + lo)
            for lo in range(0, pairs_per_record, chunk)]
    parts = pool.map(_indel_chunk, jobs) if pool else [_indel_chunk(j) for j in jobs]
    return concat_reads(parts)


def concat_reads(parts):
    from bwaflow_b200 import synth
    if len(parts) == 1:
        return parts[0]
    names, offs, base = [], [np.zeros(1, np.int64)], 0
    for p in parts:
        names += p.names
        offs.append(p.offsets[1:] + base)
        base += int(p.offsets[-1])
    return synth.ReadBatch(names, np.concatenate([p.bases for p in parts]), np.concatenate([p.quals for p in parts]), np.concatenate(offs))


def make_reference(args):
    from bwaflow_b200 import synth
    cfg = dict(CONFIGS[args.config])
    rk = dict(cfg["ref"])
    if args.ref_mbp:
        rk["n_bp"] = int(args.ref_mbp * 1e6)
    if args.contigs:
        rk["n_contigs"] = args.contigs
    return cfg, synth.make_reference(**rk)


def make_workload(args, rank, world):
    """The rank's share of the record stream for one step: records g = rank, rank + world, .. (dealer.deal)."""
    global _POOL_REF
    from bwaflow_b200 import dealer
    cfg, ref = make_reference(args)
    records = args.kflow_batches or cfg["records"]
    pairs_per_record = -(-KFLOW_BATCH_BP // (2 * cfg["read_len"]))           # getKseqBatch: stop at >= 10 Mbp, even count
    mine = dealer.deal(records * world, world, rank)
    pool = None
    if cfg["model"] == "indel":
        import multiprocessing as mp
        _POOL_REF = ref                   # inherited by the forked workers
        pool = mp.get_context("fork").Pool(min(os.cpu_count() or 1, 32))
    recs = [make_record(cfg, ref, g, pairs_per_record, pool) for g in mine]
    if pool:
        pool.close()
        pool.join()
    reads = concat_reads(recs)
    rec_ids = np.array(mine, dtype=np.int64)
    workload = {"workload": cfg["desc"] if not args.ref_mbp else cfg["desc"] + " [reference size overridden: %.0f Mbp]" % args.ref_mbp,
                "pairs_per_step_per_gpu": reads.n // 2, "kflow_records_per_step": records, "batches_in_flight_e2e": max(1, args.in_flight),
                "dealing": "record g -> rank g % N (round-robin), order restored per rank by start_idx",
                "l2": "FM-index (%.0f MB occ + %.2f GB suffix array + level tables) accessed at random, %.0f MB of reads per step and the "
                      "per-step interval/position/chain/region scratch (about 2 KB per pair, %.0f MB) rewritten every step, against a 126 MB L2: %s"
                      % (ref.total * 2 / 2 ** 20, ref.total * 16 / 1e9, reads.offsets[-1] / 1e6, reads.n * 1e-3,
                         "inputs larger than L2" if ref.total * 18 > 4 * 126e6 else
                         "the index alone is L2-sized (the reference's small CPU-runnable case); index + per-step streams exceed L2")}
    return cfg, ref, reads, rec_ids, pairs_per_record, workload


def barrier_max(x, world):
    """max over ranks (torch.distributed is only used for this and the barrier; nothing crosses ranks on the data path)"""
    if world == 1:
        return x
    import torch
    import torch.distributed as dist
    t = torch.tensor([x], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier(world):
    if world > 1:
        import torch.distributed as dist
        dist.barrier()


def dump_extension_jobs(n_target, n_threads):
    """BASELINE config[1]: extension jobs DUMPED from mem_chain2aln -- the compiled reference (oracle/_ref) runs stages 2+4 over
    2x150 bp reads of a 16 Mbp reference with every ksw_extend2 call observed by the link-time hook (oracle/ref_driver.cpp),
    until n_target distinct jobs are collected.  Returns (jobs, query pool, target pool)."""
    from concurrent.futures import ThreadPoolExecutor
    from bwaflow_b200 import synth
    from oracle import api
    ref = synth.make_reference(16_000_000, n_contigs=2, seed=2)
    td = tempfile.mkdtemp(prefix="bwaflow_c2_")
    ref.write_fasta(os.path.join(td, "r.fa"))
    api.build_index(os.path.join(td, "r.fa"), os.path.join(td, "r"))
    R = api.Ref(os.path.join(td, "r"))
    shutil.rmtree(td, ignore_errors=True)
    per_task = 20000                      # pairs per task: ~66k jobs

    def task(i):
        reads = synth.simulate_pairs(ref, per_task, 150, 0.01, seed=700 + i)
        R.hook_begin()
        b = R.batch(reads, 0)
        b.seq2chains()
        b.chains2regions()
        out = R.hook_end()
        b.free()
        return out
    parts, have, nxt = [], 0, 0
    with ThreadPoolExecutor(n_threads) as ex:
        while have < n_target:            # ~2 jobs per pair; rounds of tasks until the target is reached
            n_tasks = max(1, min(4 * n_threads, -(-(n_target - have) // (per_task * 19 // 10))))
            new = list(ex.map(task, range(nxt, nxt + n_tasks)))
            nxt += n_tasks
            parts += new
            have += sum(len(p[0]) for p in new)
    R.close()
    jobs = np.concatenate([p[0] for p in parts])
    qo = np.cumsum([0] + [len(p[1]) for p in parts])
    to = np.cumsum([0] + [len(p[2]) for p in parts])
    k = 0
    for i, p in enumerate(parts):
        jobs["q_off"][k:k + len(p[0])] += qo[i]
        jobs["t_off"][k:k + len(p[0])] += to[i]
        k += len(p[0])
    return jobs[:n_target] if len(jobs) > n_target else jobs, np.concatenate([p[1] for p in parts]), np.concatenate([p[2] for p in parts])


def ext_microbench(ctx, peak_gcups, n_jobs_target=10_000_000):
    """BASELINE config[1]: ~10M DISTINCT ksw_extend2 jobs dumped from mem_chain2aln (dump_extension_jobs; 150 bp queries, w=100,
    default scoring) through the SWTask-shaped batch API.  GCUPS counts the DP cells ksw_extend2 executes for them (counting kernel
    variant; its count equals the restatement's on a sample, checked here)."""
    from bwaflow_b200 import capi
    from oracle import api
    t0 = time.time()
    jobs, qp, tp = dump_extension_jobs(n_jobs_target, os.cpu_count() or 1)
    t_dump = time.time() - t0
    n = len(jobs)
    ej = capi.pinned((n,), capi.EXT_JOB)
    ej[:] = np.zeros(1, dtype=capi.EXT_JOB)[0]
    for k in ("qlen", "tlen", "w", "h0", "end_bonus", "zdrop", "q_off", "t_off"):
        ej[k] = jobs[k]
    qpp, tpp = capi.pinned((len(qp),), np.uint8), capi.pinned((len(tp),), np.uint8)
    qpp[:] = qp
    tpp[:] = tp
    eb = capi.ExtBatch(ctx, n, len(qpp) + 1, len(tpp) + 1)
    eb.submit(ej, qpp, tpp)
    res, t_e2e = eb.collect()
    same_as_ref = all((res[k] == jobs[k]).all() for k in ("score", "qle", "tle", "gtle", "gscore", "max_off"))
    eb.rerun(counted=True)
    res2, tc = eb.collect()
    same = all((res2[k] == res[k]).all() for k in capi.EXT_RES.names)
    cells = tc.work_units
    # cell count cross-check against the restatement on a sample of the jobs
    O = api.orc_lib()
    rng = np.random.default_rng(3)
    mat = np.array([[1 if i == j else -4 for j in range(4)] + [-1] for i in range(4)] + [[-1] * 5], dtype=np.int8).reshape(-1)
    sample = np.sort(rng.choice(n, size=min(n, 2000), replace=False))
    O.orc_counters_reset()
    for i in sample:
        j = jobs[i]
        api.orc_ksw_extend2(qp[j["q_off"]:j["q_off"] + j["qlen"]], tp[j["t_off"]:j["t_off"] + j["tlen"]], mat, 6, 1, 6, 1,
                            int(j["w"]), int(j["end_bonus"]), int(j["zdrop"]), int(j["h0"]))
    cells_sample_orc = int(O.orc_cells())
    sj = np.array(ej[sample])
    sb = capi.ExtBatch(ctx, len(sample), len(qpp) + 1, len(tpp) + 1)
    sb.submit(sj, qpp, tpp)
    sb.collect()
    sb.rerun(counted=True)
    cells_sample_gpu = int(sb.collect()[1].work_units)
    sb.close()
    ks = []
    for _ in range(5):
        eb.rerun()
        ks.append(eb.collect()[1].kernel_ms)
    k_ms = float(np.mean(ks[1:]))
    gc = cells / (k_ms * 1e-3) / 1e9
    out = {"workload": "BASELINE config[1]: %d distinct extension jobs dumped from mem_chain2aln of the compiled reference (2x150 bp @1%%, 16 Mbp "
                       "reference), mean qlen %.1f tlen %.1f, w=100, default scoring; dump took %.0f s on the host"
                       % (n, float(jobs["qlen"].mean()), float(jobs["tlen"].mean()), t_dump),
           "gcups": gc, "peak_gcups": peak_gcups, "frac": gc / peak_gcups, "cells": cells, "kernel_ms": k_ms,
           "e2e_gcups": cells / (t_e2e.total_ms * 1e-3) / 1e9, "e2e_ms": t_e2e.total_ms,
           "h2d_bytes": int(ej.nbytes + qpp.nbytes + tpp.nbytes), "d2h_bytes": int(res.nbytes), "jobs_per_s": n / (k_ms * 1e-3),
           "results_equal_reference": bool(same_as_ref), "variants_agree": bool(same),
           "cells_sample": {"jobs": int(len(sample)), "gpu": cells_sample_gpu, "oracle": cells_sample_orc}}
    eb.close()
    return out


def build_index(args, ref, device):
    """FM-index arrays for the run: the CUDA builder (default; the only way to have a 3.1 Gbp index in a bench run) or the host one."""
    from bwaflow_b200 import capi
    return capi.BuiltIndex(ref.concat(), ref.lengths, device=None if args.host_index else device)


def write_index_files(prefix, built, ref):
    """bwa's on-disk index from the builder's arrays, for the reference's own bwa_idx_load."""
    from oracle import api
    v = built.view
    codes = ref.concat()
    api.write_index_files(prefix, v.primary, list(v.L2), v.seq_len, built.bwt_array(), built.sa_array(), v.sa_intv,
                          built.pac_array(), v.l_pac, ref.names, [int(x) for x in ref.lengths],
                          codes=codes if (codes > 3).any() else None)


class ReferenceCPU:
    """The compiled reference's hot path (oracle/_ref) over reads [0, 2*n_pairs) with one thread per host core.  The records
    (bseq1_t arrays) are marshalled once, outside the timed passes; a pass is SeqsToChains::compute + ChainsToRegions::compute
    (src/Pipeline.cpp:333-348, 503-544) on every chunk."""

    def __init__(self, prefix, reads, n_pairs):
        from oracle import api
        self.R = api.Ref(prefix, paired=True)
        self.cores = os.cpu_count() or 1
        n_reads = 2 * n_pairs
        chunk = max(2, (n_reads // (self.cores * 4)) & ~1)
        cuts = list(range(0, n_reads, chunk)) + [n_reads]
        self.batches = [self.R.batch(reads.slice(cuts[i], cuts[i + 1]), cuts[i]) for i in range(len(cuts) - 1)]

    @staticmethod
    def _one(b):
        b.seq2chains()
        b.chains2regions()

    def one_pass(self):
        from concurrent.futures import ThreadPoolExecutor
        t0 = time.perf_counter()
        with ThreadPoolExecutor(self.cores) as ex:
            list(ex.map(self._one, self.batches))
        return time.perf_counter() - t0

    def close(self):
        for b in self.batches:
            b.free()
        self.R.close()


def reference_arm(args, local):
    """--impl reference: nothing of the product is mapped into this process; a child builds the index files."""
    cfg, ref, reads, rec_ids, ppr, workload = make_workload(args, 0, 1)
    n_pairs = reads.n // 2
    td = tempfile.mkdtemp(prefix="bwaflow_ref_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    prefix = os.path.join(td, "ref")
    t0 = time.time()
    cmd = [sys.executable, os.path.abspath(__file__), "--write-index-files", prefix, "--config", args.config]
    if args.ref_mbp:
        cmd += ["--ref-mbp", str(args.ref_mbp)]
    if args.contigs:
        cmd += ["--contigs", str(args.contigs)]
    if args.host_index:
        cmd += ["--host-index"]
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE")}
    env["LOCAL_RANK"] = str(local)
    subprocess.run(cmd, check=True, env=env)
    t_index = time.time() - t0
    sample = args.cpu_sample_pairs or n_pairs
    cpu = ReferenceCPU(prefix, reads, sample)
    shutil.rmtree(td, ignore_errors=True)
    for _ in range(args.warmup):
        cpu.one_pass()
    ts = [cpu.one_pass() for _ in range(args.steps)]
    sec = float(np.mean(ts))
    val = sample / sec
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "pairs/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int32", "data": "synthetic", "config": workload,
            "cpu_baseline": {"value": val, "unit": "pairs/s", "cores": cpu.cores, "kind": "reference", "per_core": val / cpu.cores,
                             "sample": "%d pairs per step, stages 2+4 of the compiled reference (oracle/_ref), one thread per core" % sample},
            "e2e": {"value": val, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0,
            "index_build_s": t_index}
    print(json.dumps(line))
    cpu.close()


_JSON_FD = None


def main():
    args = parse()
    rank, world, local = dist_env()
    if args.write_index_files:            # child of the reference arm: build the index (CUDA builder) and write bwa's files
        _, ref = make_reference(args)
        built = build_index(args, ref, local)
        write_index_files(args.write_index_files, built, ref)
        return
    if args.impl == "reference":
        if rank == 0:
            reference_arm(args, local)
        return
    if world > 1:
        # stdout carries ONE JSON line: whatever the libraries print there while the job runs (NCCL's version banner at communicator
        # creation) goes to stderr; the line itself is written to the saved descriptor
        global _JSON_FD
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cfg, ref, reads, rec_ids, pairs_per_record, workload = make_workload(args, rank, world)
    n_pairs = reads.n // 2

    # ------------------------------------------------------------------ our arm
    from bwaflow_b200 import capi, stages, dealer
    t0 = time.time()
    built = build_index(args, ref, local)
    index, keep = built.index(), built
    t_index = time.time() - t0
    ctx = capi.Context(local, index, keepalive=keep)
    nt4_np = reads.nt4()
    bases = capi.pinned(nt4_np.shape, np.uint8)
    bases[:] = nt4_np
    read_off = capi.pinned((reads.n + 1,), np.int64)
    read_off[:] = reads.offsets
    hps = [stages.FusedHotPath(ctx, reads.n, len(bases)) for _ in range(max(1, args.in_flight))]
    hp = hps[0]
    sw = capi.Stopwatch(ctx)
    reads_per_record = 2 * pairs_per_record
    rec = stages.seqs_record(int(rec_ids[0]) * reads_per_record, bases, read_off)
    reorder = dealer.Reorder(first_idx=0)
    rlock = threading.Lock()
    released = [0]

    def run_steps(fn, n_steps, n_workers):
        """n_steps passes of the hot path in total, pulled by n_workers host threads (worker i drives batch i)"""
        if n_workers == 1:
            return [fn(hps[0], k) for k in range(n_steps)]
        import itertools
        ctr, lock, out = itertools.count(), threading.Lock(), []

        def work(i):
            while True:
                with lock:
                    k = next(ctr)
                if k >= n_steps:
                    return
                out.append(fn(hps[i], k))
        th = [threading.Thread(target=work, args=(i,)) for i in range(n_workers)]
        for t in th:
            t.start()
        for t in th:
            t.join()
        return out

    def step_e2e(h, k):
        """pinned host reads in, regions out (H2D + D2H inside); the step's records leave in record order through the reorder
        buffer (step k's records follow step k-1's; batches in flight complete out of order)"""
        out = h.run(rec)
        off = out["alnreg_off"]
        with rlock:
            for j in range(len(rec_ids)):
                lo, hi = int(off[j * reads_per_record]), int(off[min((j + 1) * reads_per_record, reads.n)])
                seq = (k * len(rec_ids) + j) * reads_per_record       # position in this rank's output stream
                released[0] += len(reorder.push(seq, reads_per_record, (lo, hi)))
        return out

    def step_resident(h, k):
        return h.rerun()[2]                # reads already in HBM from the previous submit

    for h in hps:
        out = h.run(rec)                   # first touch: allocations, local-memory setup, capacity growth
    n_regs = len(out["alnreg"])
    n_exact_reads = hps[-1].batch.exact_reads()
    second_round = hps[-1].batch.second_round()
    n_chk = min(reads.n, args.parity_reads) & ~1   # kept for the parity check against the reference in the CPU-baseline leg
    chk_off = np.array(out["alnreg_off"][:n_chk + 1], dtype=np.int64)
    chk_regs = np.array(out["alnreg"][:chk_off[-1]])
    tc = hp.rerun(counted=True)[2]         # counting instances, untimed: the reference's executed DP cells and occ-block touches for this batch
    cells, blocks_smem = tc.cells, tc.smem_blocks
    nw = len(hps)
    # ---- resident pass: one batch in flight, per-stage CUDA events on its stream -> value, breakdown and rooflines
    run_steps(step_resident, args.warmup, 1)
    clocks = ClockSampler(local)          # rank 0 samples its own GPU; one nvidia-smi poller per box, not one per rank
    barrier(world)
    if rank == 0:
        clocks.start()
    sw.start()
    parts = run_steps(step_resident, args.steps, 1)
    ms_res = sw.stop()
    barrier(world)
    ms_one = ms_res
    ms_res = barrier_max(ms_res, world)
    # ---- end to end from pinned host buffers, nw batches in flight so that one batch's copies overlap the other's kernels
    n_w = max(min(args.warmup, 2), nw)
    run_steps(step_e2e, n_w, nw)
    assert reorder.drained()
    reorder.next_idx = 0
    released[0] = 0
    barrier(world)
    sw.start()
    run_steps(step_e2e, args.steps, nw)
    ms_e2e = sw.stop()
    clk = clocks.stop() if rank == 0 else None
    barrier(world)
    ms_e2e = barrier_max(ms_e2e, world)
    assert reorder.drained() and released[0] == args.steps * len(rec_ids), (released[0], args.steps, len(rec_ids))

    total_pairs = n_pairs * world
    ms_step = ms_res / args.steps
    ms_step_one = ms_one / args.steps
    value = total_pairs / (ms_step * 1e-3)
    e2e_val = total_pairs / (ms_e2e / args.steps * 1e-3)
    if rank != 0:
        for h in hps:
            h.close()
        ctx.close()
        import torch.distributed as dist
        dist.destroy_process_group()
        return
    # per-stage means over the timed steps (CUDA events on the batch's stream)
    k_smem = float(np.mean([p.smem_ms for p in parts]))
    k_pos = float(np.mean([p.seedpos_ms for p in parts]))
    k_chain = float(np.mean([p.chain_ms for p in parts]))
    k_c2a = float(np.mean([p.c2a_ms for p in parts]))
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    traffic = {}
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    except Exception:
        pass
    smem_gbs = blocks_smem * 64 / (k_smem * 1e-3) / 1e9
    roofline = {"kernel": "seeding stage: pack_reads + smem_sweep_kernel + smem_fast_p3_kernel (timed and counted together)", "bound": "hbm",
                "achieved": smem_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": smem_gbs / hbm_peak,
                "traffic": traffic.get("seeding_stage") if args.config == "c4" and not args.ref_mbp else None,
                "peak_source": "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback",
                "algorithmic_bytes_per_launch": blocks_smem * 64, "kernel_ms": k_smem, "share_of_step": k_smem / ms_step_one,
                "note": "achieved = the REFERENCE's 64-byte occ-block touches for this batch (counted by the exact instances, equal to the "
                        "oracle's count in the tests) per second of the production kernels, which serve most of those lookups from the level "
                        "tables and the 2-bit text instead of the index: DRAM traffic below the algorithmic bytes is the point"}
    dpx16, dpx32 = ctx.dpx_peak()
    gcups = cells / (k_c2a * 1e-3) / 1e9
    peak_gcups = dpx16 / 5.0 / 1e9
    roofline_ext = {"kernel": "stage 4: ext_pair_kernel / ext_kernel waves + c2a_replay + c2a_kernel (+prep/sort/compact)", "bound": "int-alu/dpx", "achieved": gcups, "peak": peak_gcups, "unit": "GCUPS", "frac": gcups / peak_gcups,
                    "cells_per_launch": cells, "kernel_ms": k_c2a, "share_of_step": k_c2a / ms_step_one,
                    "peak_source": "measured DPX s16x2 issue rate %.3g lane-instr/s / 5 lane-instr per cell (SURVEY 8d); s32 rate %.3g" % (dpx16, dpx32)}
    h2d = int(bases.nbytes + read_off.nbytes)
    d2h = int(n_regs * 88 + (reads.n + 1) * 8)
    line = {"metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32",
            "data": "synthetic", "config": workload, "clocks": clk,
            "e2e": {"value": e2e_val, "unit": "pairs/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "records_released_in_order": released[0]},
            "gpu_launches": int(args.steps * parts[-1].n_launches),
            "roofline": roofline, "roofline_ext": roofline_ext,
            "breakdown_ms": {"smem_stage": k_smem, "seedpos_stage": k_pos, "device_chaining": k_chain, "c2a_stage": k_c2a,
                             "step": ms_step, "e2e_step": ms_e2e / args.steps},
            "index_build_s": t_index, "dpx_peak_lane_ops_per_s": {"s16x2": dpx16, "s32": dpx32},
            "counts": {"n_intv": parts[-1].n_intv, "n_pos": parts[-1].n_pos, "n_chains": parts[-1].n_chains,
                       "n_seeds": parts[-1].n_seeds, "n_regs": n_regs, "stage4_reads_left_by_first_replay": second_round[0], "stage4_second_round_jobs": second_round[1],
                       "stage4_reads_on_exact_kernel": n_exact_reads}}
    # ---- stage 5 (RegionsToSam) on the first record of the step: regions of the hot path in, SAM text out (one worker; in the
    # pipeline every stage worker runs its own record).  Not part of `value` / `e2e`: the headline metric is the path up to mem_alnreg_t.
    sam_text = None
    if world == 1 and not args.no_sam_stage:
        n_sr = min(reads.n, reads_per_record)
        s_off = np.array(read_off[:n_sr + 1], dtype=np.int64)
        r_off = np.array(out["alnreg_off"][:n_sr + 1], dtype=np.int64)
        s_regs = np.array(out["alnreg"][:r_off[-1]])
        ctg_names = ["chr%d" % (i + 1) for i in range(len(ref.lengths))] if not hasattr(ref, "names") else [n if isinstance(n, str) else n.decode() for n in ref.names]
        sb = capi.SamBatch(ctx, n_sr, int(s_off[-1]) + 16, ctg_names)
        start_idx = int(rec_ids[0]) * reads_per_record
        tms = []
        for it in range(4):
            sam_text, _, st = sb.process(bases[:s_off[-1]], reads.quals[:s_off[-1]], s_off, reads.names[:n_sr], s_regs, r_off, start_idx)
            if it:
                tms.append(st)
        mean = lambda k: float(np.mean([getattr(t, k) for t in tms]))
        line["sam_stage"] = {"what": "bwaflow_sam_process on one kflow record: mem_sort_dedup_patch, mem_pestat, mate rescue, pairing, CIGAR / MD / NM, SAM text",
                             "reads": int(n_sr), "ms": mean("total_ms"), "pairs_per_s_one_worker": n_sr / 2 / (mean("total_ms") * 1e-3),
                             "phases_ms": {k: mean(k + "_ms") for k in ("dedup", "matesw", "pair", "cigar", "text")},
                             "device_ms": {"ksw_align2": mean("matesw_kernel_ms"), "mem_reg2aln": mean("cigar_kernel_ms")},
                             "patch_alignments": int(tms[-1].n_patch_jobs), "rescue_windows": int(tms[-1].n_matesw_jobs),
                             "rescues_used": int(tms[-1].n_matesw_used), "cigar_jobs": int(tms[-1].n_cigar_jobs), "sam_bytes": int(tms[-1].sam_bytes)}
        sb.close()
    if not args.no_ext_microbench:
        line["ext_microbench"] = ext_microbench(ctx, peak_gcups, args.ext_jobs)
    if world == 1 and not args.no_cpu_baseline:
        sample = args.cpu_sample_pairs or n_pairs
        td = tempfile.mkdtemp(prefix="bwaflow_ref_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
        write_index_files(os.path.join(td, "ref"), built, ref)
        cpu = ReferenceCPU(os.path.join(td, "ref"), reads, sample)
        shutil.rmtree(td, ignore_errors=True)
        sec = cpu.one_pass()
        # full-scale parity: the reference's own regions for the first n_chk reads of the batch, byte for byte (untimed), with its
        # ksw_extend2 calls observed by the hook
        cpu.R.hook_begin()
        rb = cpu.R.batch(reads.slice(0, n_chk), 0)
        rb.seq2chains()
        rb.chains2regions()
        jobs = cpu.R.hook_end()[0]
        r_regs, r_off = rb.regions()
        same_off = bool((np.asarray(r_off) == chk_off).all())
        line["parity_check"] = {"reads": int(n_chk), "regions": int(len(r_regs)), "offsets_identical": same_off,
                                "regions_identical": bool(same_off and np.asarray(r_regs).tobytes() == chk_regs.tobytes()),
                                "ksw_extend2_calls": int(len(jobs)), "second_band_tries": int((jobs["w"] > 100).sum()),
                                "max_off_max": int(jobs["max_off"].max()) if len(jobs) else 0,
                                "against": "oracle/_ref (the compiled reference: mem_seq2chain_wrapper + mem_chain2aln) on the same index"}
        rb.free()
        if sam_text is not None:
            # SAM parity at full record size: the reference's stages 2-5 on the same record (same start_idx), byte for byte (untimed)
            n_sr = len(sam_text)
            rb = cpu.R.batch(reads.slice(0, n_sr), int(rec_ids[0]) * reads_per_record)
            rb.seq2chains()
            rb.chains2regions()
            r_sam, r_soff = rb.regions2sam()
            line["sam_stage"]["sam_identical_to_reference"] = bool(bytes(r_sam) == b"".join(sam_text))
            line["sam_stage"]["reference_sam_bytes"] = int(len(r_sam))
            rb.free()
        line["cpu_baseline"] = {"value": sample / sec, "unit": "pairs/s", "cores": cpu.cores, "kind": "reference", "per_core": sample / sec / cpu.cores,
                                "sample": "%d pairs, stages 2+4 of the compiled reference (oracle/_ref: bwa/*.c + src/bwa_wrapper.cpp), "
                                          "one thread per core, %.1f s" % (sample, sec)}
        cpu.close()
    if _JSON_FD is not None:
        os.write(_JSON_FD, (json.dumps(line) + "\n").encode())
    else:
        print(json.dumps(line))
    for h in hps:
        h.close()
    ctx.close()
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
