#!/usr/bin/env python
"""bench.py -- hot-path throughput of the B200 backend (and, with --impl reference, of the reference's own CPU path).

Metric: read pairs/s of the BWA-MEM per-read compute path (SMEM seeding -> chaining -> banded extension, i.e. kflow
stages 2..4: reads in, mem_alnreg_t lists out) on the configuration BASELINE.json's metric is quoted on (config[3]): a
synthetic human-scale 3.1 Gbp reference in 24 contigs, FM-index (4.65 GB + 0.78 GB pac) replicated per GPU, and 2x150 bp
pairs at 1% substitutions.  The index is built in HBM by the library's CUDA builder (seconds; bit-identical to `bwa index`,
tests/test_index_build_gpu.py) before the timed region.  One step = one pass of the hot path over one device batch
(16 kflow records of 10 Mbp = 533,344 pairs by default).  Per rank the work is fixed (weak scaling): every rank builds and
holds its own index replica and processes its own read shard; there is no data-path collective.

  value  = pairs/s with the reads already resident in HBM when the timed region starts
  e2e    = pairs/s through the public stage API from pinned HOST buffers (H2D of the reads and D2H of the regions inside)
  roofline      for the dominant kernel of the step (seeding: 64-byte occ-block touches vs measured HBM GB/s)
  roofline_ext  extension kernel: executed DP cells (GCUPS) vs the measured DPX issue peak / 5 lane-instr per cell
  cpu_baseline  the compiled reference (oracle/_ref) on this box's host cores over a bounded sample of the same reads

Launch: python bench.py [--gpus N --steps K --warmup W]; for N>1 under torchrun (one rank per GPU).
"""
import argparse
import json
import os
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


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--ref-mbp", type=float, default=3100.0)
    ap.add_argument("--contigs", type=int, default=24)
    ap.add_argument("--host-index", action="store_true", help="build the index with the host builder (small references only)")
    ap.add_argument("--kflow-batches", type=int, default=16, help="10 Mbp records per device batch (= one step)")
    ap.add_argument("--read-len", type=int, default=150)
    ap.add_argument("--sub-rate", type=float, default=0.01)
    ap.add_argument("--cpu-sample-pairs", type=int, default=0, help="pairs in the CPU baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ext-microbench", action="store_true")
    ap.add_argument("--in-flight", type=int, default=2,
                    help="device batches in flight per GPU, each driven by its own host thread and stream (the FPGA stages give every "
                         "worker two Task buffers for the same reason, src/fpga/FPGAPipeline.cpp:374-386) in the end-to-end leg; the "
                         "resident leg (value, per-stage breakdown, rooflines) always has one batch in flight")
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


def make_workload(args, rank):
    from bwaflow_b200 import synth
    ref = synth.make_reference(int(args.ref_mbp * 1e6), n_contigs=args.contigs, seed=3100)
    pairs_per_record = -(-KFLOW_BATCH_BP // (2 * args.read_len))           # getKseqBatch: stop at >= 10 Mbp, even count
    n_pairs = pairs_per_record * args.kflow_batches
    reads = synth.simulate_pairs(ref, n_pairs, args.read_len, args.sub_rate, seed=3 + rank, first_id=rank * n_pairs)
    return ref, reads, n_pairs, pairs_per_record


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


def ext_microbench(ctx, peak_gcups, n_jobs_target=10_000_000):
    """BASELINE config[1]: ~10M ksw_extend2 jobs of the shape mem_chain2aln issues for 2x150 bp reads (150 bp queries split at a
    seed, w=100, default scoring), through the SWTask-shaped batch API.  The jobs are synthesised (synth.extension_jobs, no aligner,
    a few hundred thousand distinct jobs tiled); GCUPS counts the DP cells ksw_extend2 executes for them (exact kernel variant)."""
    from bwaflow_b200 import capi, synth
    ref = synth.make_reference(4_000_000, n_contigs=2, seed=2)
    reads = synth.simulate_pairs(ref, 100_000, 150, 0.01, seed=7)
    jobs, qp, tp = synth.extension_jobs(ref, reads)
    n0 = len(jobs["qlen"])
    rep = max(1, n_jobs_target // n0)
    # pinned host staging, as the FPGA tasks map their buffers (src/fpga/SWTask.cpp:30-60)
    ej = capi.pinned((n0 * rep,), capi.EXT_JOB)
    ej[:] = np.zeros(1, dtype=capi.EXT_JOB)[0]
    for k in ("qlen", "tlen", "w", "h0", "end_bonus", "zdrop"):
        ej[k] = np.tile(jobs[k], rep)
    ej["q_off"] = np.tile(jobs["q_off"], rep) + np.repeat(np.arange(rep, dtype=np.int64) * len(qp), n0)
    ej["t_off"] = np.tile(jobs["t_off"], rep) + np.repeat(np.arange(rep, dtype=np.int64) * len(tp), n0)
    qpp, tpp = capi.pinned((len(qp) * rep,), np.uint8), capi.pinned((len(tp) * rep,), np.uint8)
    qpp.reshape(rep, -1)[:] = qp
    tpp.reshape(rep, -1)[:] = tp
    eb = capi.ExtBatch(ctx, len(ej), len(qpp) + 1, len(tpp) + 1)
    eb.submit(ej, qpp, tpp)
    res, t_e2e = eb.collect()
    eb.rerun(counted=True)
    res2, tc = eb.collect()
    same = all((res2[k] == res[k]).all() for k in capi.EXT_RES.names) and all((res[k][:n0] == res[k][-n0:]).all() for k in capi.EXT_RES.names)
    cells = tc.work_units
    ks = []
    for _ in range(5):
        eb.rerun()
        ks.append(eb.collect()[1].kernel_ms)
    k_ms = float(np.mean(ks[1:]))
    gc = cells / (k_ms * 1e-3) / 1e9
    out = {"workload": "BASELINE config[1] shape: %d synthetic extension jobs (%d distinct), mean qlen %.1f tlen %.1f, w=100, default scoring"
                       % (len(ej), n0, float(jobs["qlen"].mean()), float(jobs["tlen"].mean())),
           "gcups": gc, "peak_gcups": peak_gcups, "frac": gc / peak_gcups, "cells": cells, "kernel_ms": k_ms,
           "e2e_gcups": cells / (t_e2e.total_ms * 1e-3) / 1e9, "e2e_ms": t_e2e.total_ms,
           "h2d_bytes": int(ej.nbytes + qpp.nbytes + tpp.nbytes), "d2h_bytes": int(res.nbytes), "jobs_per_s": len(ej) / (k_ms * 1e-3),
           "variants_agree": bool(same)}
    eb.close()
    return out


def cpu_hot_path(R, reads, lo, hi):
    """stages 2 and 4 of the reference on reads [lo,hi): SeqsToChains::compute + ChainsToRegions::compute"""
    b = R.batch(reads.slice(lo, hi), lo)
    b.seq2chains()
    b.chains2regions()
    b.free()


def build_index(args, ref, device):
    """FM-index arrays for the run: the CUDA builder (default; the only way to have a 3.1 Gbp index in a bench run) or the host one."""
    from bwaflow_b200 import capi
    return capi.BuiltIndex(ref.concat(), ref.lengths, device=None if args.host_index else device)


def run_reference_cpu(args, ref, reads, n_pairs_sample, built):
    """Times the compiled reference's hot path (oracle/_ref) with one thread per host core over a bounded sample.  The reference
    loads the index with its own bwa_idx_load from files in bwa's on-disk format, written from the arrays the builder produced
    (`bwa index` itself would need hours for 3.1 Gbp; the arrays are bit-identical, tests/test_index_build*.py)."""
    from oracle import api
    from concurrent.futures import ThreadPoolExecutor
    import shutil
    td = tempfile.mkdtemp(prefix="bwaflow_ref_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    prefix = os.path.join(td, "ref")
    v = built.view
    api.write_index_files(prefix, v.primary, list(v.L2), v.seq_len, built.bwt_array(), built.sa_array(), v.sa_intv,
                          built.pac_array(), v.l_pac, ref.names, [int(x) for x in ref.lengths])
    R = api.Ref(prefix, paired=True)
    shutil.rmtree(td, ignore_errors=True)
    cores = os.cpu_count() or 1
    n_reads = 2 * n_pairs_sample
    chunk = max(2, (n_reads // (cores * 4)) & ~1)
    cuts = list(range(0, n_reads, chunk)) + [n_reads]

    def one_pass():
        t0 = time.perf_counter()
        with ThreadPoolExecutor(cores) as ex:
            list(ex.map(lambda i: cpu_hot_path(R, reads, cuts[i], cuts[i + 1]), range(len(cuts) - 1)))
        return time.perf_counter() - t0
    return R, one_pass, cores


def main():
    args = parse()
    rank, world, local = dist_env()
    if world > 1 and args.impl == "ours":
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if args.impl == "reference" and rank != 0:
        return
    ref, reads, n_pairs, pairs_per_record = make_workload(args, rank)
    workload = {"workload": "BASELINE config[3] shape: synthetic %.0f Mbp reference (%d contigs, FM-index replicated per GPU), 2x%dbp pairs @%.0f%% subst"
                            % (args.ref_mbp, args.contigs, args.read_len, args.sub_rate * 100),
                "pairs_per_step_per_gpu": n_pairs, "kflow_records_per_step": args.kflow_batches, "batches_in_flight_e2e": max(1, args.in_flight),
                "l2": "inputs larger than L2: %.0f MB FM-index + sampled SA + pac (random access), %.0f MB of reads per step and >1 GB of "
                      "per-step interval/region scratch rewritten every step, against a 126 MB L2"
                      % (args.ref_mbp * 1.75, 2 * n_pairs * args.read_len / 1e6)}

    if args.impl == "reference":
        if rank != 0:
            return
        sample = args.cpu_sample_pairs or max(pairs_per_record, n_pairs // 4)
        t0 = time.time()
        built = build_index(args, ref, local)
        workload["index_build_s"] = time.time() - t0
        R, one_pass, cores = run_reference_cpu(args, ref, reads, sample, built)
        for _ in range(min(args.warmup, 1)):
            one_pass()
        ts = [one_pass() for _ in range(args.steps)]
        sec = float(np.mean(ts))
        val = sample / sec
        line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "pairs/s", "n_gpus": args.gpus, "steps": args.steps,
                "warmup": min(args.warmup, 1), "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "int32", "data": "synthetic", "config": workload,
                "cpu_baseline": {"value": val, "unit": "pairs/s", "cores": cores, "kind": "reference",
                                 "sample": "%d pairs per step, stages 2+4 of the compiled reference (oracle/_ref), one thread per core" % sample},
                "e2e": {"value": val, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
        print(json.dumps(line))
        return

    # ------------------------------------------------------------------ our arm
    from bwaflow_b200 import capi, stages
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
    rec = stages.seqs_record(rank * reads.n, bases, read_off)

    def run_steps(fn, n_steps, n_workers):
        """n_steps passes of the hot path in total, pulled by n_workers host threads (worker i drives batch i)"""
        if n_workers == 1:
            return [fn(hps[0]) for _ in range(n_steps)]
        import itertools
        ctr, lock, out = itertools.count(), threading.Lock(), []

        def work(i):
            while True:
                with lock:
                    k = next(ctr)
                if k >= n_steps:
                    return
                out.append(fn(hps[i]))
        th = [threading.Thread(target=work, args=(i,)) for i in range(n_workers)]
        for t in th:
            t.start()
        for t in th:
            t.join()
        return out

    def step_e2e(h):
        return h.run(rec)                  # pinned host reads in, regions out (H2D + D2H inside)

    def step_resident(h):
        return h.rerun()[2]                # reads already in HBM from the previous submit

    for h in hps:
        out = step_e2e(h)                  # first touch: allocations, local-memory setup, capacity growth
    n_regs = len(out["alnreg"])
    n_exact_reads = hps[-1].batch.exact_reads()
    n_chk = min(reads.n, 100_000)          # kept for the parity check against the reference in the CPU-baseline leg
    chk_off = np.array(out["alnreg_off"][:n_chk + 1], dtype=np.int64)
    chk_regs = np.array(out["alnreg"][:chk_off[-1]])
    tc = hp.rerun(counted=True)[2]         # counting instances, untimed: the reference's executed DP cells and occ-block touches for this batch
    cells, blocks_smem = tc.cells, tc.smem_blocks
    nw = len(hps)
    # ---- resident pass: one batch in flight, per-stage CUDA events on its stream -> value, breakdown and rooflines
    # (more batches in flight only add contention here: there is no copy to hide; measured 8.8 / 8.7 / 8.5 M pairs/s at 1 / 2 / 3)
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
    run_steps(step_e2e, max(min(args.warmup, 2), nw), nw)
    barrier(world)
    sw.start()
    run_steps(step_e2e, args.steps, nw)
    ms_e2e = sw.stop()
    clk = clocks.stop() if rank == 0 else None
    barrier(world)
    ms_e2e = barrier_max(ms_e2e, world)

    total_pairs = n_pairs * world
    ms_step = ms_res / args.steps
    ms_step_one = ms_one / args.steps
    value = total_pairs / (ms_step * 1e-3)
    e2e_val = total_pairs / (ms_e2e / args.steps * 1e-3)
    if rank != 0:
        return
    # per-stage means over the timed steps (CUDA events on the batch's stream)
    k_smem = float(np.mean([p.smem_ms for p in parts]))
    k_pos = float(np.mean([p.seedpos_ms for p in parts]))
    k_chain = float(np.mean([p.chain_ms for p in parts]))
    k_c2a = float(np.mean([p.c2a_ms for p in parts]))
    blocks_pos = parts[-1].seedpos_blocks
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
    roofline = {"kernel": "smem_kernel (+ smem_p3_kernel: the seeding stage, timed and counted together)", "bound": "hbm", "achieved": smem_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": smem_gbs / hbm_peak,
                "traffic": traffic.get("smem_kernel"), "peak_source": "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback",
                "algorithmic_bytes_per_launch": blocks_smem * 64, "kernel_ms": k_smem, "share_of_step": k_smem / ms_step_one}
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
            "e2e": {"value": e2e_val, "unit": "pairs/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": int(args.steps * parts[-1].n_launches),
            "roofline": roofline, "roofline_ext": roofline_ext,
            "breakdown_ms": {"smem_stage": k_smem, "seedpos_stage": k_pos, "device_chaining": k_chain, "c2a_stage": k_c2a,
                             "step": ms_step, "e2e_step": ms_e2e / args.steps, "index_build_s": t_index},
            "counts": {"n_intv": parts[-1].n_intv, "n_pos": parts[-1].n_pos, "n_chains": parts[-1].n_chains,
                       "n_seeds": parts[-1].n_seeds, "n_regs": n_regs, "stage4_reads_on_exact_kernel": n_exact_reads},
            "seedpos_gbs": blocks_pos * 64 / (k_pos * 1e-3) / 1e9 if k_pos > 0 else None}
    if not args.no_ext_microbench:
        line["ext_microbench"] = ext_microbench(ctx, peak_gcups)
    if world == 1 and not args.no_cpu_baseline:
        sample = args.cpu_sample_pairs or n_pairs
        R, one_pass, cores = run_reference_cpu(args, ref, reads, sample, built=built)
        sec = one_pass()
        # full-scale parity: the reference's own regions for the first n_chk reads of the batch, byte for byte (untimed)
        rb = R.batch(reads.slice(0, n_chk), 0)
        rb.seq2chains()
        rb.chains2regions()
        r_regs, r_off = rb.regions()
        same_off = bool((np.asarray(r_off) == chk_off).all())
        line["parity_check"] = {"reads": int(n_chk), "regions": int(len(r_regs)), "offsets_identical": same_off,
                                "regions_identical": bool(same_off and np.asarray(r_regs).tobytes() == chk_regs.tobytes()),
                                "against": "oracle/_ref (the compiled reference: mem_seq2chain_wrapper + mem_chain2aln) on the same 3.1 Gbp index"}
        rb.free()
        line["cpu_baseline"] = {"value": sample / sec, "unit": "pairs/s", "cores": cores, "kind": "reference",
                                "sample": "%d pairs, stages 2+4 of the compiled reference (oracle/_ref: bwa/*.c + src/bwa_wrapper.cpp), "
                                          "one thread per core, %.1f s" % (sample, sec)}
    print(json.dumps(line))
    for h in hps:
        h.close()
    ctx.close()
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
