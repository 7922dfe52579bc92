"""Host-side mirrors of the two kflow stage slots the CUDA backend fills, in the shape of the FPGA stages
(src/fpga/FPGAPipeline.h:14-48), plus the host hop between them:

    SeqsToChainsCUDA.compute(SeqsRecord)      -> ChainsRecord{bwtintvs, tag=0}   (SeqsToChainsFPGA::compute, FPGAPipeline.cpp:582)
    ChainsPipe.compute(ChainsRecord)          -> ChainsRecord{chains,  tag=1}    (src/Pipeline.cpp:408-438)
    ChainsToRegionsCUDA.compute(ChainsRecord) -> RegionsRecord{alnreg}           (ChainsToRegionsFPGA::compute, FPGAPipeline.cpp:367)

Records are plain dicts with the reference's field names (src/Pipeline.h:39-66); arrays are flat numpy arrays
(seqs -> `bases` nt4 + `read_off`).  All device work goes through the C ABI (capi); nothing here falls back to a CPU path.
"""
from __future__ import annotations

import numpy as np

from . import capi


def seqs_record(start_idx, bases_nt4, read_off):
    """SeqsRecord (src/Pipeline.h:46-51) after the in-place nt4 conversion of stage 2 (src/bwa_wrapper.cpp:124-125)."""
    read_off = np.ascontiguousarray(read_off, dtype=np.int64)
    return {"name": "SeqsRecord", "start_idx": int(start_idx), "batch_num": len(read_off) - 1,
            "bases": np.ascontiguousarray(bases_nt4, dtype=np.uint8), "read_off": read_off}


class SeqsToChainsCUDA:
    """Stage slot 2.  Output keeps the intervals (tag=0) like the FPGA stage, and additionally carries the SA positions
    of every seed (computed on the device) so that ChainsPipe does not walk the index on the host."""

    def __init__(self, ctx: capi.Context, max_reads, max_bases):
        self.ctx = ctx
        self.batch = capi.SmemBatch(ctx, max_reads, max_bases)
        self.last = {}

    def compute(self, rec):
        b = self.batch
        b.submit(rec["bases"], rec["read_off"])
        intv, intv_off, t1 = b.collect()
        pos, pos_off, t2 = b.seed_positions()
        self.last = {"smem": t1, "seedpos": t2}
        out = dict(rec)
        out.update(name="ChainsRecord", tag=0, bwtintvs=intv, bwtintv_off=intv_off, seed_pos=pos, seed_pos_off=pos_off,
                   chains=None)
        return out

    def close(self):
        self.batch.close()


class ChainsPipe:
    """The host hop (src/Pipeline.cpp:408-438): pass-through for tag=1, chaining + filtering for tag=0."""

    def __init__(self, index_like, ctg_off, ctg_len, opt: capi.ChainOpt | None = None, ctg_alt=None, n_threads=0):
        self.l_pac = int(index_like.l_pac)
        self.ctg_off, self.ctg_len, self.ctg_alt = ctg_off, ctg_len, ctg_alt
        self.opt = opt
        self.n_threads = n_threads

    def compute(self, rec):
        if rec.get("tag") == 1:
            return rec
        chains, seeds, chain_off = capi.chain_build(self.l_pac, self.ctg_off, self.ctg_len, rec["read_off"], rec["bwtintvs"],
                                                    rec["bwtintv_off"], rec["seed_pos"], rec["seed_pos_off"], self.opt,
                                                    self.ctg_alt, self.n_threads)
        out = dict(rec)
        out.update(tag=1, chains=chains, seeds=seeds, chain_off=chain_off, bwtintvs=None, bwtintv_off=None,
                   seed_pos=None, seed_pos_off=None)
        return out


class ChainsToRegionsCUDA:
    """Stage slot 4."""

    def __init__(self, ctx: capi.Context, max_reads, max_bases, max_chains=None, max_seeds=None):
        self.ctx = ctx
        self.cap = [max_reads, max_bases, max_chains or 4 * max_reads, max_seeds or 16 * max_reads]
        self.batch = capi.C2ABatch(ctx, *self.cap)
        self.last = {}

    def _fit(self, n_chains, n_seeds):
        if n_chains > self.cap[2] or n_seeds > self.cap[3]:
            self.batch.close()
            self.cap[2] = max(self.cap[2], n_chains + n_chains // 4)
            self.cap[3] = max(self.cap[3], n_seeds + n_seeds // 4)
            self.batch = capi.C2ABatch(self.ctx, *self.cap)
            return True
        return False

    def compute(self, rec):
        assert rec.get("tag") == 1, "ChainsToRegionsCUDA needs chains (run ChainsPipe first)"
        self._fit(len(rec["chains"]), len(rec["seeds"]))
        b = self.batch
        b.submit(rec["bases"], rec["read_off"], rec["chains"], rec["chain_off"], rec["seeds"])
        regs, reg_off, t = b.collect()
        self.last = {"c2a": t}
        return {"name": "RegionsRecord", "start_idx": rec["start_idx"], "batch_num": rec["batch_num"], "bases": rec["bases"],
                "read_off": rec["read_off"], "chains": None, "alnreg": regs, "alnreg_off": reg_off}

    def close(self):
        self.batch.close()


class FusedHotPath:
    """reads -> alignment regions on the device only (bwaflow_hot_*): stage 2 leaves its chains in HBM (tag=1 record whose
    `chains` is a device handle), ChainsPipe passes through, stage 4 consumes them.  Same results as HotPath."""

    def __init__(self, ctx: capi.Context, max_reads, max_bases, chain_opt=None, ctg_alt=None):
        self.batch = capi.HotBatch(ctx, max_reads, max_bases, chain_opt, ctg_alt)
        self.last = None

    def run(self, rec):
        self.batch.submit(rec["bases"], rec["read_off"])
        regs, reg_off, t = self.batch.collect()
        self.last = t
        return {"name": "RegionsRecord", "start_idx": rec["start_idx"], "batch_num": rec["batch_num"], "bases": rec["bases"],
                "read_off": rec["read_off"], "chains": None, "alnreg": regs, "alnreg_off": reg_off}

    def rerun(self, counted=False):
        self.batch.rerun(counted)
        regs, reg_off, t = self.batch.collect()
        self.last = t
        return regs, reg_off, t

    def close(self):
        self.batch.close()


class HotPath:
    """reads -> alignment regions: stage 2, the host hop and stage 4 chained the way compute_flow wires them
    (src/main.cpp:302-335), with the chaining hop on the host CPU as in the FPGA flow."""

    def __init__(self, ctx: capi.Context, index_like, ctg_off, ctg_len, max_reads, max_bases, chain_opt=None, n_threads=0):
        self.s2 = SeqsToChainsCUDA(ctx, max_reads, max_bases)
        self.hop = ChainsPipe(index_like, ctg_off, ctg_len, chain_opt, None, n_threads)
        self.s4 = ChainsToRegionsCUDA(ctx, max_reads, max_bases)

    def run(self, rec):
        return self.s4.compute(self.hop.compute(self.s2.compute(rec)))

    def timings(self):
        d = {}
        d.update(self.s2.last)
        d.update(self.s4.last)
        return d

    def close(self):
        self.s2.close()
        self.s4.close()
