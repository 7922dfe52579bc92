"""Multi-GPU dealing: the path shards by whole read batches with no exchange step.  The index is
replicated per GPU, batches (the 10 Mbp records of src/Pipeline.cpp:98-146, which must stay whole because mem_pestat is
per batch, src/Pipeline.cpp:579) are dealt round-robin to ranks, and output order is restored by start_idx the way
SamsReorder does (src/Pipeline.cpp:651-737).  No collective is on the data path; torch.distributed is used only to
gather results / timings at the end."""
from __future__ import annotations


def deal(n_batches, world_size, rank):
    """Indices of the batches rank `rank` processes (round-robin, as the MPI mode's pull scatter approximates,
    src/mpi_main.cpp:295-317)."""
    return list(range(rank, n_batches, world_size))


class Reorder:
    """Releases records in start_idx order (SamsReorder, src/Pipeline.cpp:651-737).  push() returns the records that
    became releasable."""

    def __init__(self, first_idx=0):
        self.next_idx = first_idx
        self.pending = {}

    def push(self, start_idx, n_reads, payload):
        self.pending[start_idx] = (n_reads, payload)
        out = []
        while self.next_idx in self.pending:
            n, p = self.pending.pop(self.next_idx)
            out.append((self.next_idx, p))
            self.next_idx += n
        return out

    def drained(self):
        return not self.pending
