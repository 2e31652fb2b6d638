"""Multi-GPU plumbing: one process per GPU (torch.distributed), z steps or whole pose lists sharded, one
gather of the score table at the end — the B200 replacement of the reference's MPI decomposition
(tools/correlate.c:140-147 split of z over ranks, :295-356 MPI_Gatherv x5).  There is no collective on the
compute path: cells are independent given the replicated read-only tables.
"""
import numpy as np


def z_digit(index, L):
    nb, N = L + 1, 2 * L + 1
    return np.asarray(index).astype(np.int64) // (nb * nb * N ** 3)


def shard_z_ranges(index, L, znum, world):
    """Contiguous z ranges [lo, hi) per rank with roughly equal numbers of listed poses.

    The reference deals z steps round-robin to ranks; contiguous ranges give the same independence and let
    every rank keep its translated slabs for a range of z resident.  Ranks beyond the number of occupied z
    steps get an empty range."""
    z = z_digit(index, L)
    z = z[(z >= 0) & (z < znum)]
    per_z = np.bincount(z, minlength=znum).astype(np.int64)
    total = int(per_z.sum())
    cum = np.concatenate([[0], np.cumsum(per_z)])
    bounds = [0]
    for r in range(1, world):
        target = total * r / world
        b = int(np.searchsorted(cum, target, side="left"))
        b = min(max(b, bounds[-1]), znum)
        bounds.append(b)
    bounds.append(znum)
    return [(bounds[r], bounds[r + 1]) for r in range(world)]


def merge_tables(tables, owners):
    """tables[r] = (scores, c1, c2) of rank r over the FULL pose list (untouched outside its z range),
    owners[i] = rank that scored pose i (or -1).  Returns the merged table."""
    out = [np.array(t, copy=True) for t in tables[0]]
    for r in range(1, len(tables)):
        m = owners == r
        for k in range(3):
            out[k][m] = tables[r][k][m]
    return tuple(out)


def owners_of(index, L, ranges):
    z = z_digit(index, L)
    own = np.full(len(z), -1, dtype=np.int64)
    for r, (lo, hi) in enumerate(ranges):
        own[(z >= lo) & (z < hi)] = r
    return own


def all_gather_tables(local3, group=None):
    """local3: torch tensor [3][n] (scores, c1, c2) of this rank's own pose list -> [world][3][n] on every rank
    (NCCL on GPUs, gloo on CPU)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rows, n = local3.shape
    out = torch.empty((world * rows, n), dtype=local3.dtype, device=local3.device)  # concatenation along dim 0
    dist.all_gather_into_tensor(out, local3.contiguous(), group=group)
    return out.view(world, rows, n)


class RowShardGather:
    """ONE pose list split over the ranks by z range (strong scaling — the reference's MPI scheme, every rank filters the
    rows of its own z steps, tools/correlate.c:140-147,169-251) and the final gather of the score table into input
    order (MPI_Gatherv x5, :295-356) as ONE all-gather of the padded [3][rows] shards plus one index copy.

    Every rank holds the whole index list (like every MPI rank reads the whole Euler file), so all ranks derive the
    same partition without talking to each other; only scores travel: 24 bytes per pose.
    """

    def __init__(self, index, L, znum, world, rank, device=None):
        import torch
        self.world, self.rank = world, rank
        self.ranges = shard_z_ranges(index, L, znum, world)
        own = owners_of(index, L, self.ranges)
        self.rows = [np.flatnonzero(own == r) for r in range(world)]
        self.n_total = len(index)
        self.n_max = max(1, max(len(r) for r in self.rows))
        self.my_rows = self.rows[rank]
        # flat destination of every slot of the gathered [world][n_max] block; padding slots go to a dump column
        dest = np.full((world, self.n_max), self.n_total, dtype=np.int64)
        for r in range(world):
            dest[r, :len(self.rows[r])] = self.rows[r]
        self.dest = torch.from_numpy(dest.reshape(-1))
        if device is not None:
            self.dest = self.dest.to(device)
        self.local = torch.zeros((3, self.n_max), dtype=torch.float64, device=device)
        self.gathered = torch.empty((world * 3, self.n_max), dtype=torch.float64, device=device)
        self.table = torch.zeros((3, self.n_total + 1), dtype=torch.float64, device=device)

    def gather(self, group=None):
        """self.local[:, :len(my_rows)] holds this rank's scores -> self.table[:, :n_total] in input order, every rank"""
        import torch.distributed as dist
        if self.world > 1:
            dist.all_gather_into_tensor(self.gathered, self.local, group=group)
            g = self.gathered.view(self.world, 3, self.n_max).permute(1, 0, 2).reshape(3, -1)
        else:
            g = self.local
        self.table.index_copy_(1, self.dest, g)
        return self.table[:, :self.n_total]
