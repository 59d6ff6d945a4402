"""Partitioning of the recon path over one 8xB200 box; pure host logic + thin
torch.distributed plumbing, so it is testable with gloo on CPU.

Two ways to shard (neither exists in the reference, which asserts B == 1, g2vlm.py:1006):

* scenes: independent units -> rank r takes scenes r, r+R, ...; no data-path collective.
* views of ONE long scene ("sequence parallel"): contiguous view ranges per rank.
    - DINO is sharded by *attention segment*, not by image: the reference applies cu_seqlens built from
      patch counts P to rows laid out with S = P+5 rows per image (quirk Q1), so segment i = flattened
      rows [iP, (i+1)P) straddles images. Rank r owns the rows of its views' segments [v0*P, v1*P) (the
      last rank also the 5N uncovered tail rows); every non-attention DINO op is row-local. Afterwards the
      rows of image v live at [v*S, (v+1)*S): rank r needs rows up to v1*S = v1*P + 5*v1, i.e. the first
      5*v1 rows of rank r+1 -> ONE neighbour exchange.
    - MoT: every op is row-local except the shared attention, which needs all keys: per layer one
      all-gather of the rank's K/V rows (GQA: K+V is 1/3 the width of Q); the K0 prefix rows are
      replicated.
    - Pi3 heads: per-view local; the global-points decoder's context is view 0's hidden -> one broadcast
      from rank 0.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Tuple


def scenes_for_rank(n_scenes: int, rank: int, world: int) -> List[int]:
    """Round-robin scene assignment (data parallel over independent scenes)."""
    return list(range(rank, n_scenes, world))


@dataclass(frozen=True)
class ViewShard:
    rank: int
    world: int
    n_views: int          # views of the whole scene
    P: int                # patches per view
    n_reg: int            # cls + register rows per image (5)
    v0: int               # first view owned
    v1: int               # one past the last view owned

    @property
    def S(self) -> int:
        return self.P + self.n_reg

    @property
    def n_local(self) -> int:
        return self.v1 - self.v0

    # ---- DINO rows (flattened [N*S] sequence) owned by this rank -------------------------------
    @property
    def dino_rows(self) -> Tuple[int, int]:
        """Rows of the segments [v0, v1): [v0*P, v1*P); the last rank also owns the uncovered tail
        [N*P, N*S) (rows in no segment, attention output defined as zero)."""
        end = self.v1 * self.P
        if self.rank == self.world - 1:
            end = self.n_views * self.S
        return self.v0 * self.P, end

    @property
    def dino_covered_rows(self) -> Tuple[int, int]:
        """Rows covered by this rank's attention segments."""
        return self.v0 * self.P, self.v1 * self.P

    @property
    def dino_images(self) -> Tuple[int, int]:
        """Images whose rows intersect `dino_rows` (their patch embeddings are needed locally)."""
        a, b = self.dino_rows
        return a // self.S, (b - 1) // self.S + 1

    # ---- re-association of rows with images after the encoder --------------------------------
    @property
    def token_rows_needed(self) -> Tuple[int, int]:
        """Global DINO rows holding the tokens of the owned views: [v0*S, v1*S)."""
        return self.v0 * self.S, self.v1 * self.S

    @property
    def recv_from_next(self) -> Tuple[int, int]:
        """Global rows this rank must receive from rank+1: (own end, v1*S) — empty for the last rank."""
        own_end = self.dino_rows[1]
        need_end = self.token_rows_needed[1]
        return (own_end, need_end) if need_end > own_end else (own_end, own_end)

    @property
    def send_to_prev(self) -> Tuple[int, int]:
        """Global rows this rank must send to rank-1: its first 5*v0 rows [v0*P, v0*S)."""
        if self.rank == 0:
            return 0, 0
        return self.v0 * self.P, self.v0 * self.S

    def token_row_index(self) -> List[int]:
        """For every token of the owned views (view-major, patch order) the global DINO row holding it."""
        out = []
        for v in range(self.v0, self.v1):
            out.extend(range(v * self.S + self.n_reg, (v + 1) * self.S))
        return out

    # ---- MoT rows (packed order: P+2 rows per view) -------------------------------------------
    @property
    def packed_rows(self) -> Tuple[int, int]:
        return self.v0 * (self.P + 2), self.v1 * (self.P + 2)


def view_ranges(n_views: int, world: int) -> List[Tuple[int, int]]:
    """Contiguous view range of every rank; the first n_views % world ranks own one view more."""
    if n_views < world:
        raise ValueError(f"view-sharding needs at least one view per rank ({n_views} views, {world} ranks)")
    base, rem = divmod(n_views, world)
    out, v = [], 0
    for r in range(world):
        n = base + (1 if r < rem else 0)
        out.append((v, v + n))
        v += n
    return out


def shard_views(n_views: int, P: int, rank: int, world: int, n_reg: int = 5) -> ViewShard:
    ranges = view_ranges(n_views, world)
    v0, v1 = ranges[rank]
    sh = ViewShard(rank=rank, world=world, n_views=n_views, P=P, n_reg=n_reg, v0=v0, v1=v1)
    # the neighbour exchange only reaches rank+1 if the shifted rows fit inside ITS block
    if rank < world - 1 and n_reg * sh.v1 > (ranges[rank + 1][1] - ranges[rank + 1][0]) * P:
        raise ValueError("too many ranks for this scene: DINO row shift exceeds one rank's block")
    return sh
