"""Tile planner for full-frame inference.

A frame larger than the flat-stream conv kernel's widest panel (254 px), larger than L2, or
spread over several GPUs is cut into equally sized tiles that all lie INSIDE the frame and
overlap by at least 2*halo, where halo = the network's receptive-field radius (VDSR 20,
ESPCN 4 LR px, ENet 13 LR px).  Each tile runs through the network with ordinary per-layer SAME
padding inside the crop; every output pixel is taken from the one tile that "owns" it, i.e. a
tile in which the pixel sits >= halo away from any crop edge that is not a true frame edge --
there the tiled result equals the full-frame result exactly (crop-edge error advances one pixel
per 3x3 layer and never reaches an owned pixel).  Pure host-side integer logic; numpy only.
"""
from __future__ import annotations

from dataclasses import dataclass

MAX_PANEL_W = 254  # widest image the flat-stream kernels serve: a tile's 2*Wp + 128 resident rows must leave look-ahead room in the smem ring


@dataclass(frozen=True)
class Tile:
    frame: int
    y0: int
    x0: int
    own_y0: int
    own_y1: int
    own_x0: int
    own_x1: int

    def as_tuple(self):
        return (self.frame, self.y0, self.x0, self.own_y0, self.own_y1, self.own_x0, self.own_x1)


def _split_axis(size: int, halo: int, max_len: int | None):
    """-> (tile_len, [(start, own_lo, own_hi)]) with own ranges in tile-local coordinates."""
    if max_len is None or size <= max_len:
        return size, [(0, 0, size)]
    if max_len <= 2 * halo:
        raise ValueError(f"max tile length {max_len} does not exceed twice the halo {halo}")
    core = max_len - 2 * halo
    parts = -(-(size - 2 * halo) // core)  # ceil
    length = -(-(size - 2 * halo) // parts) + 2 * halo
    while True:  # integer rounding of the starts can cost one pixel of overlap
        starts = [round(j * (size - length) / (parts - 1)) for j in range(parts)]
        if all(starts[j] + length - starts[j + 1] >= 2 * halo for j in range(parts - 1)):
            break
        length += 1
    assert length <= max_len
    # ownership boundaries: midpoint of each overlap
    bounds = [0]
    for j in range(parts - 1):
        lo, hi = starts[j + 1], starts[j] + length  # overlap [lo, hi)
        assert hi - lo >= 2 * halo, "internal: overlap smaller than 2*halo"
        bounds.append((lo + hi) // 2)
    bounds.append(size)
    out = []
    for j in range(parts):
        out.append((starts[j], bounds[j] - starts[j], bounds[j + 1] - starts[j]))
    return length, out


def plan_tiles(n_frames: int, FH: int, FW: int, halo: int, max_w: int | None = MAX_PANEL_W, max_h: int | None = None,
               halo_x: int | None = None):
    """Returns (Ht, Wt, [Tile...]): every tile is Ht x Wt, inside the frame, owned rects partition it.
    `halo_x` (default: `halo`) is the column overlap radius: 1 when the panels of a band exchange their seam columns after
    every layer (`srk_fpa_halo_exchange`) instead of recomputing a receptive-field halo."""
    Wt, xs = _split_axis(FW, halo if halo_x is None else halo_x, max_w)
    Ht, ys = _split_axis(FH, halo, max_h)
    tiles = []
    for f in range(n_frames):
        for (y0, oy0, oy1) in ys:
            for (x0, ox0, ox1) in xs:
                tiles.append(Tile(f, y0, x0, oy0, oy1, ox0, ox1))
    return Ht, Wt, tiles


def shard_tiles(tiles, rank: int, world: int):
    """Contiguous, balanced shard of the tile list for one rank (no data-path collective)."""
    n = len(tiles)
    lo = (rank * n) // world
    hi = ((rank + 1) * n) // world
    return tiles[lo:hi]


def plan_seam_exchange(n_frames: int, FH: int, FW: int, halo: int, max_w: int | None, max_h: int | None, rank: int = 0, world: int = 1,
                       group: int | None = None):
    """Tile plan for the seam-exchange mode of tiled inference (`srk_fpa_halo_exchange`): column panels overlap by one column
    per side and swap their seam columns after every layer, row bands keep the receptive-field `halo`.
    Returns (Ht, Wt, tiles of this rank, exchange, max_cols).  `exchange` is False -- and the plan falls back to
    receptive-field halos in both directions -- unless every row band of panels lies completely inside this rank's shard and
    inside one launch group of `group` tiles (a panel must find both neighbours in the same FPA batch).
    `max_cols` bounds the number of non-owned columns of any panel (argument of the exchange kernel)."""
    Ht, Wt, tiles = plan_tiles(n_frames, FH, FW, halo, max_w, max_h, halo_x=1)
    per_band = len({t.x0 for t in tiles})
    mine = shard_tiles(tiles, rank, world)
    # the mode is a property of the whole job: every rank must reach the same decision, so every shard boundary is checked
    bounds = [(r * len(tiles)) // world for r in range(world + 1)]
    largest = max(b1 - b0 for b0, b1 in zip(bounds, bounds[1:]))
    exchange = per_band > 1 and all(b % per_band == 0 for b in bounds) and (group is None or group >= largest)
    if not exchange:
        Ht, Wt, tiles = plan_tiles(n_frames, FH, FW, halo, max_w, max_h)
        mine = shard_tiles(tiles, rank, world)
    max_cols = 2 * max((max(t.own_x0, Wt - t.own_x1) for t in mine), default=0) if exchange else 0
    return Ht, Wt, mine, exchange, max_cols


def rank_grid(world: int, H: int, W: int, halo: int, strip_w: int = 126, max_panel_w: int = 251) -> tuple[int, int]:
    """(gy, gx), gy * gx == world: the grid of frame regions that costs a rank the least work.  Every region is extended by `halo`
    on its interior sides (a 2-D grid recomputes fewer halo pixels than row bands), and inside a rank the region is cut
    into equal column panels of at most `max_panel_w` pixels whose rows occupy whole 126-pixel strips of the column-strip kernel:
    the cost of a grid is rows x strips of its largest region, so 4 x 2 beats 2 x 4 for a 4K frame on 8 GPUs (eight 245-px panels
    fill 16 strips, five 202-px panels of a 1000-px region waste a fifth of theirs)."""
    best = None
    for gy in range(1, world + 1):
        if world % gy:
            continue
        gx = world // gy
        rows = -(-H // gy) + (2 * halo if gy > 2 else (halo if gy == 2 else 0))
        width = -(-W // gx) + (2 * halo if gx > 2 else (halo if gx == 2 else 0))
        k = 1
        while -(-width // k) + 2 > max_panel_w:
            k += 1
        panel = -(-width // k) + (2 if k > 1 else 0)
        strips = k * (-(-(panel + 1) // strip_w))
        cost = rows * strips
        if best is None or cost < best[0]:
            best = (cost, gy, gx)
    return best[1], best[2]


def rank_region(world: int, rank: int, H: int, W: int, halo: int):
    """The frame region rank `rank` of `world` owns in tile-sharded inference, (y0, y1, x0, x1), and the region it reads (the owned
    one extended by `halo` on its interior sides, clipped at the frame), (ya, yb, xa, xb)."""
    gy, gx = rank_grid(world, H, W, halo)
    ry, rx = divmod(rank, gx)
    y0, y1, x0, x1 = H * ry // gy, H * (ry + 1) // gy, W * rx // gx, W * (rx + 1) // gx
    return (y0, y1, x0, x1), (max(0, y0 - halo), min(H, y1 + halo), max(0, x0 - halo), min(W, x1 + halo))
