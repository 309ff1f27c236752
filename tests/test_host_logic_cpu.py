"""CPU suite: host-side logic -- tile planner, parameter arena, session shim, and the data-parallel /
tile-sharded math on a world_size-2 gloo group."""
import os
import socket

import numpy as np
import pytest
import torch

from ml_super_resolution_b200.tiling import MAX_PANEL_W, plan_tiles, shard_tiles


@pytest.mark.parametrize("FH,FW,halo,max_h", [(2160, 3840, 20, None), (1080, 1920, 4, None), (41, 41, 20, None), (300, 255, 20, 100), (64, 224, 4, None),
                                             (96, 200, 8, 40), (2160, 3840, 20, 540), (17, 1000, 13, None)])
def test_tiles_cover_frame_once_with_halo(FH, FW, halo, max_h):
    Ht, Wt, tiles = plan_tiles(2, FH, FW, halo, MAX_PANEL_W, max_h)
    assert Wt <= MAX_PANEL_W and (max_h is None or Ht <= max_h or FH <= max_h)
    owner = np.zeros((2, FH, FW), np.int32)
    for t in tiles:
        assert 0 <= t.y0 and t.y0 + Ht <= FH and 0 <= t.x0 and t.x0 + Wt <= FW, "tile leaves the frame"
        owner[t.frame, t.y0 + t.own_y0:t.y0 + t.own_y1, t.x0 + t.own_x0:t.x0 + t.own_x1] += 1
        # owned pixels are >= halo away from every crop edge that is not a frame edge
        if t.x0 > 0:
            assert t.own_x0 >= halo
        if t.x0 + Wt < FW:
            assert Wt - t.own_x1 >= halo
        if t.y0 > 0:
            assert t.own_y0 >= halo
        if t.y0 + Ht < FH:
            assert Ht - t.own_y1 >= halo
    assert np.all(owner == 1), "every output pixel must be owned by exactly one tile"
    # sharding is a partition
    parts = [shard_tiles(tiles, r, 3) for r in range(3)]
    assert sum(len(p) for p in parts) == len(tiles) and [t for p in parts for t in p] == tiles


def test_session_shim_fetch_structures():
    from ml_super_resolution_b200.session import Handle, Session, placeholder

    class G:
        def execute(self, keys, feeds):
            return {k: (k, feeds.get(ph)) for k in keys}

    ph = placeholder([None, 3], "x")
    g = G()
    with Session() as s:
        out = s.run({"a": Handle(g, "a"), "b": [Handle(g, "b"), None]}, feed_dict={ph: 5})
    assert out == {"a": ("a", 5), "b": [("b", 5), None]}


def test_param_arena_roundtrip_cpu():
    from collections import OrderedDict
    from ml_super_resolution_b200.params import ParamArena
    p = OrderedDict([("conv2d/kernel:0", np.arange(54, dtype=np.float32).reshape(3, 3, 3, 2)), ("conv2d/bias:0", np.array([1, 2], np.float32))])
    a = ParamArena(p, device="cpu")
    assert a.size % 4 == 0 and all(o % 4 == 0 for o in a.offsets.values())
    back = a.to_numpy()
    assert all(np.array_equal(back[k], p[k]) for k in p)
    assert a.decay_mask.sum() == 54 and a.view("conv2d/bias:0").tolist() == [1, 2]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _dp_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.distributed.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import models as OM
    torch.set_num_threads(1)
    L = 3
    params = OM.vdsr_init(seed=5, num_layers=L)
    sd = OM.synthetic_images(1, 4, 9, 9, 3)
    hd = OM.synthetic_images(2, 4, 9, 9, 3)
    lo, hi = rank * 2, rank * 2 + 2
    # per-rank loss scaled by the GLOBAL element count (what srk_mse_fwd_bwd's numel_total does), no l2 term
    p = OM._to_t(params, np.float64, requires_grad=True)
    sr = OM.vdsr_forward_t(p, OM._t(sd[lo:hi], np.float64), L)
    loss = ((sr - OM._t(hd[lo:hi], np.float64)) ** 2).sum() / sd.size
    loss.backward()
    flat = torch.cat([p[k].grad.reshape(-1) for k in params])
    torch.distributed.all_reduce(flat)  # the one exchange step of data-parallel training
    ltot = loss.detach().clone()
    torch.distributed.all_reduce(ltot)
    if rank == 0:
        q.put((flat.numpy(), float(ltot)))
    torch.distributed.destroy_process_group()


def test_data_parallel_gradients_equal_single_gpu_gloo():
    import torch.multiprocessing as mp
    from oracle import models as OM
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    flat, ltot = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    params = OM.vdsr_init(seed=5, num_layers=3)
    sd = OM.synthetic_images(1, 4, 9, 9, 3)
    hd = OM.synthetic_images(2, 4, 9, 9, 3)
    _, mse, grads, _ = OM.vdsr_loss_and_grads(params, sd, hd, num_layers=3, weight_decay=0.0)
    ref = np.concatenate([grads[k].reshape(-1) for k in params])
    assert abs(ltot - mse) < 1e-12
    assert np.abs(flat - ref).max() < 1e-12


@pytest.mark.parametrize("k", [3, 5, 9])
def test_srcnn_tap_block_decomposition(k):
    """The k x k weight gradient is computed as 3x3 tap blocks over row-shifted activation views: the block centres must
    tile the kernel exactly once and the assembly must put every tap back in its place (srcnn/srcnn.py:100-130 backward)."""
    from ml_super_resolution_b200.srcnn.srcnn import assemble_tap_blocks, tap_block_centres
    centres = tap_block_centres(k)
    h = k // 2
    seen = {}
    ci, co = 2, 3
    blocks = torch.zeros((len(centres), 9, ci, co))
    for blk, (a, b) in enumerate(centres):
        for u in (-1, 0, 1):
            for v in (-1, 0, 1):
                tap = (a + u, b + v)  # offset of this block tap relative to the kernel centre
                if -h <= tap[0] <= h and -h <= tap[1] <= h:
                    assert tap not in seen, f"tap {tap} covered twice"
                    seen[tap] = blk
                # a block tap's gradient is identified by its (row, column) offset
                blocks[blk, (u + 1) * 3 + (v + 1)] = float((tap[0] + 16) * 100 + (tap[1] + 16))
    assert len(seen) == k * k
    full = assemble_tap_blocks(blocks, k)
    assert full.shape == (k, k, ci, co)
    for i in range(k):
        for j in range(k):
            assert float(full[i, j, 0, 0]) == float((i - h + 16) * 100 + (j - h + 16))


@pytest.mark.parametrize("FH,FW,halo,max_w,max_h,world", [(2160, 3840, 20, 254, None, 1), (2160, 3840, 20, 254, 570, 4), (2160, 3840, 20, 254, 305, 8),
                                                          (96, 200, 8, 80, None, 3), (96, 200, 8, 72, 56, 2), (270, 480, 13, 63, None, 1),
                                                          (48, 56, 13, 40, None, 3), (48, 56, 13, 40, None, 2)])
def test_seam_exchange_plan(FH, FW, halo, max_w, max_h, world):
    """Seam-exchange tiling (srk_fpa_halo_exchange): exchange is only chosen when every rank holds whole row bands; in that mode
    neighbouring panels of a band overlap by at least two columns, each non-owned column is owned by the direct neighbour,
    and the owned rectangles of all ranks still partition the frame exactly once."""
    from ml_super_resolution_b200.tiling import plan_seam_exchange
    cover = np.zeros((FH, FW), np.int32)
    modes = set()
    for rank in range(world):
        Ht, Wt, tiles, exchange, max_cols = plan_seam_exchange(1, FH, FW, halo, max_w, max_h, rank, world)
        modes.add(exchange)
        assert Wt <= max_w
        for i, t in enumerate(tiles):
            assert 0 <= t.x0 and t.x0 + Wt <= FW and 0 <= t.y0 and t.y0 + Ht <= FH
            cover[t.y0 + t.own_y0:t.y0 + t.own_y1, t.x0 + t.own_x0:t.x0 + t.own_x1] += 1
            if exchange:
                assert t.own_x0 + (Wt - t.own_x1) <= max_cols
                if t.own_x0 > 0:  # left seam: the previous tile is the same band's left neighbour and owns those columns
                    p = tiles[i - 1]
                    assert p.y0 == t.y0 and p.x0 + p.own_x0 <= t.x0 and t.x0 + t.own_x0 <= p.x0 + p.own_x1
                if t.own_x1 < Wt:
                    q = tiles[i + 1]
                    assert q.y0 == t.y0 and q.x0 + q.own_x0 <= t.x0 + t.own_x1 and t.x0 + Wt <= q.x0 + q.own_x1
    assert len(modes) == 1, "all ranks must agree on the mode"
    assert (cover == 1).all()
    if (FW, max_w, world) == (200, 80, 3):
        assert modes == {False}  # three panels over three ranks: bands are split, receptive-field halos are used
    if world in (1, 4, 8) and FW == 3840:
        assert modes == {True}


def test_input_pipeline_draws_follow_the_reference_rng_sequence():
    """The device input pipeline's host side must consume numpy's RandomState exactly like vdsr/vdsr/dataset.py:85-110
    (shuffle per epoch; per sample randint(w-S), randint(h-S), choice([0,1]), choice(scales)): cropping / flipping with
    its draws in numpy reproduces the restated generator's hd batch bit for bit, across an epoch boundary."""
    from oracle import ops as O
    sys_path_mod = __import__("importlib").import_module("ml_super_resolution_b200.vdsr.dataset")
    rng = np.random.default_rng(4)
    images = [rng.integers(0, 256, (int(rng.integers(42, 70)), int(rng.integers(42, 90)), 3), dtype=np.uint8) for _ in range(5)]
    images.insert(2, rng.integers(0, 256, (20, 100, 3), dtype=np.uint8))  # skipped: smaller than the crop
    S, B = 41, 4
    ref = O.vdsr_image_batches(images, [2.0, 3.0, 4.0], S, B, np.random.RandomState(77))
    draws = sys_path_mod.draw_samples([im.shape for im in images], [2.0, 3.0, 4.0], S, B, np.random.RandomState(77))
    for _ in range(4):  # 16 samples > 5 usable images: crosses epoch boundaries (re-shuffles)
        _, hd_ref = next(ref)
        crops, scales = next(draws)
        assert crops.shape == (B, 4) and scales.shape == (B,) and set(np.unique(scales)) <= {2.0, 3.0, 4.0}
        for b, (i, y, x, flip) in enumerate(crops):
            patch = images[i][y:y + S, x:x + S, :]
            if flip:
                patch = patch[:, ::-1, :]
            hd = np.divide(patch, 255, dtype=np.float32) * 2.0 - 1.0
            assert np.array_equal(hd.astype(np.float32), hd_ref[b].astype(np.float32))


def test_enet_resample_tables_match_the_pinned_restatement():
    """The product's host-side Pillow coefficient tables (enet/datasets.py) equal the oracle's, which is pinned against Pillow."""
    from oracle import ops as O
    from ml_super_resolution_b200.enet.datasets import resample_tables
    for (i, o, interp) in [(128, 32, "bilinear"), (32, 128, "bicubic"), (91, 40, "bilinear"), (40, 91, "bicubic")]:
        ks, b, k = resample_tables(i, o, interp)
        ks2, b2, k2 = O.pil_resample_coeffs(i, o, interp)
        assert ks == ks2 and np.array_equal(b, b2) and np.array_equal(k, k2)


def test_strip_form_rule_matches_the_kernel_geometry():
    """ops.strip_lanes / ops.conv_form mirror cs_geom / conv_strip_applicable of csrc/conv_strip.cu: 126-pixel strips for wide rows,
    K images side by side (each with its zero column) for narrow ones; the strip form from 80 % lane occupancy, or when the row is
    beyond the flat-stream kernel's 254 pixels."""
    from ml_super_resolution_b200 import ops
    assert ops.strip_lanes(41) == 1.0           # 3 x 42 lanes: VDSR's training patches
    assert abs(ops.strip_lanes(32) - 99 / 126) < 1e-12
    assert abs(ops.strip_lanes(128) - 129 / 252) < 1e-12
    assert abs(ops.strip_lanes(242) - 243 / 252) < 1e-12
    assert ops.strip_lanes(3840) > 0.98
    form = lambda w: ops.conv_form(w).form  # noqa: E731
    assert [form(w) for w in (41, 99, 100, 125, 128, 200, 242, 254, 300, 3840)] == ["flat", "flat", "strip", "strip", "flat", "flat", "strip", "flat", "strip", "strip"]
    assert form("strip") == "strip" and form("flat") == "flat"


def test_rank_grid_minimises_halo_pixels():
    """tiling.rank_grid: the rank grid of tile-sharded inference -- the grid whose largest region costs the fewest
    strip rows; bands for frames that are much taller than wide, one region for one rank; gy * gx is always the world size."""
    from ml_super_resolution_b200.tiling import rank_grid
    assert rank_grid(8, 2160, 3840, 20) == (4, 2)   # eight 245-px panels per region fill their 16 strips; 2 x 4 would waste a fifth
    assert rank_grid(1, 2160, 3840, 20) == (1, 1)
    assert rank_grid(2, 2160, 3840, 20) in ((1, 2), (2, 1))
    assert rank_grid(8, 8000, 100, 20)[0] == 8
    for world in (1, 2, 3, 4, 6, 8):
        gy, gx = rank_grid(world, 270, 3840, 20)
        assert gy * gx == world


@pytest.mark.parametrize("world,H,W", [(1, 2160, 3840), (2, 2160, 3840), (4, 2160, 3840), (8, 2160, 3840), (3, 120, 330), (8, 120, 330), (6, 41, 41)])
def test_rank_regions_partition_the_frame(world, H, W):
    """tiling.rank_region: the owned regions of the ranks tile the frame exactly once; every rank reads its region plus the
    receptive-field halo, clipped at the frame (what bench.py copies host -> device and VdsrNet.forward computes on)."""
    from ml_super_resolution_b200.tiling import rank_region
    halo = 20
    cover = np.zeros((H, W), np.int32)
    for r in range(world):
        (y0, y1, x0, x1), (ya, yb, xa, xb) = rank_region(world, r, H, W, halo)
        cover[y0:y1, x0:x1] += 1
        assert ya == max(0, y0 - halo) and yb == min(H, y1 + halo) and xa == max(0, x0 - halo) and xb == min(W, x1 + halo)
    assert cover.min() == 1 and cover.max() == 1
