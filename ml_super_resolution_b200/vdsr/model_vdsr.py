"""VDSR on the B200 conv hot path -- drop-in for vdsr/vdsr/model_vdsr.py of the reference.

`build_model(sd_images, hd_images=None, num_layers=20, use_adam=False)` keeps the reference's
signature and dict keys (`conv.{i}`, `relu.{i}`, `sd_images`, `sr_images`, `step`, `loss`, `trainer`,
`hd_images`, `learning_rate`; reference :6,72,76,101,108-109,186-190); the values are graph handles
evaluated by `session.Session.run(fetches, feed_dict)` the way `tf.Session.run` evaluates tensors.
Underneath, `VdsrNet` drives the libsrk kernels:

  layer 1      srk_conv_first_tc(3x3, C->64, ReLU)                       reference :62-70 (i = 0)
  layers 2..19 srk_conv_tc      (tcgen05 shift-GEMM, 64->64, ReLU)        reference :62-70
  layer 20     srk_conv_tc_last (64->C, fused `sd_images + residual`)     reference :85-104
  loss         srk_mse_fwd_bwd + srk_sumsq_masked (MSE mean + 1e-4*l2)    reference :120-125
  backward     srk_conv_tc(dgrad) / srk_conv_wgrad_tc / first+last wgrad  reference :146-148 (autodiff)
  optimiser    srk_adam_step(_dev) | srk_momentum_clip_step               reference :145-184
"""
from __future__ import annotations

import math
from collections import OrderedDict

import numpy as np
import torch

from .. import ops
from ..initializers import tf_conv_name, vdsr_params
from ..params import ParamArena
from ..session import Handle, Placeholder
from ..tiling import MAX_PANEL_W, plan_seam_exchange, plan_tiles, rank_region

WEIGHT_DECAY = 1e-4  # tf.contrib.layers.l2_regularizer(0.0001), reference :34


class VdsrNet:
    def __init__(self, params: "OrderedDict[str, np.ndarray] | None" = None, num_layers=20, channels=3, device="cuda", seed=0):
        assert num_layers >= 3
        self.L = num_layers
        self.C = channels
        self.device = device
        if params is None:
            params = vdsr_params(seed, num_layers, channels)
        order = OrderedDict()
        for i in range(num_layers):
            for suffix in ("kernel:0", "bias:0"):
                k = f"{tf_conv_name(i)}/{suffix}"
                order[k] = np.asarray(params[k], np.float32)
        self.arena = ParamArena(order, device)
        self.step = 0
        self.learning_rate = 0.1  # `learning_rate` variable initial value, reference :136-141
        self._build_pack_plan()
        self.repack()
        self._train_bufs = None
        self._infer_bufs = {}

    # ------------------------------------------------------------------ weights
    def _kname(self, i):
        return f"{tf_conv_name(i)}/kernel:0"

    def _bname(self, i):
        return f"{tf_conv_name(i)}/bias:0"

    def _build_pack_plan(self):
        a, L = self.arena, self.L
        plan = ops.PackPlan(self.device)
        self._fwd_idx, self._dg_idx = {}, {}
        self._fwd_idx[0] = plan.add(a.offsets[self._kname(0)], 3, self.C, 64, ops.PACK_FIRST)
        for i in range(1, L - 1):
            self._fwd_idx[i] = plan.add(a.offsets[self._kname(i)], 3, 64, 64, ops.PACK_FWD, 64, 64)
            self._dg_idx[i] = plan.add(a.offsets[self._kname(i)], 3, 64, 64, ops.PACK_DGRAD, 64, 64)
        self._fwd_idx[L - 1] = plan.add(a.offsets[self._kname(L - 1)], 3, 64, self.C, ops.PACK_FWD, 16, 64)
        self._dg_idx[L - 1] = plan.add(a.offsets[self._kname(L - 1)], 3, 64, self.C, ops.PACK_FIRST_ROT180T)
        plan.finalize()
        self.plan = plan
        self.bias_last = torch.zeros(16, dtype=torch.float32, device=self.device)

    def repack(self):
        """fp32 HWIO arena -> GEMM-ready bf16 blocks for every layer: one launch."""
        self.plan.run(self.arena.w)
        self.bias_last[: self.C].copy_(self.arena.view(self._bname(self.L - 1)))

    def wf(self, i):
        return self.plan.views[self._fwd_idx[i]]

    def wd(self, i):
        return self.plan.views[self._dg_idx[i]]

    def load_params(self, params: dict):
        self.arena.load_numpy(params)
        self.repack()

    # ------------------------------------------------------------------ inference
    def forward(self, sd: torch.Tensor, taps: dict | None = None, out: torch.Tensor | None = None, tile_rows: int | None = None,
                rank: int = 0, world: int = 1, max_panel_w: int = MAX_PANEL_W, _form_width: int | None = None) -> torch.Tensor:
        """sd fp32 [N,H,W,C] on device -> sr.  Frames wider than 254 px (or taller than `tile_rows`) are
        cut into halo-overlapped tiles; with world > 1 this rank computes only its shard of the tiles
        (tile-sharded multi-GPU inference, no collective; pixels it does not own are left untouched)."""
        n, H, W, C = sd.shape
        assert C == self.C
        a = self.arena
        assert max_panel_w <= MAX_PANEL_W
        form_width = W if _form_width is None else _form_width
        if world > 1:
            # Tile sharding over ranks: a gy x gx grid of frame regions, one per rank, each extended by the receptive-field halo
            # (num_layers pixels) on its interior sides and run through the single-rank path below; the rank writes only the
            # pixels it owns.  A 2 x 4 grid of a 4K frame recomputes 5 % of the pixels where 8 row bands recompute 15 %.  Every
            # owned pixel sees its whole receptive field (or the frame border's zero padding) and the FRAME's kernel form, so the
            # union of the ranks' pixels equals the un-sharded frame bit for bit.
            (y0, y1, x0, x1), (ya, yb, xa, xb) = rank_region(world, rank, H, W, self.L)
            if y1 <= y0 or x1 <= x0:
                return out if out is not None else torch.empty_like(sd)
            sub = self.forward(sd[:, ya:yb, xa:xb].contiguous(), tile_rows=tile_rows, max_panel_w=max_panel_w, _form_width=form_width)
            if out is None:
                out = torch.empty_like(sd)
            out[:, y0:y1, x0:x1].copy_(sub[:, y0 - ya:y1 - ya, x0 - xa:x1 - xa])
            return out
        need_tiles = W > max_panel_w or (tile_rows is not None and H > tile_rows)
        if need_tiles and ops.conv_form(form_width).form == "strip":
            max_panel_w = min(max_panel_w, 2 * ops.STRIP_W - 1)  # a panel plus its zero column fills at most two 126-pixel strips
        if out is None:
            out = torch.empty_like(sd)
        if not need_tiles:
            bufs = self._get_infer_bufs(n, H, W)
            t = ops.conv_first_tc(sd, self.wf(0), a.view(self._bname(0)), 3, "SAME", "relu", out=bufs[0])
            if taps is not None:
                taps["conv.1"] = taps["relu.1"] = ops.fpa_to_nhwc(t)
            for i in range(1, self.L - 1):
                with ops.conv_form(form_width):
                    t = ops.conv_tc(t, self.wf(i), a.view(self._bname(i)), 3, "relu", out=bufs[i % 2])
                if taps is not None:
                    taps[f"conv.{i + 1}"] = taps[f"relu.{i + 1}"] = ops.fpa_to_nhwc(t)
            ops.conv_tc_last(t, self.wf(self.L - 1), self.bias_last, 3, self.C, None, addend=sd, out=out)
            if taps is not None:
                taps[f"conv.{self.L}"] = out - sd
            return out
        assert taps is None, "feature-map taps are only available for un-tiled frames"
        # Column seams: panels of one band overlap by one column per side and swap their seam columns after every layer
        # (srk_fpa_halo_exchange) instead of recomputing a 20-px halo (16 % more pixels at 252-px panels).  That needs all
        # panels of a band in this rank's shard and in one launch group; otherwise fall back to the receptive-field halo.
        Ht1, Wt1, _ = plan_tiles(n, H, W, halo=self.L, max_w=max_panel_w, max_h=tile_rows, halo_x=1)
        Ht, Wt, tiles, exchange, max_cols = plan_seam_exchange(n, H, W, self.L, max_panel_w, tile_rows, rank, world,
                                                               group=self._tile_group_size(Ht1, Wt1))
        group = self._tile_group_size(Ht, Wt)
        if not tiles:
            return out
        for g0 in range(0, len(tiles), group):
            chunk = tiles[g0:g0 + group]
            key = (tuple(t.as_tuple() for t in chunk), str(sd.device))
            panels = self._panel_cache(key)
            bufs = self._get_infer_bufs(len(chunk), Ht, Wt)
            t = ops.conv_first_tc(sd, self.wf(0), a.view(self._bname(0)), 3, "SAME", "relu", panels=panels, panel_hw=(Ht, Wt),
                                  out=bufs[0])
            if exchange:  # the gather pads a panel's window edge with zeros: its seam columns come from the neighbour too
                ops.fpa_halo_exchange(t, panels, max_cols)
            for i in range(1, self.L - 1):
                with ops.conv_form(form_width):  # the FRAME's width decides the kernel form: panels compute what the un-tiled frame would
                    t = ops.conv_tc(t, self.wf(i), a.view(self._bname(i)), 3, "relu", out=bufs[i % 2])
                if exchange:
                    ops.fpa_halo_exchange(t, panels, max_cols)
            ops.conv_tc_last(t, self.wf(self.L - 1), self.bias_last, 3, self.C, None, addend=sd, panels=panels, frame_shape=(n, H, W),
                             out=out)
        return out

    tile_group_bytes = 1 << 62  # set to e.g. 48 MB to keep a tile group's activations L2-resident

    def _tile_group_size(self, Ht, Wt):
        per_tile = ops.fpa_rows(1, Ht, Wt) * 128
        return max(1, int(self.tile_group_bytes // per_tile))

    def _panel_cache(self, key):
        cache = self.__dict__.setdefault("_panels", {})
        if key not in cache:
            cache[key] = ops.make_panels(list(key[0]), self.device)
        return cache[key]

    def _get_infer_bufs(self, n, H, W):
        key = (n, H, W)
        if key not in self._infer_bufs:
            if len(self._infer_bufs) > 8:
                self._infer_bufs.clear()
            self._infer_bufs[key] = [ops.fpa_empty(n, H, W, 64, self.device) for _ in range(2)]
        return self._infer_bufs[key]

    # ------------------------------------------------------------------ training
    def _get_train_bufs(self, n, H, W):
        key = (n, H, W)
        if self._train_bufs is None or self._train_bufs["key"] != key:
            self.arena.enable_training()
            self._train_bufs = {
                "key": key,
                "acts": [ops.fpa_empty(n, H, W, 64, self.device) for _ in range(self.L - 1)],
                # every layer's dY is kept so that all weight gradients run in one batched launch after the dgrad chain
                "dy": [ops.fpa_empty(n, H, W, 64, self.device) for _ in range(self.L - 1)],
                "sr": torch.empty((n, H, W, self.C), dtype=torch.float32, device=self.device),
                "dsr": torch.empty((n, H, W, self.C), dtype=torch.float32, device=self.device),
                "loss": torch.zeros(2, dtype=torch.float32, device=self.device),  # [mse, l2 regulariser]
                "wg_stride": (ops.wgrad_workspace_bytes(n, H, W) + 1023) // 1024 * 1024,
                "lr_t": torch.zeros(1, dtype=torch.float32, device=self.device),
            }
            b = self._train_bufs
            # every layer's weight gradient runs on the tensor-core wgrad kernel (first / last layer over operands
            # zero-padded to 64 channels); one workspace slice per layer so a single launch folds all partial sums
            b["wg_ws"] = torch.empty(b["wg_stride"] * self.L, dtype=torch.uint8, device=self.device)
            b["sd_fpa"] = ops.fpa_empty(n, H, W, 64, self.device)
            b["dsr_fpa"] = ops.fpa_empty(n, H, W, 64, self.device)
            a = self.arena
            ents = []
            for i in range(self.L):
                ci_n = self.C if i == 0 else 64
                co_n = self.C if i == self.L - 1 else 64
                ents.append((a.view(self._kname(i), "g"), a.view(self._bname(i), "g"), ci_n, co_n))
            b["wg_dsts"] = ops.make_wgrad_dsts(ents, self.device)
        return self._train_bufs

    def forward_backward(self, sd: torch.Tensor, hd: torch.Tensor, numel_total: float | None = None):
        """Forward, MSE+L2 loss and all parameter gradients (into arena.g, overwritten).  Returns the
        buffer dict (`loss` = [mse, reg], `sr`).  `numel_total` = GLOBAL element count under data
        parallelism so that summing rank gradients reproduces the single-GPU MEAN reduction."""
        n, H, W, C = sd.shape
        assert W <= MAX_PANEL_W, "training patches wider than 254 px are not supported by the flat-stream kernels"
        a, L = self.arena, self.L
        b = self._get_train_bufs(n, H, W)
        acts, dyb = b["acts"], b["dy"]
        # ---- forward, keeping every activation for the backward pass
        ops.conv_first_tc(sd, self.wf(0), a.view(self._bname(0)), 3, "SAME", "relu", out=acts[0])
        for i in range(1, L - 1):
            ops.conv_tc(acts[i - 1], self.wf(i), a.view(self._bname(i)), 3, "relu", out=acts[i])
        ops.conv_tc_last(acts[L - 2], self.wf(L - 1), self.bias_last, 3, C, None, addend=sd, out=b["sr"])
        # ---- loss + d(loss)/d(sr)
        b["loss"].zero_()
        ops.mse_fwd_bwd(b["sr"], hd, b["loss"][0:1], b["dsr"], numel_total)
        ops.sumsq_masked(a.w, a.decay_mask, 0.5 * WEIGHT_DECAY, b["loss"][1:2])
        # ---- backward: the dgrad chain stores every layer's dY; then ONE batched tensor-core wgrad launch over all layers
        # (20 separate launches of ~880 chunks each paid 20 fills and drains) and ONE fixed-order reduce launch
        ops.nhwc_to_fpa_pad(b["dsr"], 64, out=b["dsr_fpa"])
        dys = [None] * L
        dys[L - 1] = b["dsr_fpa"]
        dys[L - 2] = ops.conv_first_tc(b["dsr"], self.wd(L - 1), None, 3, "SAME", None, out=dyb[L - 2], mask_src=acts[L - 2], mask_kind="relu")
        for i in range(L - 2, 0, -1):
            dys[i - 1] = ops.conv_tc(dys[i], self.wd(i), None, 3, None, out=dyb[i - 1], mask_src=acts[i - 1], mask_kind="relu")
        ops.nhwc_to_fpa_pad(sd, 64, out=b["sd_fpa"])
        xs = [b["sd_fpa"]] + [acts[i - 1] for i in range(1, L)]
        ops.conv_wgrad_tc_batched(xs, dys, b["wg_ws"], b["wg_stride"], b["wg_dsts"])
        return b

    def apply_gradients(self, lr: float, use_adam=True, lr_t_dev: torch.Tensor | None = None):
        a = self.arena
        self.step += 1
        if use_adam:
            if lr_t_dev is not None:
                ops.adam_step_dev(a.w, a.g, a.m, a.v, lr_t_dev, weight_decay=WEIGHT_DECAY, decay_mask=a.decay_mask)
            else:
                ops.adam_step(a.w, a.g, a.m, a.v, lr, self.step, weight_decay=WEIGHT_DECAY, decay_mask=a.decay_mask)
        else:
            ops.momentum_clip_step(a.w, a.g, a.m, lr, 0.9, 0.01, WEIGHT_DECAY, a.decay_mask)
        self.repack()

    def train_step(self, sd, hd, lr: float | None = None, use_adam=True, group=None):
        """One optimiser step (reference `session.run(trainer)`): returns the pre-update loss (device scalar)."""
        lr = self.learning_rate if lr is None else lr
        world = 1
        if group is not None or (torch.distributed.is_available() and torch.distributed.is_initialized()):
            world = torch.distributed.get_world_size(group)
        b = self.forward_backward(sd, hd, numel_total=float(sd.numel()) * world)
        if world > 1:
            ops.comm_init(group)
            ops.allreduce_grads(self.arena.g)  # srk_allreduce_grads: NCCL sum over NVLink, the path's one exchange step
        self.apply_gradients(lr, use_adam)
        return b["loss"].sum()

    def make_graphed_step(self, sd_static: torch.Tensor, hd_static: torch.Tensor, group=None, peer_exchange: bool = True):
        """Capture the Adam training step into CUDA graphs (the ~85 launches of a step are latency-bound at
        64 patches of 41x41).  Returns `step(lr) -> loss buffer [mse, reg]`; new batches are copied INTO
        `sd_static` / `hd_static` before each call.  ONE graph holds forward + loss + backward, the data-parallel exchange
        step and Adam (learning rate read from device memory, so the same graph serves every step) + re-pack.  The exchange is
        fused with Adam over NVLink peer memory (srk_allreduce_adam_step_dev: every rank sums all ranks' gradients through the peer
        mappings and updates its replica in one kernel; `peer_exchange=False` keeps the NCCL all-reduce of srk_allreduce_grads)."""
        world = 1
        if group is not None or (torch.distributed.is_available() and torch.distributed.is_initialized()):
            world = torch.distributed.get_world_size(group)
        numel_total = float(sd_static.numel()) * world
        a = self.arena
        # the exchange step: fused with Adam over NVLink peer memory (srk_allreduce_adam_step_dev: one kernel, no NCCL call) where
        # the ranks are peers of one node; NCCL all-reduce + Adam otherwise
        fused = world > 1 and peer_exchange and ops.peer_init(a.w.numel(), group)
        if world > 1 and not fused:
            ops.comm_init(group)

        def exchange_and_update(lr_t):
            if fused:
                ops.allreduce_adam_step_dev(a.w, a.g, a.m, a.v, lr_t, weight_decay=WEIGHT_DECAY, decay_mask=a.decay_mask)
            else:
                if world > 1:
                    ops.allreduce_grads(a.g)
                ops.adam_step_dev(a.w, a.g, a.m, a.v, lr_t, weight_decay=WEIGHT_DECAY, decay_mask=a.decay_mask)

        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):  # warm-up outside capture: allocates buffers, sets kernel attributes, NCCL sets up its channels
            self.forward_backward(sd_static, hd_static, numel_total)
            b = self._train_bufs
            exchange_and_update(b["lr_t"])  # lr_t == 0: no-op update
            self.repack()
        torch.cuda.current_stream().wait_stream(side)
        a.m.zero_()
        a.v.zero_()
        torch.cuda.synchronize()
        g_step = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g_step):
            self.forward_backward(sd_static, hd_static, numel_total)
            exchange_and_update(b["lr_t"])
            self.repack()
        lr_feed = ops.PinnedScalarFeed()

        def step(lr: float):
            self.step += 1
            lr_feed.push(self.adam_lr_t(lr, self.step), b["lr_t"])
            g_step.replay()
            return b["loss"]

        step.graphs = (g_step,)
        step.fused_exchange = bool(fused)
        return step

    @staticmethod
    def adam_lr_t(lr: float, t: int, beta1=0.9, beta2=0.999) -> float:
        return lr * math.sqrt(1.0 - beta2 ** t) / (1.0 - beta1 ** t)


# ---------------------------------------------------------------------------------------------
# reference-shaped builder
# ---------------------------------------------------------------------------------------------


def build_model(sd_images, hd_images=None, num_layers=20, use_adam=False, params=None, channels=3, device="cuda", seed=0):
    """Same positional signature and dict keys as vdsr/vdsr/model_vdsr.py:6 `build_model`.

    `sd_images` / `hd_images` are `session.Placeholder`s (or None); the returned dict maps the
    reference's keys to handles that `session.Session.run` evaluates.  Extra keyword arguments
    (`params`, `channels`, `device`, `seed`) choose the initial weights; they default to the
    reference's Xavier initialisation."""
    net = VdsrNet(params, num_layers, channels, device, seed)
    graph = _VdsrGraph(net, sd_images, hd_images, use_adam)
    model = {}
    for i in range(num_layers):
        model[f"conv.{i + 1}"] = Handle(graph, f"conv.{i + 1}")
        if i < num_layers - 1:
            model[f"relu.{i + 1}"] = Handle(graph, f"relu.{i + 1}")
    model["sd_images"] = sd_images
    model["sr_images"] = Handle(graph, "sr_images")
    if hd_images is None:
        return model
    model["step"] = Handle(graph, "step")
    model["loss"] = Handle(graph, "loss")
    model["trainer"] = Handle(graph, "trainer")
    model["hd_images"] = hd_images
    model["learning_rate"] = Placeholder("learning_rate", [], variable_of=graph)
    model["psnr"] = Handle(graph, "psnr")
    return model


class _VdsrGraph:
    """Evaluates fetches for one VDSR model; the `tf.Session.run` counterpart lives in session.py."""

    def __init__(self, net: VdsrNet, sd_ph, hd_ph, use_adam):
        self.net, self.sd_ph, self.hd_ph, self.use_adam = net, sd_ph, hd_ph, use_adam

    def set_variable(self, name, value):
        if name == "learning_rate":
            self.net.learning_rate = float(value)

    def execute(self, keys, feeds):
        net = self.net
        out = {}
        if keys == {"step"}:
            return {"step": net.step}
        learning_rate = net.learning_rate  # a fed value overrides the variable for THIS run only, as a TF feed does
        for ph, val in feeds.items():
            if isinstance(ph, Placeholder) and ph.variable_of is self and ph.name == "learning_rate":
                learning_rate = float(val)
        sd = _to_device(feeds[self.sd_ph], net.device)
        train = "trainer" in keys
        if train or "loss" in keys or "psnr" in keys:
            hd = _to_device(feeds[self.hd_ph], net.device)
        if train:
            out["step"] = net.step
            loss = net.train_step(sd, hd, learning_rate, self.use_adam)
            out["trainer"] = None
            out["loss"] = float(loss)
            sr = net._train_bufs["sr"]
        else:
            taps = {} if any(k.startswith(("conv.", "relu.")) for k in keys) else None
            sr = net.forward(sd, taps)
            if taps:
                for k in keys:
                    if k in taps:
                        out[k] = taps[k].cpu().numpy()
            if "loss" in keys:
                acc = torch.zeros(2, device=net.device)
                ops.mse_fwd_bwd(sr, hd, acc[0:1], None)
                ops.sumsq_masked(net.arena.w, net.arena.decay_mask, 0.5 * WEIGHT_DECAY, acc[1:2])
                out["loss"] = float(acc.sum())
            out.setdefault("step", net.step)
        if "sr_images" in keys:
            out["sr_images"] = sr.cpu().numpy()
        if "psnr" in keys:
            # tf.image.psnr(sr, hd, max_val=2.0), vdsr/vdsr/experiment_train.py:80
            srh, hdh = sr.cpu().numpy().astype(np.float64), np.asarray(feeds[self.hd_ph], np.float64)
            mse = ((srh - hdh) ** 2).reshape(srh.shape[0], -1).mean(axis=1)
            out["psnr"] = 20.0 * np.log10(2.0) - 10.0 * np.log10(mse)
        return out


def _to_device(x, device):
    if isinstance(x, torch.Tensor):
        return x.to(device=device, dtype=torch.float32).contiguous()
    return torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)).to(device)
