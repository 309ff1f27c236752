"""ESPCN on the B200 conv hot path -- drop-in for espcn/espcn/model_espcn.py of the reference.

  f1  srk_conv_first_tc 5x5 C->64 tanh                          reference :30-38 / :117-120
  f2  srk_conv_tc      3x3 64->32 tanh (tcgen05, N=32 tiles)      reference :40-48 / :123-126
  f3  srk_conv_tc_last 3x3 32->C*r^2 linear, fused depth_to_space reference :54-62 / :132-134 and the
                       host un-pack of espcn/espcn/experiment_test.py:173-177

`build_model` / `build_test_model` / `extract_weights` keep the reference's names, arguments and dict
keys; `sr_result(s)` stays in PACKED (un-shuffled) space exactly like the reference, and the shuffled
image is available as the extra key `hr_images` (pixel shuffle fused into the f3 epilogue).
"""
from __future__ import annotations

from collections import OrderedDict

import numpy as np
import torch

from .. import ops
from ..initializers import espcn_params
from ..params import ParamArena
from ..session import Handle, Placeholder
from ..tiling import MAX_PANEL_W, plan_tiles, shard_tiles

HALO = 4  # receptive-field radius in LR pixels: 2 (5x5) + 1 + 1


class EspcnNet:
    def __init__(self, params=None, scaling_factor=3, channels=3, device="cuda", seed=0):
        if params is None:
            params = espcn_params(seed, scaling_factor, channels)
        order = OrderedDict((k, np.asarray(params[k], np.float32)) for k in
                            ("f1/kernel:0", "f1/bias:0", "f2/kernel:0", "f2/bias:0", "f3/kernel:0", "f3/bias:0"))
        self.C = order["f1/kernel:0"].shape[2]
        self.cout3 = order["f3/kernel:0"].shape[3]
        self.r = int(round((self.cout3 // self.C) ** 0.5))
        assert self.C * self.r * self.r == self.cout3
        self.device = device
        self.arena = ParamArena(order, device)
        a = self.arena
        self.np3 = ops.pad_cout(self.cout3)
        plan = ops.PackPlan(device)
        self._i1 = plan.add(a.offsets["f1/kernel:0"], 5, self.C, 64, ops.PACK_FIRST)
        self._i2 = plan.add(a.offsets["f2/kernel:0"], 3, 64, 32, ops.PACK_FWD, 32, 64)
        self._i3 = plan.add(a.offsets["f3/kernel:0"], 3, 32, self.cout3, ops.PACK_FWD, self.np3, 32)
        self.np3f = (self.cout3 + 15) // 16 * 16  # the fused kernel pads f3's outputs to 16, not to a power of two (4x RGB: 48)
        self._i3f = self._i3 if self.np3f == self.np3 else plan.add(a.offsets["f3/kernel:0"], 3, 32, self.cout3, ops.PACK_FWD, self.np3f, 32)
        plan.finalize()
        self.plan = plan
        self.bias3 = torch.zeros(self.np3, dtype=torch.float32, device=device)
        self.repack()
        self._bufs = {}
        self._panels = {}

    def repack(self):
        self.plan.run(self.arena.w)
        self.bias3[: self.cout3].copy_(self.arena.view("f3/bias:0"))

    def load_params(self, params):
        self.arena.load_numpy(params)
        self.repack()

    # ------------------------------------------------------------------ training (reference build_model :76-94)
    def _enable_training(self, n, H, W):
        """Training runs every layer 64 channels wide (f2's 32 outputs / f3's 32 inputs zero-padded) so that the
        64x64 tensor-core dgrad / wgrad kernels serve all three layers; loss is MSE in packed space (:76-77)."""
        key = (n, H, W)
        if getattr(self, "_tb", None) is not None and self._tb["key"] == key:
            return self._tb
        a = self.arena
        a.enable_training()
        plan = ops.PackPlan(self.device)
        i = {}
        i["f1"] = plan.add(a.offsets["f1/kernel:0"], 5, self.C, 64, ops.PACK_FIRST)
        i["f2"] = plan.add(a.offsets["f2/kernel:0"], 3, 64, 32, ops.PACK_FWD, 64, 64)
        i["f3"] = plan.add(a.offsets["f3/kernel:0"], 3, 32, self.cout3, ops.PACK_FWD, 32 if self.cout3 <= 32 else 64, 64)
        i["d3"] = plan.add(a.offsets["f3/kernel:0"], 3, 32, self.cout3, ops.PACK_DGRAD, 64, 64)
        i["d2"] = plan.add(a.offsets["f2/kernel:0"], 3, 64, 32, ops.PACK_DGRAD, 64, 64)
        plan.finalize()
        stride = (ops.wgrad_workspace_bytes(n, H, W) + 1023) // 1024 * 1024
        np3 = plan.views[i["f3"]].shape[1]
        self._tb = {
            "key": key, "plan": plan, "idx": i, "stride": stride,
            "a1": ops.fpa_empty(n, H, W, 64, self.device), "a2": ops.fpa_empty(n, H, W, 64, self.device),
            "d": [ops.fpa_empty(n, H, W, 64, self.device) for _ in range(2)], "dP": ops.fpa_empty(n, H, W, 64, self.device),
            "sr": torch.empty((n, H, W, self.cout3), dtype=torch.float32, device=self.device),
            "dsr": torch.empty((n, H, W, self.cout3), dtype=torch.float32, device=self.device),
            "loss": torch.zeros(1, dtype=torch.float32, device=self.device),
            "bias2": torch.zeros(64, dtype=torch.float32, device=self.device),
            "bias3": torch.zeros(np3, dtype=torch.float32, device=self.device),
            "ws": torch.empty(stride * 2, dtype=torch.uint8, device=self.device),
            "dsts": ops.make_wgrad_dsts([(a.view("f2/kernel:0", "g"), a.view("f2/bias:0", "g"), 64, 32),
                                         (a.view("f3/kernel:0", "g"), a.view("f3/bias:0", "g"), 32, self.cout3)], self.device),
        }
        self.step = getattr(self, "step", 0)
        self._repack_train()
        return self._tb

    def _repack_train(self):
        b, a = self._tb, self.arena
        b["plan"].run(a.w)
        b["bias2"][:32].copy_(a.view("f2/bias:0"))
        b["bias3"][: self.cout3].copy_(a.view("f3/bias:0"))

    def forward_backward(self, lr: torch.Tensor, hr_packed: torch.Tensor, numel_total: float | None = None):
        n, H, W, _ = lr.shape
        assert W <= MAX_PANEL_W
        b, a = self._enable_training(n, H, W), self.arena
        V, ix = b["plan"].views, b["idx"]
        ops.conv_first_tc(lr, V[ix["f1"]], a.view("f1/bias:0"), 5, "SAME", "tanh", out=b["a1"])
        ops.conv_tc(b["a1"], V[ix["f2"]], b["bias2"], 3, "tanh", out=b["a2"])
        ops.conv_tc_last(b["a2"], V[ix["f3"]], b["bias3"], 3, self.cout3, None, out=b["sr"])
        b["loss"].zero_()
        ops.mse_fwd_bwd(b["sr"], hr_packed, b["loss"], b["dsr"], numel_total)
        # backward: f3 (linear) -> f2 (tanh) -> f1 (tanh)
        st = b["stride"]
        ops.nhwc_to_fpa_pad(b["dsr"], 64, out=b["dP"])
        ops.conv_wgrad_tc(b["a2"], b["dP"], None, None, workspace=b["ws"][st:2 * st])
        d2 = ops.conv_tc(b["dP"], V[ix["d3"]], None, 3, None, out=b["d"][0], mask_src=b["a2"], mask_kind="tanh")
        ops.conv_wgrad_tc(b["a1"], d2, None, None, workspace=b["ws"][0:st])
        d1 = ops.conv_tc(d2, V[ix["d2"]], None, 3, None, out=b["d"][1], mask_src=b["a1"], mask_kind="tanh")
        ops.wgrad_reduce_many(b["ws"], st, 2, n, H, W, b["dsts"])
        a.view("f1/kernel:0", "g").zero_()
        a.view("f1/bias:0", "g").zero_()
        ops.conv_first_wgrad(lr, d1, 5, a.view("f1/kernel:0", "g"), a.view("f1/bias:0", "g"))
        return b

    def train_step(self, lr: torch.Tensor, hr_packed: torch.Tensor, learning_rate: float, group=None):
        """One Adam step (reference `session.run(model['optimizer'], feed_dict={learning_rate: ...})`); returns the pre-update loss."""
        world = 1
        if group is not None or (torch.distributed.is_available() and torch.distributed.is_initialized()):
            world = torch.distributed.get_world_size(group)
        b = self.forward_backward(lr, hr_packed, float(hr_packed.numel()) * world)
        a = self.arena
        if world > 1:
            ops.comm_init(group)
            ops.allreduce_grads(a.g)
        self.step += 1
        ops.adam_step(a.w, a.g, a.m, a.v, learning_rate, self.step)
        self._repack_train()
        self.repack()
        return b["loss"]

    def make_graphed_step(self, lr_static: torch.Tensor, hr_static: torch.Tensor, group=None):
        """The Adam training step (forward, packed-space MSE, backward, data-parallel all-reduce, Adam, re-pack) captured into ONE
        CUDA graph; returns `step(learning_rate) -> loss buffer`.  New batches are copied INTO the static tensors before each call
        (reference loop: espcn/espcn/experiment_train.py:60-95 around model_espcn.py:64-96)."""
        world = 1
        if group is not None or (torch.distributed.is_available() and torch.distributed.is_initialized()):
            world = torch.distributed.get_world_size(group)
        a = self.arena
        a.enable_training()
        numel = float(hr_static.numel()) * world
        update, _ = ops.make_exchange_and_adam(a, group)  # fused exchange + Adam over NVLink peer memory, or NCCL + Adam

        def body(lr_t):
            self.forward_backward(lr_static, hr_static, numel)
            update(lr_t)
            self._repack_train()
            self.repack()

        gstep = ops.graph_training_step(body, a)

        def step(learning_rate: float):
            self.step += 1
            gstep(learning_rate)
            return self._tb["loss"]

        step.graph = gstep.graph
        return step

    def _get_bufs(self, n, H, W):
        key = (n, H, W)
        if key not in self._bufs:
            if len(self._bufs) > 8:
                self._bufs.clear()
            self._bufs[key] = (ops.fpa_empty(n, H, W, 64, self.device), ops.fpa_empty(n, H, W, 32, self.device))
        return self._bufs[key]

    def forward_fused(self, lr: torch.Tensor, shuffle=True, out: torch.Tensor | None = None, rank=0, world=1, uint8=False) -> torch.Tensor:
        """The test graph (reference model_espcn.py:117-134 + the un-pack of experiment_test.py:173-177) as ONE kernel
        (srk_espcn_forward): activations stay in tensor memory.  Rank `rank` of `world` produces its LR row band of every
        frame (reads 4 halo rows from `lr`, writes only its band of `out`)."""
        n, H, W, C = lr.shape
        assert C == self.C
        a, V = self.arena, self.plan.views
        y0, y1 = (H * rank) // world, (H * (rank + 1)) // world
        return ops.espcn_forward(lr, V[self._i1], a.view("f1/bias:0"), V[self._i2], a.view("f2/bias:0"), V[self._i3f], a.view("f3/bias:0"),
                                 self.r, shuffle, out, (y0, y1), uint8)

    # ------------------------------------------------------------------ host-to-host inference (the session.run seam)
    def _host_state(self, n, H, W, shuffle, uint8):
        key = (n, H, W, shuffle, uint8)
        st = getattr(self, "_hs", None)
        if st is None or st["key"] != key:
            r, C = self.r, self.C
            oshape = (n, H * r, W * r, C) if shuffle else (n, H, W, self.cout3)
            odt = torch.uint8 if uint8 else torch.float32
            st = {"key": key, "lr_dev": torch.empty((n, H, W, C), dtype=torch.float32, device=self.device),
                  "lr_u8_dev": torch.empty((n, H, W, C), dtype=torch.uint8, device=self.device),
                  "lr_u8_pin": torch.empty((n, H, W, C), dtype=torch.uint8).pin_memory(),
                  "out_dev": torch.empty(oshape, dtype=odt, device=self.device),
                  "lr_pin": torch.empty((n, H, W, C), dtype=torch.float32).pin_memory(),
                  "out_pin": torch.empty(oshape, dtype=odt).pin_memory()}
            self._hs = st
        return st

    def forward_host(self, lr_host, out_host=None, shuffle=True, uint8=False, band_rows=540):
        """Host array in, host array out -- what `session.run(sr, feed_dict={lr: frames})` does in the reference
        (espcn/espcn/experiment_test.py:164-184) -- with the copies hidden: every frame is cut into row bands, and the
        host->device copy of band k+1, the fused kernel on band k and the device->host copy of band k-1 run on three streams.
        `lr_host` / `out_host`: numpy arrays or CPU tensors; page-locked ones (`torch.Tensor.pin_memory()`, or the arrays
        `pinned_like()` returns) are used in place, pageable ones go through a pinned staging copy.  uint8=True returns
        saturate_cast(x * 127.5 + 127.5) as the reference's PNG writer does (:179-184).  Returns `out_host` (or a staging
        buffer that the next call overwrites when none was given)."""
        if isinstance(lr_host, torch.Tensor):
            lr_t = lr_host
        else:
            lr_host = np.asarray(lr_host)
            lr_t = torch.from_numpy(np.ascontiguousarray(lr_host, dtype=np.uint8 if lr_host.dtype == np.uint8 else np.float32))
        n, H, W, C = lr_t.shape
        # a uint8 frame is the RAW image: it crosses PCIe at one byte per sample and the drivers' `image / 127.5 - 1.0`
        # (experiment_test.py:159) runs on the device (srk_u8_to_pm1_f64: bit-identical to the host arithmetic)
        raw = lr_t.dtype == torch.uint8
        assert C == self.C and (raw or lr_t.dtype == torch.float32) and lr_t.is_contiguous()
        st = self._host_state(n, H, W, shuffle, uint8)
        if not lr_t.is_pinned():
            pin = st["lr_u8_pin"] if raw else st["lr_pin"]
            pin.copy_(lr_t)
            lr_t = pin
        out_t = st["out_pin"]
        user_out = None
        if out_host is not None:
            cand = out_host if isinstance(out_host, torch.Tensor) else torch.from_numpy(out_host)
            assert cand.shape == st["out_dev"].shape and cand.dtype == st["out_dev"].dtype and cand.is_contiguous()
            if cand.is_pinned():
                out_t = cand
            else:
                user_out = cand
        a, V = self.arena, self.plan.views
        ops.espcn_forward_host(lr_t, out_t, V[self._i1], a.view("f1/bias:0"), V[self._i2], a.view("f2/bias:0"), V[self._i3f], a.view("f3/bias:0"), self.r,
                               shuffle, uint8, st["lr_dev"], st["lr_u8_dev"] if raw else None, st["out_dev"], band_rows)
        if user_out is not None:
            user_out.copy_(out_t)
            out_t = user_out
        return out_t if isinstance(out_host, torch.Tensor) or out_host is None else out_host

    def forward(self, lr: torch.Tensor, shuffle=True, out: torch.Tensor | None = None, rank=0, world=1, tile_rows=None, fused=True) -> torch.Tensor:
        """lr fp32 [N,h,w,C] -> shuffled [N,h*r,w*r,C] (shuffle=True) or packed [N,h,w,C*r^2].  fused=False runs the three
        layers as separate kernels through FPA buffers in HBM (the form training uses; kept for A/B measurements)."""
        n, H, W, C = lr.shape
        assert C == self.C
        if fused and tile_rows is None:
            return self.forward_fused(lr, shuffle, out, rank, world)
        a, r = self.arena, (self.r if shuffle else 1)
        if out is None:
            out = torch.empty((n, H * r, W * r, self.cout3 // (r * r)), dtype=torch.float32, device=lr.device)
        b1, b2 = a.view("f1/bias:0"), a.view("f2/bias:0")
        w1p, w2p, w3p = self.plan.views[self._i1], self.plan.views[self._i2], self.plan.views[self._i3]
        if W <= MAX_PANEL_W and world == 1 and (tile_rows is None or H <= tile_rows):
            t1, t2 = self._get_bufs(n, H, W)
            ops.conv_first_tc(lr, w1p, b1, 5, "SAME", "tanh", out=t1)
            ops.conv_tc(t1, w2p, b2, 3, "tanh", out=t2)
            ops.conv_tc_last(t2, w3p, self.bias3, 3, self.cout3, None, shuffle_r=r, out=out)
            return out
        Ht, Wt, tiles = plan_tiles(n, H, W, HALO, MAX_PANEL_W, tile_rows)
        tiles = shard_tiles(tiles, rank, world)
        if not tiles:
            return out
        key = tuple(t.as_tuple() for t in tiles)
        if key not in self._panels:
            self._panels[key] = ops.make_panels(list(key), self.device)
        panels = self._panels[key]
        t1, t2 = self._get_bufs(len(tiles), Ht, Wt)
        ops.conv_first_tc(lr, w1p, b1, 5, "SAME", "tanh", panels=panels, panel_hw=(Ht, Wt), out=t1)
        ops.conv_tc(t1, w2p, b2, 3, "tanh", out=t2)
        ops.conv_tc_last(t2, w3p, self.bias3, 3, self.cout3, None, shuffle_r=r, panels=panels, frame_shape=(n, H, W), out=out)
        return out


# ---------------------------------------------------------------------------------------------
# reference-shaped builders
# ---------------------------------------------------------------------------------------------


class _EspcnGraph:
    def __init__(self, net, lr_ph, hr_ph=None):
        self.net, self.lr_ph, self.hr_ph = net, lr_ph, hr_ph
        self.lr_u8_ph = Placeholder("lr_source_u8", getattr(lr_ph, "shape", None))  # raw uint8 frames (normalised on the device)

    @staticmethod
    def _to_device(x, device):
        """One path for every feed: numpy array, CPU tensor or device tensor -> contiguous fp32 device tensor."""
        t = x if isinstance(x, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32))
        return t.to(device, torch.float32, non_blocking=True).contiguous()

    def execute(self, keys, feeds, outs=None):
        net = self.net
        out = {}
        outs = outs or {}
        raw = self.lr_u8_ph in feeds
        x = feeds[self.lr_u8_ph] if raw else feeds[self.lr_ph]
        if raw:
            assert not (keys & {"optimizer", "loss"}) and not (isinstance(x, torch.Tensor) and x.is_cuda), "lr_source_u8 feeds host frames to the inference fetches"
        if not (keys & {"optimizer", "loss"}) and not (isinstance(x, torch.Tensor) and x.is_cuda):
            # inference on host arrays: band-pipelined copies around the fused kernel (EspcnNet.forward_host)
            from ..session import as_host_tensor
            lr_h = as_host_tensor(x, np.uint8 if raw else np.float32)
            for key, shuffle, u8 in (("sr_result", False, False), ("sr_results", False, False), ("hr_images", True, False), ("hr_images_u8", True, True)):
                if key in keys and key not in out:
                    dst = outs.get(key)
                    res = net.forward_host(lr_h, None if dst is None else as_host_tensor(dst), shuffle, u8)
                    val = dst if dst is not None else res.numpy().copy()  # a fresh array, as session.run returns; pass out= to avoid the copy
                    out[key] = val
                    if key in ("sr_result", "sr_results"):
                        out["sr_result"] = out["sr_results"] = val
            out["step"] = getattr(net, "step", 0)
            out["scaling_factor"] = net.r
            return out
        lr = self._to_device(x, net.device)
        if "optimizer" in keys:
            hr = self._to_device(feeds[self.hr_ph], net.device)
            rate = [v for k, v in feeds.items() if isinstance(k, Placeholder) and k.name == "learning_rate"]
            step_before = getattr(net, "step", 0)
            loss = net.train_step(lr, hr, float(rate[0]) if rate else 0.01)
            out.update(optimizer=None, loss=float(loss), step=step_before)
            if keys & {"sr_result", "sr_results"}:
                out["sr_result"] = out["sr_results"] = net._tb["sr"].cpu().numpy()
            return out
        if keys & {"sr_result", "sr_results", "loss"}:
            packed = net.forward(lr, shuffle=False)
            out["sr_result"] = out["sr_results"] = packed.cpu().numpy()
            if "loss" in keys:
                hr = self._to_device(feeds[self.hr_ph], net.device)
                acc = torch.zeros(1, device=net.device)
                ops.mse_fwd_bwd(packed, hr, acc, None)
                out["loss"] = float(acc)
        if "hr_images" in keys:
            out["hr_images"] = net.forward(lr, shuffle=True).cpu().numpy()
        if "hr_images_u8" in keys:
            out["hr_images_u8"] = net.forward_fused(lr, shuffle=True, uint8=True).cpu().numpy()
        out["step"] = getattr(net, "step", 0)
        out["scaling_factor"] = net.r
        return out


def build_model(lr_source, scaling_factor=3, hr_target=None, params=None, channels=3, device="cuda", seed=0):
    """espcn/espcn/model_espcn.py:6 `build_model(lr_source, scaling_factor=3, hr_target=None)`."""
    net = EspcnNet(params, scaling_factor, channels, device, seed)
    g = _EspcnGraph(net, lr_source, hr_target)
    model = {"lr_source": lr_source, "lr_source_u8": g.lr_u8_ph, "sr_result": Handle(g, "sr_result"), "hr_images": Handle(g, "hr_images"),
             "hr_images_u8": Handle(g, "hr_images_u8")}
    if hr_target is None:
        return model
    model["hr_target"] = hr_target
    model["step"] = Handle(g, "step")
    model["loss"] = Handle(g, "loss")
    model["optimizer"] = Handle(g, "optimizer")
    model["learning_rate"] = Placeholder("learning_rate", [])
    return model


def extract_weights(meta_path, ckpt_path):
    """espcn/espcn/model_espcn.py:150-166 returns {tf_var_name: ndarray}.  `ckpt_path` is an `.npz` file keyed by the TF variable
    names or a TensorFlow Saver checkpoint prefix (read without TensorFlow); `meta_path` is accepted and ignored: the graph is
    rebuilt from the variable shapes, not imported."""
    from ..params import load_params
    return {k: v for k, v in load_params(ckpt_path).items() if k.endswith(("kernel:0", "bias:0"))}


def build_test_model(meta_path, ckpt_path, device="cuda"):
    """espcn/espcn/model_espcn.py:99-147: constant-weight test graph; `sr_results` is packed."""
    variables = extract_weights(meta_path, ckpt_path)
    scaling_factor = int((variables["f3/bias:0"].size // 3) ** 0.5)
    lr_sources = Placeholder("lr_sources", [None, None, None, 3])
    net = EspcnNet(variables, scaling_factor, 3, device)
    g = _EspcnGraph(net, lr_sources)
    return {"lr_sources": lr_sources, "sr_results": Handle(g, "sr_results"), "scaling_factor": scaling_factor,
            "hr_images": Handle(g, "hr_images"), "hr_images_u8": Handle(g, "hr_images_u8")}
