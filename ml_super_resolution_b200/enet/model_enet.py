"""EnhanceNet on the B200 conv hot path -- drop-in for enet/enet/model_enet.py: `build_generator` / `residual_block` (below) and
`build_enet` with the discriminator / VGG-19 perceptual / texture-matching losses (`losses.py`).

  conv2d            3x3 3->64 ReLU                      srk_conv_first_tc                    reference :63-70
  10 x residual     3x3 64->64 ReLU ; 1x1 64->64 ;      srk_conv_tc ; srk_conv_tc(k=1) with  reference :8-31
                    ReLU(x + y)                         the block residual + ReLU fused in the epilogue
  2 x upsample      NN x2 ; 3x3 64->64 ReLU             srk_fpa_upsample2 ; srk_conv_tc      reference :77-89
  conv2d_23         3x3 64->64 ReLU                     srk_conv_tc                          reference :92-99
  conv2d_24         3x3 64->3, + bq_images              srk_conv_tc_last (fused addend)      reference :102-113

Backward of the generator for a supplied d(sr) (`forward_backward`; what `minimize(..., var_list=g_variables)` adds,
reference :336-341 -- the VGG / texture / adversarial loss terms that produce d(sr) are outside the hot path):
  every stored gradient is taken w.r.t. the PRE-activation of the layer that produced the tensor, so one dgrad launch
  = conv with SRK_PACK_DGRAD weights * ReLU'(saved input); at a residual junction the skip gradient joins before the
  mask (srk_conv_tc relu_after_add = 2); NN-upsample backward = srk_fpa_upsample2_bwd of the masked gradient;
  wgrad = srk_conv_wgrad_tc per layer (1x1 layers keep the centre tap) + one srk_wgrad_reduce_many per resolution.
"""
from __future__ import annotations

from collections import OrderedDict

import numpy as np
import torch

from .. import ops
from ..initializers import ENET_G_LAYERS, enet_g_params, tf_conv_name
from ..params import ParamArena
from ..session import Handle
from ..tiling import MAX_PANEL_W, plan_seam_exchange


class EnetGenerator:
    def __init__(self, params=None, scope_name="g_", device="cuda", seed=0):
        self.scope = scope_name
        if params is None:
            params = enet_g_params(seed, scope_name)
        order = OrderedDict()
        for i in range(len(ENET_G_LAYERS)):
            for suf in ("kernel:0", "bias:0"):
                k = f"{scope_name}/{tf_conv_name(i)}/{suf}"
                order[k] = np.asarray(params[k], np.float32)
        self.device = device
        self.arena = ParamArena(order, device, decay_suffix=None)
        a = self.arena
        plan = ops.PackPlan(device)
        self._idx = {}
        for i, (k, cin, cout) in enumerate(ENET_G_LAYERS):
            off = a.offsets[self._k(i)]
            if i == 0:
                self._idx[i] = plan.add(off, 3, cin, 64, ops.PACK_FIRST)
            elif cout == 64:
                self._idx[i] = plan.add(off, k, 64, 64, ops.PACK_FWD, 64, 64)
            else:
                self._idx[i] = plan.add(off, 3, 64, cout, ops.PACK_FWD, 16, 64)
        plan.finalize()
        self.plan = plan
        self.bias_last = torch.zeros(16, dtype=torch.float32, device=device)
        self.repack()

    def _k(self, i):
        return f"{self.scope}/{tf_conv_name(i)}/kernel:0"

    def _b(self, i):
        return f"{self.scope}/{tf_conv_name(i)}/bias:0"

    def repack(self):
        self.plan.run(self.arena.w)
        self.bias_last[:3].copy_(self.arena.view(self._b(len(ENET_G_LAYERS) - 1)))

    HALO = 13  # receptive-field radius in LR pixels: row bands recompute it, column panels exchange seams

    def forward(self, sd: torch.Tensor, bq: torch.Tensor, out: torch.Tensor | None = None, tile_rows: int | None = None,
                rank: int = 0, world: int = 1, max_panel_w: int = MAX_PANEL_W // 4) -> torch.Tensor:
        """sd fp32 [N,h,w,3], bq fp32 [N,4h,4w,3] -> sr = bq + generator(sd).
        Frames wider than 63 LR pixels (252 HR) are cut into column panels that swap their seam columns after every 3x3 layer
        at all three resolutions (`srk_fpa_halo_exchange`), frames taller than `tile_rows` LR rows into row bands with a 13-px
        halo; with world > 1 this rank computes its shard of the tile list and writes only the pixels it owns (seams are only exchanged
        when the shard consists of whole bands; otherwise the column halo is recomputed as well)."""
        n, h, w, _ = sd.shape
        assert max_panel_w <= MAX_PANEL_W // 4
        a, W = self.arena, self.plan.views
        if out is None:
            out = torch.empty_like(bq)
        need_tiles = w > max_panel_w or (tile_rows is not None and h > tile_rows) or world > 1
        panels = [None, None, None]
        max_cols = 0
        if need_tiles:
            Ht, Wt, tiles, exchange, max_cols = plan_seam_exchange(n, h, w, self.HALO, max_panel_w, tile_rows, rank, world)
            # (exchange False: a band is split between ranks -> the plan carries the 13-px halo in x as well, no seam traffic)
            if not tiles:
                return out
            key = (tuple(t.as_tuple() for t in tiles), str(sd.device))
            cache = self.__dict__.setdefault("_panels", {})
            if key not in cache:
                cache[key] = [ops.make_panels([(t.frame,) + tuple(s * v for v in t.as_tuple()[1:]) for t in tiles], self.device)
                              for s in (1, 2, 4)]
            panels = cache[key]
            n_img, hh, ww = len(tiles), Ht, Wt
        else:
            n_img, hh, ww = n, h, w

        def seam(t, level):  # refresh the panels' non-owned columns after a 3x3 layer
            if need_tiles and max_cols > 0:
                ops.fpa_halo_exchange(t, panels[level], max_cols << level)
            return t

        t = seam(ops.conv_first_tc(sd, W[self._idx[0]], a.view(self._b(0)), 3, "SAME", "relu", panels=panels[0],
                                   panel_hw=(hh, ww) if need_tiles else None), 0)
        i = 1
        # the FRAME's width at each resolution decides the kernel form of the 3x3 layers (ops.conv_form): panels compute what
        # the un-tiled frame would
        for _ in range(10):
            with ops.conv_form(w):
                x = seam(ops.conv_tc(t, W[self._idx[i]], a.view(self._b(i)), 3, "relu"), 0)
            t = ops.conv_tc(x, W[self._idx[i + 1]], a.view(self._b(i + 1)), 1, None, addend=t, relu_after_add=True)
            i += 2
        for level in (1, 2):
            with ops.conv_form(w << level):
                t = seam(ops.conv_tc(ops.fpa_upsample2(t), W[self._idx[i]], a.view(self._b(i)), 3, "relu"), level)
            i += 1
        with ops.conv_form(4 * w):
            t = seam(ops.conv_tc(t, W[self._idx[i]], a.view(self._b(i)), 3, "relu"), 2)
        return ops.conv_tc_last(t, W[self._idx[i + 1]], self.bias_last, 3, 3, None, addend=bq, panels=panels[2],
                                frame_shape=(n, 4 * h, 4 * w) if need_tiles else None, out=out)

    # ------------------------------------------------------------------------------------------ training
    def _enable_training(self, n, h, w):
        key = (n, h, w)
        if getattr(self, "_tb", None) is not None and self._tb["key"] == key:
            return self._tb
        a, dev = self.arena, self.device
        a.enable_training()
        L = len(ENET_G_LAYERS)
        plan = ops.PackPlan(dev)
        fw, dg = {}, {}
        for i, (k, cin, cout) in enumerate(ENET_G_LAYERS):
            off = a.offsets[self._k(i)]
            if i == 0:
                fw[i] = plan.add(off, 3, cin, 64, ops.PACK_FIRST)
            elif i == L - 1:
                fw[i] = plan.add(off, 3, 64, cout, ops.PACK_FWD, 16, 64)
                dg[i] = plan.add(off, 3, 64, cout, ops.PACK_FIRST_ROT180T)
            else:
                fw[i] = plan.add(off, k, 64, 64, ops.PACK_FWD, 64, 64)
                dg[i] = plan.add(off, k, 64, 64, ops.PACK_DGRAD, 64, 64)
        plan.finalize()
        geo = {"lo": (h, w), "mid": (2 * h, 2 * w), "hi": (4 * h, 4 * w)}
        F = lambda g: ops.fpa_empty(n, geo[g][0], geo[g][1], 64, dev)  # noqa: E731
        st = {g: (ops.wgrad_workspace_bytes(n, *geo[g]) + 1023) // 1024 * 1024 for g in geo}
        gv = lambda i, b=False: a.view(self._b(i) if b else self._k(i), "g")  # noqa: E731
        tmp1 = torch.zeros((10, 9, 64, 64), dtype=torch.float32, device=dev)  # 1x1 layers: all nine taps, the centre one is kept
        # wgrad jobs per resolution, in launch order: (layer index, destination kernel buffer, ci_n, co_n)
        jobs = {"hi": [(24, gv(24), 64, 3), (23, gv(23), 64, 64), (22, gv(22), 64, 64)], "mid": [(21, gv(21), 64, 64)], "lo": []}
        for b in range(9, -1, -1):
            jobs["lo"].append((2 + 2 * b, tmp1[b], 64, 64))
            jobs["lo"].append((1 + 2 * b, gv(1 + 2 * b), 64, 64))
        jobs["lo"].append((0, gv(0), 3, 64))
        self._tb = {
            "key": key, "plan": plan, "fw": fw, "dg": dg, "geo": geo, "st": st, "tmp1": tmp1, "jobs": jobs,
            "t": [F("lo") for _ in range(11)], "hb": [F("lo") for _ in range(10)],
            "u1": F("mid"), "m1": F("mid"), "u2": F("hi"), "m2": F("hi"), "m3": F("hi"),
            # every layer's dY is kept: the weight gradients of a resolution run in one batched launch after the dgrad chain
            "dhi": [F("hi") for _ in range(3)], "dmid": [F("mid") for _ in range(2)], "dlo": [F("lo") for _ in range(21)],
            "sdF": F("lo"), "dP": F("hi"),
            "ws": {g: torch.empty(st[g] * max(1, len(jobs[g])), dtype=torch.uint8, device=dev) for g in geo},
            "dsts": {g: ops.make_wgrad_dsts([(dw, gv(i, True), ci, co) for (i, dw, ci, co) in jobs[g]], dev) for g in geo},
            "sr": torch.empty((n, 4 * h, 4 * w, 3), dtype=torch.float32, device=dev),
        }
        self._tb["plan"].run(a.w)
        return self._tb

    def forward_backward(self, sd: torch.Tensor, bq: torch.Tensor, dsr):
        """Generator forward, then its backward for the upstream gradient d(loss)/d(sr): fills the gradient arena
        (`arena.g`, every kernel and bias of the 25 layers) and returns sr.  `dsr` is that gradient as a tensor, or a
        callable `dsr = f(sr)` evaluated on the forward result (the loss head), so one forward serves loss and backward."""
        n, h, w, _ = sd.shape
        assert 4 * w <= MAX_PANEL_W
        b, a = self._enable_training(n, h, w), self.arena
        V, fw, dg = b["plan"].views, b["fw"], b["dg"]
        bias = lambda i: a.view(self._b(i))  # noqa: E731
        # ---- forward, keeping every activation
        t = b["t"]
        ops.conv_first_tc(sd, V[fw[0]], bias(0), 3, "SAME", "relu", out=t[0])
        for k in range(10):
            i = 1 + 2 * k
            ops.conv_tc(t[k], V[fw[i]], bias(i), 3, "relu", out=b["hb"][k])
            ops.conv_tc(b["hb"][k], V[fw[i + 1]], bias(i + 1), 1, None, out=t[k + 1], addend=t[k], relu_after_add=1)
        ops.fpa_upsample2(t[10], out=b["u1"])
        ops.conv_tc(b["u1"], V[fw[21]], bias(21), 3, "relu", out=b["m1"])
        ops.fpa_upsample2(b["m1"], out=b["u2"])
        ops.conv_tc(b["u2"], V[fw[22]], bias(22), 3, "relu", out=b["m2"])
        ops.conv_tc(b["m2"], V[fw[23]], bias(23), 3, "relu", out=b["m3"])
        ops.conv_tc_last(b["m3"], V[fw[24]], self.bias_last, 3, 3, None, addend=bq, out=b["sr"])
        if callable(dsr):
            dsr = dsr(b["sr"])
        # ---- backward: dgrad chain first (all dY kept), then one batched wgrad + reduce launch per resolution, in the order
        # of b["jobs"] (hi: conv2d_24, 23, 22; mid: conv2d_21; lo: blocks 9..0 as (1x1, 3x3), then the first layer)
        dhi, dmid, dlo = b["dhi"], b["dmid"], b["dlo"]
        xs = {"hi": [], "mid": [], "lo": []}
        dys = {"hi": [], "mid": [], "lo": []}

        def wgrad(g, x, dy):
            xs[g].append(x)
            dys[g].append(dy)

        ops.nhwc_to_fpa_pad(dsr, 64, out=b["dP"])
        wgrad("hi", b["m3"], b["dP"])                                                                            # conv2d_24
        d = ops.conv_first_tc(dsr, V[dg[24]], None, 3, "SAME", None, out=dhi[0], mask_src=b["m3"], mask_kind="relu")
        wgrad("hi", b["m2"], d)                                                                                  # conv2d_23
        d = ops.conv_tc(d, V[dg[23]], None, 3, None, out=dhi[1], mask_src=b["m2"], mask_kind="relu")
        wgrad("hi", b["u2"], d)                                                                                  # conv2d_22
        d = ops.conv_tc(d, V[dg[22]], None, 3, None, out=dhi[2], mask_src=b["u2"], mask_kind="relu")
        d = ops.fpa_upsample2_bwd(d, out=dmid[0])
        wgrad("mid", b["u1"], d)                                                                                 # conv2d_21
        d = ops.conv_tc(d, V[dg[21]], None, 3, None, out=dmid[1], mask_src=b["u1"], mask_kind="relu")
        ds = ops.fpa_upsample2_bwd(d, out=dlo[0])
        nxt = 1
        for k in range(9, -1, -1):                                                                               # residual blocks, last first
            i = 1 + 2 * k
            wgrad("lo", b["hb"][k], ds)                                                                          # 1x1
            dh = ops.conv_tc(ds, V[dg[i + 1]], None, 1, None, out=dlo[nxt], mask_src=b["hb"][k], mask_kind="relu")
            wgrad("lo", t[k], dh)                                                                                # 3x3
            ds = ops.conv_tc(dh, V[dg[i]], None, 3, None, out=dlo[nxt + 1], mask_src=t[k], mask_kind="relu", addend=ds, relu_after_add=2)
            nxt += 2
        ops.nhwc_to_fpa_pad(sd, 64, out=b["sdF"])
        wgrad("lo", b["sdF"], ds)                                                                                # conv2d (first layer)
        for g in ("hi", "mid", "lo"):
            ops.conv_wgrad_tc_batched(xs[g], dys[g], b["ws"][g], b["st"][g], b["dsts"][g])
        for k in range(10):
            a.view(self._k(2 + 2 * k), "g").copy_(b["tmp1"][k, 4].view(1, 1, 64, 64))
        return b["sr"]


    def make_graphed_step(self, sd_static: torch.Tensor, bq_static: torch.Tensor, loss_head, group=None):
        """Capture one generator training step -- forward, `loss_head(sr) -> d(loss)/d(sr)` (kernels writing static buffers),
        backward, the data-parallel gradient all-reduce (srk_allreduce_grads, when a process group is up), Adam(lr, .9, .999) and
        the weight re-pack -- into ONE CUDA graph: the ~130 launches of a step are latency-bound at 64 patches of 32x32.
        Returns `step(lr)`; new batches are copied INTO the static tensors before each call
        (reference: `session.run(g_trainer)` of enet/enet/experiment_train.py:100-130 with model_enet.py:336-341)."""
        a = self.arena
        a.enable_training()
        update, _ = ops.make_exchange_and_adam(a, group)  # fused exchange + Adam over NVLink peer memory, or NCCL + Adam

        def body(lr_t):
            self.forward_backward(sd_static, bq_static, loss_head)
            update(lr_t)
            self._tb["plan"].run(a.w)
            self.repack()

        gstep = ops.graph_training_step(body, a)

        def step(lr: float = 1e-4):
            gstep(lr)

        step.graph = gstep.graph
        return step


class _EnetGraph:
    def __init__(self, net, sd_ph, bq_ph):
        self.net, self.sd_ph, self.bq_ph = net, sd_ph, bq_ph

    def execute(self, keys, feeds):
        dev = self.net.device
        to = lambda x: (x if isinstance(x, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32))).to(dev).contiguous()
        return {"sr_images": self.net.forward(to(feeds[self.sd_ph]), to(feeds[self.bq_ph])).cpu().numpy()}


def build_generator(sd_images, bq_images, hd_images=None, scope_name="g_", params=None, device="cuda", seed=0):
    """enet/enet/model_enet.py:44 `build_generator(sd_images, bq_images, hd_images, scope_name)` -> sr_images handle."""
    net = EnetGenerator(params, scope_name, device, seed)
    return Handle(_EnetGraph(net, sd_images, bq_images), "sr_images")


# ---------------------------------------------------------------------------------------------
# build_enet: generator + the training losses (enet/enet/model_enet.py:264-348)
# ---------------------------------------------------------------------------------------------


class EnetPat:
    """ENet-P / -PA / -PAT training state: generator, discriminator, constant VGG-19, and one training step
    (`g_trainer` = Adam(1e-4) on g_losses w.r.t. the generator, `d_trainer` = Adam(1e-4) on a_loss w.r.t. the discriminator)."""

    def __init__(self, pat_model="pat", vgg_weights=None, params=None, d_params=None, hd_size=128, device="cuda", seed=0):
        from . import losses as L
        self.pat = pat_model
        self.gen = EnetGenerator(params, "g_", device, seed)
        self.vgg = L.Vgg19(vgg_weights if vgg_weights is not None else L.vgg19_random_weights(seed), device)
        self.disc = L.Discriminator(hd_size, d_params, "d_", device, seed + 1) if "a" in pat_model else None
        self.device = device
        self.step = 0
        self.last = {}

    def losses_and_dsr(self, sr: torch.Tensor, hd: torch.Tensor) -> torch.Tensor:
        """The loss section of build_enet (:288-322) on the device: fills `self.last` with device scalars p_loss / g_loss /
        t_loss / g_loss_all and returns d(g_losses)/d(sr)."""
        from . import losses as L
        from .. import nn
        z = lambda: torch.zeros(1, dtype=torch.float32, device=sr.device)  # noqa: E731
        p_loss, g_loss, t_loss = z(), z(), z()
        hd_acts, sr_acts = self.vgg.forward(hd), self.vgg.forward(sr)
        taps = L.perceptual_loss(self.vgg, sr_acts, hd_acts, p_loss)
        if "t" in self.pat:
            taps.update(L.texture_matching_loss(sr_acts, hd_acts, t_loss))
        dsr = torch.zeros_like(sr)
        self.vgg.backward(sr_acts, taps, dsr, accumulate=False)
        if "a" in self.pat:
            scale = 2.0 if "t" in self.pat else 1.0
            g_scaled = z()
            nn.axpby(self.disc.generator_loss(sr, g_scaled, scale), dsr, 1.0, 1.0)
            nn.axpby(g_scaled, g_loss, 1.0 / scale, 0.0)
        total = z()
        nn.axpby(p_loss, total, 1.0, 0.0)
        if "a" in self.pat:
            nn.axpby(g_loss, total, 2.0 if "t" in self.pat else 1.0, 1.0)
        if "t" in self.pat:
            nn.axpby(t_loss, total, 1.0, 1.0)
        self.last = {"p_loss": p_loss, "g_loss": g_loss, "t_loss": t_loss, "g_loss_all": total}
        return dsr

    def train_step(self, sd, bq, hd, g_train=True, d_train=True, learning_rate=1e-4):
        """One `session.run({g_trainer, d_trainer, losses})` (enet/enet/experiment_train.py): both trainers use the generator output
        computed with the weights from before the step, as TF evaluates them in one graph run."""
        gen = self.gen
        sr = gen.forward_backward(sd, bq, lambda s: self.losses_and_dsr(s, hd))
        if self.disc is not None:
            a_loss = torch.zeros(1, dtype=torch.float32, device=sr.device)
            self.disc.discriminator_loss_and_grads(sr, hd, a_loss)
            self.last["a_loss"] = a_loss
            if d_train:
                self.disc.adam_step(learning_rate)
        if g_train:
            a = gen.arena
            self.step += 1
            ops.adam_step(a.w, a.g, a.m, a.v, learning_rate, self.step)
            gen._tb["plan"].run(a.w)
            gen.repack()
        return sr


class _EnetPatGraph:
    def __init__(self, trainer: EnetPat, sd_ph, bq_ph, hd_ph):
        self.t, self.sd_ph, self.bq_ph, self.hd_ph = trainer, sd_ph, bq_ph, hd_ph

    def execute(self, keys, feeds):
        t = self.t
        if keys == {"step"}:
            return {"step": t.step}
        dev = t.device
        to = lambda x: (x if isinstance(x, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32))).to(dev).contiguous()  # noqa: E731
        sd, bq = to(feeds[self.sd_ph]), to(feeds[self.bq_ph])
        out = {"step": t.step}
        if keys <= {"sr_images", "step"}:
            out["sr_images"] = t.gen.forward(sd, bq).cpu().numpy()
            return out
        hd = to(feeds[self.hd_ph])
        sr = t.train_step(sd, bq, hd, g_train="g_trainer" in keys, d_train="d_trainer" in keys)
        out.update(g_trainer=None, d_trainer=None, sr_images=sr.cpu().numpy() if "sr_images" in keys else None)
        for k, v in t.last.items():
            out[k] = float(v)
        return out


def build_enet(sd_images, bq_images, hd_images, pat_model, vgg19_path, params=None, d_params=None, hd_size=128, device="cuda", seed=0):
    """enet/enet/model_enet.py:264 `build_enet(sd_images, bq_images, hd_images, pat_model, vgg19_path)` -> the same dict keys
    (`sd_images, bq_images, sr_images [, hd_images, step, p_loss, g_loss_all, g_trainer, a_loss, g_loss, d_trainer, t_loss]`).
    `vgg19_path`: the keras VGG-19 weights as `.npz` (`<layer>_W_1:0`, `<layer>_b_1:0`); None = random stand-in weights."""
    from . import losses as L
    trainer = EnetPat(pat_model, L.load_vgg_weights(vgg19_path) if vgg19_path else None, params, d_params, hd_size, device, seed)
    g = _EnetPatGraph(trainer, sd_images, bq_images, hd_images)
    model = {"sd_images": sd_images, "bq_images": bq_images, "sr_images": Handle(g, "sr_images")}
    if hd_images is None:
        return model
    model["hd_images"] = hd_images
    for k in ("step", "p_loss", "g_loss_all", "g_trainer"):
        model[k] = Handle(g, k)
    if "a" in pat_model:
        for k in ("a_loss", "g_loss", "d_trainer"):
            model[k] = Handle(g, k)
    if "t" in pat_model:
        model["t_loss"] = Handle(g, "t_loss")
    return model
