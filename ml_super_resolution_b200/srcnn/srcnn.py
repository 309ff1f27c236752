"""SRCNN 9-1-5 on the B200 conv hot path -- drop-in for srcnn/srcnn.py of the reference.

The reference is a single script that reads a global `FLAGS` (srcnn/srcnn.py:8-23) and builds, in one graph,
the JPEG reader, the bicubic degrade, the three VALID convolutions, the crop, the loss and Adam
(`build_srcnn`, :81-166).  Here the same flag names live on `FLAGS`, `sanity_check()` is the same arithmetic
(:28-43, python-2 integer division), and `build_srcnn(hi_images)` returns the same dict keys
(`step, loss, trainer, hd_images, sd_images, sr_images`) as session handles.  The file reader / queue runners are
out of scope: `hi_images` is fed like any other placeholder.

  lo = resize_bicubic(resize_bicubic(hi, S/r), S)   srk_resize_bicubic_tf1 x2        reference :89-93
  patch_extraction     9x9 VALID C->64 ReLU         srk_conv_first_tc               reference :100-108
  non_linear_mapping   1x1 VALID 64->32 ReLU        srk_conv_tc (k=1, N=32)         reference :111-119
  reconstruction       5x5 VALID 32->C tanh         srk_conv_tc_last (k=5, crop)    reference :122-130
  loss                 mean_rows ||sr - hi||_2      srk_l2norm_rows_mean_fwd_bwd    reference :142-144
  trainer              Adam(1e-3, .5, .9)           backward below + srk_adam_step  reference :155-157

Backward (what `minimize` adds, reference :155-157).  A VALID k x k layer is the SAME layer over the larger of its two
geometries with the smaller tensor embedded behind a zero border of k/2 pixels, so every gradient runs on the 3x3 /
64-channel tensor-core kernels the other models use:
  wgrad of a k x k kernel = ceil(k/3)^2 calls of srk_conv_wgrad_tc, one per 3x3 block of taps, with the activation
      pointer shifted by (a*Wp + b) rows (tap block centre (a, b)); the zero border of the embedded gradient
      guarantees that no shifted read that matters wraps around an image row;
  dgrad of the 5x5 reconstruction layer = srk_conv_first_tc over the embedded gradient frame with
      SRK_PACK_FIRST_ROT180T weights (+ ReLU' mask); dgrad of the 1x1 layer = srk_conv_tc with SRK_PACK_DGRAD weights.
"""
from __future__ import annotations

from collections import OrderedDict
from types import SimpleNamespace

import numpy as np
import torch

from .. import ops
from ..initializers import srcnn_params
from ..params import ParamArena
from ..session import Handle

FLAGS = SimpleNamespace(ckpt_dir_path="./ckpts/", logs_dir_path="./logs/", training_images_path=None, sr_source_path=None,
                        sr_target_path=None, train=False, batch_size=64, upscaling_factor=3, crop_image_size=256, crop_image_side=6,
                        srcnn_fsub=33, srcnn_f1=9, srcnn_f2=1, srcnn_f3=5, srcnn_n1=64, srcnn_n2=32)


def sanity_check(flags=FLAGS):
    """srcnn/srcnn.py:28-43 (py2 `/` on ints is floor division)."""
    smaller_output_size = flags.srcnn_fsub - flags.srcnn_f1 - flags.srcnn_f2 - flags.srcnn_f3 + 3
    boundary = (flags.srcnn_fsub - smaller_output_size) // 2
    crop_size = (flags.crop_image_size - boundary * 2) // smaller_output_size
    flags.crop_image_side = boundary
    flags.crop_image_size = crop_size * smaller_output_size + boundary * 2
    if not flags.train:
        flags.batch_size = 1


def tap_block_centres(k: int):
    """Centres (a, b) of the 3x3 tap blocks that tile a k x k kernel (taps -k//2 .. k//2), row-major; the last block of an
    axis may reach past the kernel (k = 5: centres -1 and 2 cover taps -2..3, tap 3 is discarded)."""
    nb = -(-k // 3)
    c = [3 * i + 1 - k // 2 for i in range(nb)]
    return [(a, b) for a in c for b in c]


def assemble_tap_blocks(blocks: torch.Tensor, k: int) -> torch.Tensor:
    """[nb*nb, 9, ci, co] per-block 3x3 tap gradients (block order of `tap_block_centres`) -> the [k, k, ci, co] kernel."""
    nb = -(-k // 3)
    ci, co = blocks.shape[-2:]
    return blocks.reshape(nb, nb, 3, 3, ci, co).permute(0, 2, 1, 3, 4, 5).reshape(3 * nb, 3 * nb, ci, co)[:k, :k]


class SrcnnNet:
    NAMES = ("patch_extraction", "non_linear_mapping", "reconstruction")

    def __init__(self, params=None, channels=3, device="cuda", seed=0, flags=FLAGS):
        f = (flags.srcnn_f1, flags.srcnn_f2, flags.srcnn_f3)
        assert f == (9, 1, 5) and (flags.srcnn_n1, flags.srcnn_n2) == (64, 32), "kernels are instantiated for the 9-1-5 / 64-32 SRCNN"
        if params is None:
            params = srcnn_params(seed, channels, f, (64, 32))
        order = OrderedDict()
        for n in self.NAMES:
            order[f"{n}/weights:0"] = np.asarray(params[f"{n}/weights:0"], np.float32)
            order[f"{n}/biases:0"] = np.asarray(params[f"{n}/biases:0"], np.float32)
        self.C = order["patch_extraction/weights:0"].shape[2]
        self.device = device
        self.r = flags.upscaling_factor
        self.arena = ParamArena(order, device, decay_suffix=None)
        a = self.arena
        plan = ops.PackPlan(device)
        self._i1 = plan.add(a.offsets["patch_extraction/weights:0"], 9, self.C, 64, ops.PACK_FIRST)
        self._i2 = plan.add(a.offsets["non_linear_mapping/weights:0"], 1, 64, 32, ops.PACK_FWD, 32, 64)
        self._i3 = plan.add(a.offsets["reconstruction/weights:0"], 5, 32, self.C, ops.PACK_FWD, 16, 32)
        plan.finalize()
        self.plan = plan
        self.bias3 = torch.zeros(16, dtype=torch.float32, device=device)
        self.repack()
        self._panels = {}

    def repack(self):
        self.plan.run(self.arena.w)
        self.bias3[: self.C].copy_(self.arena.view("reconstruction/biases:0"))

    def degrade(self, hi: torch.Tensor) -> torch.Tensor:
        """In-graph bicubic down then up, TF1 legacy kernel (reference :89-93; sizes are py2 integer divisions)."""
        n, h, w, c = hi.shape
        lo = ops.resize_bicubic_tf1(hi, h // self.r, w // self.r)
        return ops.resize_bicubic_tf1(lo, h, w)

    def forward(self, lo: torch.Tensor) -> torch.Tensor:
        """lo fp32 [N,S,S,C] -> sr [N,S-12,S-12,C] (three VALID convolutions)."""
        n, H, W, C = lo.shape
        a = self.arena
        t1 = ops.conv_first_tc(lo, self.plan.views[self._i1], a.view("patch_extraction/biases:0"), 9, "VALID", "relu")
        t2 = ops.conv_tc(t1, self.plan.views[self._i2], a.view("non_linear_mapping/biases:0"), 1, "relu")
        h2, w2 = t2.H, t2.W
        key = (n, h2, w2)
        if key not in self._panels:  # VALID 5x5 == SAME over the FPA, keeping only outputs 2 px inside the border
            self._panels[key] = ops.make_panels([(i, -2, -2, 2, h2 - 2, 2, w2 - 2) for i in range(n)], self.device)
        return ops.conv_tc_last(t2, self.plan.views[self._i3], self.bias3, 5, C, "tanh", panels=self._panels[key],
                                frame_shape=(n, h2 - 4, w2 - 4))

    # ------------------------------------------------------------------------------------------ training
    def _enable_training(self, n, S):
        key = (n, S)
        if getattr(self, "_tb", None) is not None and self._tb["key"] == key:
            return self._tb
        a, dev, C = self.arena, self.device, self.C
        a.enable_training()
        H1, bb = S - 8, S - 12
        plan = ops.PackPlan(dev)
        ix = {
            "f1": plan.add(a.offsets["patch_extraction/weights:0"], 9, C, 64, ops.PACK_FIRST),
            "f2": plan.add(a.offsets["non_linear_mapping/weights:0"], 1, 64, 32, ops.PACK_FWD, 64, 64),
            "f3": plan.add(a.offsets["reconstruction/weights:0"], 5, 32, C, ops.PACK_FWD, 16, 64),
            "d3": plan.add(a.offsets["reconstruction/weights:0"], 5, 32, C, ops.PACK_FIRST_ROT180T),
            "d2": plan.add(a.offsets["non_linear_mapping/weights:0"], 1, 64, 32, ops.PACK_DGRAD, 64, 64),
        }
        plan.finalize()
        # geometry G0 = S x S (input), G1 = H1 x H1 (both hidden layers)
        Wp0 = S + 1
        margin = ((4 * Wp0 + 8 + 7) // 8) * 8  # rows the shifted activation views may reach before / behind the buffer
        rows0 = ops.fpa_rows(n, S, S)
        lo_store = torch.zeros((margin + rows0 + margin, 64), dtype=torch.bfloat16, device=dev)
        Wp1 = H1 + 1
        margin1 = ((3 * Wp1 + 8 + 7) // 8) * 8
        rows1 = ops.fpa_rows(n, H1, H1)
        t2_store = torch.zeros((margin1 + rows1 + margin1, 64), dtype=torch.bfloat16, device=dev)
        st0 = (ops.wgrad_workspace_bytes(n, S, S) + 1023) // 1024 * 1024
        st1 = (ops.wgrad_workspace_bytes(n, H1, H1) + 1023) // 1024 * 1024
        tmp1 = torch.zeros((9, 9, C, 64), dtype=torch.float32, device=dev)    # [block][tap][ci][co] of the 9x9 kernel
        tmp3 = torch.zeros((4, 9, 32, C), dtype=torch.float32, device=dev)    # [block][tap][ci][co] of the 5x5 kernel (6x6 taps)
        tmp2 = torch.zeros((9, 64, 32), dtype=torch.float32, device=dev)      # 1x1 kernel = centre tap
        dummy = torch.zeros(64, dtype=torch.float32, device=dev)
        g = lambda name: a.view(name, "g")  # noqa: E731
        dsts0 = [(tmp1[blk], g("patch_extraction/biases:0") if blk == 4 else dummy, C, 64) for blk in range(9)]
        dsts1 = [(tmp3[blk], g("reconstruction/biases:0") if blk == 0 else dummy, 32, C) for blk in range(4)]
        dsts1.append((tmp2, g("non_linear_mapping/biases:0"), 64, 32))
        self._tb = {
            "key": key, "plan": plan, "idx": ix, "st0": st0, "st1": st1, "margin": margin, "margin1": margin1,
            "lo_store": lo_store, "loF": ops.Fpa(lo_store[margin:margin + rows0], n, S, S),
            "t1": ops.fpa_empty(n, H1, H1, 64, dev),
            "t2_store": t2_store, "t2": ops.Fpa(t2_store[margin1:margin1 + rows1], n, H1, H1),
            "sr": torch.empty((n, bb, bb, C), dtype=torch.float32, device=dev),
            "dpre": torch.empty((n, bb, bb, C), dtype=torch.float32, device=dev),
            "dpre_e": torch.zeros((n, H1, H1, C), dtype=torch.float32, device=dev),   # zero border of 2 px stays zero
            "dP3": ops.fpa_empty(n, H1, H1, 64, dev),
            "d2": ops.fpa_empty(n, H1, H1, 64, dev), "d1": ops.fpa_empty(n, H1, H1, 64, dev),
            "d1e": ops.Fpa(torch.zeros((rows0, 64), dtype=torch.bfloat16, device=dev), n, S, S),   # zero border of 4 px stays zero
            "loss": torch.zeros(1, dtype=torch.float32, device=dev),
            "bias2": torch.zeros(64, dtype=torch.float32, device=dev),
            "ws0": torch.empty(st0 * 9, dtype=torch.uint8, device=dev), "ws1": torch.empty(st1 * 5, dtype=torch.uint8, device=dev),
            "tmp1": tmp1, "tmp2": tmp2, "tmp3": tmp3, "dummy": dummy,   # the dsts tables hold raw pointers: keep the targets alive
            "dsts0": ops.make_wgrad_dsts(dsts0, dev), "dsts1": ops.make_wgrad_dsts(dsts1, dev),
        }
        self.step = getattr(self, "step", 0)
        self._repack_train()
        return self._tb

    def _repack_train(self):
        b, a = self._tb, self.arena
        b["plan"].run(a.w)
        b["bias2"][:32].copy_(a.view("non_linear_mapping/biases:0"))

    @staticmethod
    def _shifted(store: torch.Tensor, margin: int, rows: int, shift: int, like: "ops.Fpa") -> "ops.Fpa":
        return ops.Fpa(store[margin + shift:margin + shift + rows], like.n_img, like.H, like.W)

    def forward_backward(self, hi: torch.Tensor):
        """hi fp32 [N,S,S,C] in [-1,1] -> fills the gradient arena; returns the buffer dict (loss, sr, lo)."""
        n, S, S2, C = hi.shape
        assert S == S2 and C == self.C and S + 1 <= 255
        b, a = self._enable_training(n, S), self.arena
        V, ix = b["plan"].views, b["idx"]
        H1, bb = S - 8, S - 12
        lo = self.degrade(hi)
        # ---- forward (64 channels wide: the padded halves of W2 / bias2 are zero, so t2[..., 32:] == 0)
        t1 = ops.conv_first_tc(lo, V[ix["f1"]], a.view("patch_extraction/biases:0"), 9, "VALID", "relu", out=b["t1"])
        t2 = ops.conv_tc(t1, V[ix["f2"]], b["bias2"], 1, "relu", out=b["t2"])
        pk = (n, H1)
        if pk not in self._panels:
            self._panels[pk] = ops.make_panels([(i, -2, -2, 2, H1 - 2, 2, H1 - 2) for i in range(n)], self.device)
        ops.conv_tc_last(t2, V[ix["f3"]], self.bias3, 5, C, "tanh", panels=self._panels[pk], frame_shape=(n, bb, bb), out=b["sr"])
        side = (S - bb) // 2
        hic = hi[:, side:side + bb, side:side + bb, :].contiguous()
        b["loss"].zero_()
        ops.l2norm_rows_mean_fwd_bwd(b["sr"], hic, bb * bb, b["loss"], b["dpre"], sr_act="tanh")
        # ---- backward: reconstruction (5x5) -> non_linear_mapping (1x1) -> patch_extraction (9x9)
        b["dpre_e"][:, 2:2 + bb, 2:2 + bb, :].copy_(b["dpre"])
        ops.nhwc_to_fpa_pad(b["dpre_e"], 64, out=b["dP3"])
        Wp1, rows1 = H1 + 1, b["t2"].data.shape[0]
        d2 = ops.conv_first_tc(b["dpre_e"], V[ix["d3"]], None, 5, "SAME", None, out=b["d2"], mask_src=t2, mask_kind="relu")
        d1 = ops.conv_tc(d2, V[ix["d2"]], None, 1, None, out=b["d1"], mask_src=t1, mask_kind="relu")
        # the 5x5 kernel's four 3x3 tap blocks (centres -1|2) and the 1x1 kernel: one batched wgrad + reduce launch on G1
        xs1 = [self._shifted(b["t2_store"], b["margin1"], rows1, ca * Wp1 + cb, b["t2"]) for (ca, cb) in tap_block_centres(5)] + [t1]
        ops.conv_wgrad_tc_batched(xs1, [b["dP3"]] * 4 + [d2], b["ws1"], b["st1"], b["dsts1"])
        # d1 lives on G1; embed it 4 px inside G0 next to the input frame
        src = d1.data[: n * (H1 + 1) * (H1 + 1)].view(n, H1 + 1, H1 + 1, 64)[:, 1:, :H1]
        b["d1e"].data[: n * (S + 1) * (S + 1)].view(n, S + 1, S + 1, 64)[:, 5:5 + H1, 4:4 + H1].copy_(src)
        ops.nhwc_to_fpa_pad(lo, 64, out=b["loF"])
        Wp0, rows0 = S + 1, b["loF"].data.shape[0]
        # the 9x9 kernel's nine 3x3 tap blocks (centres -3, 0, 3): one batched launch on G0
        xs0 = [self._shifted(b["lo_store"], b["margin"], rows0, ca * Wp0 + cb, b["loF"]) for (ca, cb) in tap_block_centres(9)]
        ops.conv_wgrad_tc_batched(xs0, [b["d1e"]] * 9, b["ws0"], b["st0"], b["dsts0"])
        # assemble the kernel gradients from their 3x3 tap blocks
        g = lambda name: a.view(name, "g")  # noqa: E731
        g("patch_extraction/weights:0").copy_(assemble_tap_blocks(b["tmp1"], 9))
        g("reconstruction/weights:0").copy_(assemble_tap_blocks(b["tmp3"], 5))
        g("non_linear_mapping/weights:0").copy_(b["tmp2"][4].view(1, 1, 64, 32))
        b["lo"] = lo
        return b

    def train_step(self, hi: torch.Tensor, learning_rate: float = 1e-3):
        """One Adam(lr, beta1=0.5, beta2=0.9) step (reference :155-157); returns the pre-update loss tensor."""
        b = self.forward_backward(hi)
        a = self.arena
        self.step += 1
        ops.adam_step(a.w, a.g, a.m, a.v, learning_rate, self.step, beta1=0.5, beta2=0.9)
        self._repack_train()
        self.repack()
        return b["loss"]

    def make_graphed_step(self, hi_static: torch.Tensor):
        """Capture the training step into two CUDA graphs (a 128-patch step is ~45 launches of a few microseconds each):
        graph A = degrade + forward + loss + backward, graph B = Adam (bias-corrected learning rate read from device memory,
        so one graph serves every step) + weight re-pack.  Returns `step(lr) -> loss tensor`; new batches are copied INTO
        `hi_static` before each call."""
        import math
        a = self.arena
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):  # warm-up outside capture: allocates buffers and panel tables, sets kernel attributes
            b = self.forward_backward(hi_static)
            lr_t = torch.zeros(1, dtype=torch.float32, device=self.device)
            ops.adam_step_dev(a.w, a.g, a.m, a.v, lr_t, beta1=0.5, beta2=0.9)  # lr_t == 0: no-op update
            self._repack_train()
            self.repack()
        torch.cuda.current_stream().wait_stream(side)
        a.m.zero_()
        a.v.zero_()
        torch.cuda.synchronize()
        g_fb, g_opt = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
        with torch.cuda.graph(g_fb):
            self.forward_backward(hi_static)
        with torch.cuda.graph(g_opt):
            ops.adam_step_dev(a.w, a.g, a.m, a.v, lr_t, beta1=0.5, beta2=0.9)
            self._repack_train()
            self.repack()
        lr_feed = ops.PinnedScalarFeed()

        def step(lr: float = 1e-3):
            g_fb.replay()
            self.step += 1
            lr_feed.push(lr * math.sqrt(1.0 - 0.9 ** self.step) / (1.0 - 0.5 ** self.step), lr_t)
            g_opt.replay()
            return b["loss"]

        step.graphs = (g_fb, g_opt)
        return step

    def loss(self, sr: torch.Tensor, hi: torch.Tensor):
        """mean over rows of ||reshape(sr - crop(hi), [-1, bb^2])||_2 (reference :132-144); returns (loss, dloss/dsr)."""
        bb = sr.shape[1]
        side = (hi.shape[1] - bb) // 2
        hic = hi[:, side:side + bb, side:side + bb, :].contiguous()
        acc = torch.zeros(1, device=self.device)
        dsr = torch.empty_like(sr)
        ops.l2norm_rows_mean_fwd_bwd(sr, hic, bb * bb, acc, dsr)
        return acc, dsr


class _SrcnnGraph:
    def __init__(self, net, hi_ph):
        self.net, self.hi_ph = net, hi_ph

    reader = None

    def execute(self, keys, feeds):
        net = self.net
        if keys == {"step"}:
            return {"step": net.step}
        x = feeds[self.hi_ph] if self.hi_ph is not None and self.hi_ph in feeds else next(self.reader)
        hi = (x if isinstance(x, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32))).to(net.device).contiguous()
        if "trainer" in keys:
            b = net.forward_backward(hi)
            a = net.arena
            net.step += 1
            ops.adam_step(a.w, a.g, a.m, a.v, 1e-3, net.step, beta1=0.5, beta2=0.9)   # reference :155-156: fixed hyper-parameters
            net._repack_train()
            net.repack()
            side = (hi.shape[1] - b["sr"].shape[1]) // 2
            crop = lambda t: t[:, side:side + b["sr"].shape[1], side:side + b["sr"].shape[1], :]  # noqa: E731
            return {"step": net.step, "trainer": None, "loss": float(b["loss"]), "sr_images": b["sr"].cpu().numpy(),
                    "hd_images": crop(hi).cpu().numpy(), "sd_images": crop(b["lo"]).cpu().numpy()}
        lo = net.degrade(hi)
        sr = net.forward(lo)
        side = (hi.shape[1] - sr.shape[1]) // 2
        crop = lambda t: t[:, side:side + sr.shape[1], side:side + sr.shape[1], :]
        out = {"step": net.step, "sr_images": sr.cpu().numpy(), "hd_images": crop(hi).cpu().numpy(), "sd_images": crop(lo).cpu().numpy()}
        if "loss" in keys:
            out["loss"] = float(net.loss(sr, hi)[0])
        return out


def build_dataset_reader(flags=FLAGS, seed=None):
    """srcnn/srcnn.py:46-78: an endless generator of [batch, crop, crop, 3] batches in [-1, 1] -- whole JPEG files read, randomly
    cropped and randomly mirrored -- from `training_images_path/*.jpg` (training) or the one `sr_source_path` image."""
    import glob
    import os
    from ..io.images import imread_u8
    paths = glob.glob(os.path.join(flags.training_images_path, "*.jpg")) if flags.train else [flags.sr_source_path]
    rng = np.random.RandomState(seed)
    size = flags.crop_image_size
    images = {}

    def crops():
        while True:
            for p_ in rng.permutation(paths):  # tf.train.string_input_producer: shuffled epochs
                if p_ not in images:
                    images[p_] = imread_u8(p_)
                im = images[p_]
                y, x = rng.randint(0, im.shape[0] - size + 1), rng.randint(0, im.shape[1] - size + 1)
                c = im[y:y + size, x:x + size]
                if rng.randint(2):
                    c = c[:, ::-1]
                yield c.astype(np.float32) / 127.5 - 1.0

    g = crops()
    while True:
        yield np.stack([next(g) for _ in range(flags.batch_size)])


def build_srcnn(hi_images=None, params=None, channels=3, device="cuda", seed=0, flags=FLAGS):
    """srcnn/srcnn.py:81 `build_srcnn()`: called without arguments it reads the global FLAGS and owns its dataset reader, as the
    reference does (fetches need no feeds); `hi_images` (a placeholder) replaces the in-graph reader (:86) with a feed."""
    net = SrcnnNet(params, channels, device, seed, flags)
    g = _SrcnnGraph(net, hi_images)
    if hi_images is None and (flags.training_images_path or flags.sr_source_path):
        g.reader = build_dataset_reader(flags)
    return {k: Handle(g, k) for k in ("step", "loss", "trainer", "hd_images", "sd_images", "sr_images")}


def build_sr_result(fetched):
    """srcnn/srcnn.py:168-183: [hd | sd | sr] side by side, the batch stacked vertically (numpy, from fetched arrays)."""
    hd, sd, sr = fetched["hd_images"], fetched["sd_images"], fetched["sr_images"]
    b, w = hd.shape[0], hd.shape[1]
    return np.concatenate([hd.reshape(1, b * w, w, -1), sd.reshape(1, b * w, w, -1), sr.reshape(1, b * w, w, -1)], axis=2)


def train(flags=FLAGS):
    """srcnn/srcnn.py:203-258."""
    import glob
    import json
    import os
    from ..params import load_params
    from ..session import Session
    os.makedirs(flags.ckpt_dir_path, exist_ok=True)
    os.makedirs(flags.logs_dir_path, exist_ok=True)
    found = glob.glob(os.path.join(flags.ckpt_dir_path, "model.ckpt-*.npz"))
    source = max(found, key=lambda p_: int(p_.rsplit("-", 1)[1][:-4])) if found else None
    srcnn = build_srcnn(params=load_params(source) if source else None, flags=flags)
    net = srcnn["step"].graph.net
    if source:
        net.step = int(source.rsplit("-", 1)[1][:-4])
    log = open(os.path.join(flags.logs_dir_path, "events.jsonl"), "a")
    stop = getattr(flags, "stop_training_at_k_step", None)  # (extension: the reference loops until interrupted)
    with Session() as session:
        while True:
            fetched = session.run({"loss": srcnn["loss"], "step": srcnn["step"], "trainer": srcnn["trainer"]})
            step = fetched["step"]
            log.write(json.dumps({"step": int(step), "loss": float(fetched["loss"])}) + "\n")
            if step % 100 == 0:
                print("loss[{}]: {}".format(step, fetched["loss"]))
            if step % 5000 == 0 or step == stop:
                net.arena.save(os.path.join(flags.ckpt_dir_path, f"model.ckpt-{step}.npz"), global_step=step)
            if step == stop:
                break
    log.close()


def super_resolution(flags=FLAGS):
    """srcnn/srcnn.py:261-288: [hd | sd | sr] of one crop of `sr_source_path`, saturate_cast((x + 1) * 127.5), written as an image."""
    import glob
    import os
    from ..io.images import write_png
    from ..params import load_params
    from ..session import Session
    found = glob.glob(os.path.join(flags.ckpt_dir_path, "model.ckpt-*.npz"))
    source = max(found, key=lambda p_: int(p_.rsplit("-", 1)[1][:-4]))
    srcnn = build_srcnn(params=load_params(source), flags=flags)
    with Session() as session:
        fetched = session.run({k: srcnn[k] for k in ("hd_images", "sd_images", "sr_images")})
    image = build_sr_result(fetched)[0]
    write_png(flags.sr_target_path, np.clip((image + 1.0) * 127.5, 0, 255).astype(np.uint8))


def main(_):
    """srcnn/srcnn.py:291-299."""
    sanity_check()
    if FLAGS.train:
        train()
    else:
        super_resolution()


if __name__ == "__main__":
    from .. import flags as _flags
    for _name, _default in vars(FLAGS).items():
        _kind = {bool: _flags.DEFINE_boolean, int: _flags.DEFINE_integer}.get(type(_default), _flags.DEFINE_string)
        _kind(_name, _default, "")
    _flags.DEFINE_integer("stop_training_at_k_step", 0, "stop after k steps (0: run until interrupted, as the reference does)")
    _argv = []
    for _a in __import__("sys").argv[1:]:  # the reference spells its flags with hyphens (--ckpt-dir-path): same names here
        if _a.startswith("--"):
            _k, _eq, _v = _a[2:].partition("=")
            _a = "--" + _k.replace("-", "_") + _eq + _v
        _argv.append(_a)
    _flags.parse(_argv)
    for _name in list(vars(FLAGS)) + ["stop_training_at_k_step"]:
        setattr(FLAGS, _name, getattr(_flags.FLAGS, _name))
    if not FLAGS.stop_training_at_k_step:
        FLAGS.stop_training_at_k_step = None
    main(None)
