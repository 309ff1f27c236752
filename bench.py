#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native super-resolution conv hot path.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload espcn|vdsr_train|vdsr_infer] [--impl reference]

Headline workload (BASELINE.json configs[1]): ESPCN 3x inference on synthetic 1920x1080 Y frames; metric
"output Mpix/s".  One step = one batch of FRAMES_PER_STEP frames through f1 -> f2 -> f3(+pixel shuffle).
`--workload vdsr_train` (configs[2]: VDSR-20 training, 64 patches of 41x41 per GPU, Adam, NCCL gradient
all-reduce) and `--workload vdsr_infer` (configs[3]: VDSR 4K frame, tiles sharded over the GPUs) are the two
other halves of BASELINE.json's metric; the default run appends their numbers under "also".
Prints ONE JSON line on rank 0.  `--impl reference` times the CPU restatement of the reference graph
(oracle/, torch-CPU fp32, all host threads) on the same workload instead.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FRAMES_PER_STEP = 4
LR_H, LR_W, SCALE = 1080, 1920, 3
VDSR_LAYERS = 20
TRAIN_BATCH, TRAIN_PATCH = 64, 41
FRAME_4K = (2160, 3840)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return dict(hbm=float(d["hbm_gbs"]), tf_burst=float(d["bf16_tflops"]), tf_sust=float(d["bf16_tflops_sustained"]), src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback")


def tensor_roofline(kernel: str, tflops: float, timed_ms: float, digits: int = 4) -> dict:
    """Roofline entry of a tensor-bound workload.  The denominator follows the length of the timed region: a region shorter than
    two seconds runs at the burst clock, so it is quoted against the BURST bf16 peak of MEASURED_PEAKS.json (the sustained one
    would flatter it); `frac_sustained` is given beside it for reference."""
    pk = peaks()
    burst = timed_ms < 2000.0
    peak = pk["tf_burst"] if burst else pk["tf_sust"]
    return {"bound": "tensor", "kernel": kernel, "achieved": round(tflops, 2 if tflops < 100 else 1), "peak": peak, "peak_kind": "burst" if burst else "sustained",
            "unit": "TFLOP/s", "frac": round(tflops / peak, digits), "frac_sustained": round(tflops / pk["tf_sust"], digits), "traffic": None,
            "peak_source": pk["src"], "timed_region_ms": round(timed_ms, 1)}


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index: int):
        self.idx, self.rows, self.proc = device_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.idx}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ helpers
def dist_info():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def timed_steps(step_fn, steps: int, warmup: int, world: int, sampler: ClockSampler | None, finalize=None):
    """W warm-up steps, then exactly K steps between barrier+synchronize; device time via CUDA events on the
    launching (current) stream; MAX over ranks."""
    for _ in range(warmup):
        step_fn()
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize()
    if sampler:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step_fn()
    if finalize:
        finalize()  # e.g. make the timing stream wait for side-stream copies of the last steps
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize()
    clocks = sampler.stop() if sampler else None
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device="cuda")
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        ms = float(t)
    return ms, clocks


def event_time(fn, iters=5):
    """Average device time (ms) of one callable, CUDA events on the current stream, after one warm call."""
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def launches_of(fn) -> int:
    from ml_super_resolution_b200 import _ffi
    before = _ffi.launch_count
    fn()
    return _ffi.launch_count - before


def train_e2e_leg(static_tensors, host_tensors, run_step, loss_dev, n_steps, world):
    """End-to-end leg of a training workload: every step copies ITS batch host -> device from pinned memory (session.DeviceFeed:
    the copy of batch k+1 overlaps step k) and reads ITS loss back (the host waits for the loss of the previous step while the
    current one runs -- two pinned slots -- the way a training loop logs its metric without stalling the queue)."""
    from ml_super_resolution_b200.session import DeviceFeed
    feed = DeviceFeed(static_tensors)
    slots = [torch.zeros_like(loss_dev, device="cpu").pin_memory() for _ in range(2)]
    evs = [torch.cuda.Event() for _ in range(2)]
    loss_h = torch.zeros_like(slots[0])
    i_ = [0]
    feed.put(host_tensors)

    def e2e_step():
        i = i_[0]
        i_[0] += 1
        feed.take()
        feed.put(host_tensors)  # the next step's batch starts copying now
        run_step()
        slots[i & 1].copy_(loss_dev, non_blocking=True)
        evs[i & 1].record()
        if i > 0:
            evs[(i - 1) & 1].synchronize()
            loss_h.copy_(slots[(i - 1) & 1])

    def finalize():
        evs[(i_[0] - 1) & 1].synchronize()
        feed.take()  # (drain the one batch that was prefetched beyond the last step)

    ms_e, _ = timed_steps(e2e_step, n_steps, 2, world, None, finalize=finalize)
    return ms_e


# ------------------------------------------------------------------------------------------------ ESPCN (headline)
ESPCN_FLOP_PER_OUT_PIXEL = 5027.6  # SURVEY 8(d) cfg2, C = 1: 2 * (25*64 + 576*32 + 288*9) / 9


def ncu_traffic(kernel_substr: str, shape_key: str):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel from the tracked ncu capture
    (profiles/r2_ncu_traffic.json, written by tools/ncu_traffic.py from an `ncu --set full` report); None unless the kernel
    name and the workload shape recorded there match what is being benchmarked."""
    p = os.path.join(ROOT, "profiles", "r2_ncu_traffic.json")
    if not os.path.exists(p):
        return None
    with open(p) as f:
        for e in json.load(f):
            if kernel_substr in e["kernel"] and e.get("shape") == shape_key:
                return e["dram_bytes"]
    return None


def espcn_workload(args, rank, world):
    from ml_super_resolution_b200 import ops
    from ml_super_resolution_b200.espcn.model_espcn import EspcnNet, build_model
    from ml_super_resolution_b200.session import Session, pinned_empty, placeholder
    from ml_super_resolution_b200.tiling import plan_tiles
    C = 1
    lr_ph = placeholder([None, None, None, C], "lr_source")
    model = build_model(lr_ph, SCALE, channels=C, seed=42)
    net = model["sr_result"].graph.net
    # trained-like magnitudes so tanh is exercised (reference init is sigma=0.02)
    net.arena.w.mul_(5.0)
    net.repack()
    g = torch.Generator(device="cuda").manual_seed(1235 + rank)
    lr = torch.rand((FRAMES_PER_STEP, LR_H, LR_W, C), device="cuda", generator=g) * 2 - 1
    out = torch.empty((FRAMES_PER_STEP, LR_H * SCALE, LR_W * SCALE, C), device="cuda")
    out_pix = FRAMES_PER_STEP * LR_H * SCALE * LR_W * SCALE

    def step():  # ONE launch: srk_espcn_forward (f1 -> f2 -> f3 -> depth_to_space, activations in tensor memory)
        net.forward(lr, shuffle=True, out=out)

    ms, clocks = timed_steps(step, args.steps, args.warmup, world, ClockSampler(torch.cuda.current_device()) if rank == 0 else None)
    value = out_pix * world * args.steps / ms / 1e3  # Mpix/s, whole job
    n_launch = launches_of(step) * args.steps

    # ---- roofline of the dominant (only) kernel: average launch duration over the timed region, live CUDA events above
    pk = peaks()
    k_ms = ms / args.steps
    flops = ESPCN_FLOP_PER_OUT_PIXEL * out_pix
    tf = flops / k_ms / 1e9
    burst = ms < 2000.0  # a kernel timed alone for less than ~2 s runs at boost clocks: quote it against the burst figure
    peak_tf = pk["tf_burst"] if burst else pk["tf_sust"]
    alg_bytes = FRAMES_PER_STEP * LR_H * LR_W * (4 * C + 4 * C * SCALE * SCALE)
    roofline = {"bound": "tensor", "kernel": "espcn_fused_kernel<1,3,true> (f1+f2+f3+pixel shuffle, one launch per step)", "achieved": round(tf, 1),
                "peak": peak_tf, "unit": "TFLOP/s", "frac": round(tf / peak_tf, 4), "peak_kind": "burst" if burst else "sustained",
                "traffic": ncu_traffic("espcn_fused_kernel", f"{FRAMES_PER_STEP}x{LR_H}x{LR_W}x{C} r{SCALE} f32"), "peak_source": pk["src"],
                "algorithmic_flops": flops, "algorithmic_bytes": alg_bytes, "hbm_GBps": round(alg_bytes / k_ms / 1e6, 1),
                "hbm_frac": round(alg_bytes / k_ms / 1e6 / pk["hbm"], 4),
                "note": "bound by the UMMA operand fetch from shared memory (SS-mode N <= 96 instructions run at a 60-cycle floor; ncu l1tex__data_pipe_tc_wavefronts_mem_shared 75 %) plus the hand-over of the single-buffered rotating accumulator window; DESIGN.md 3.6"}

    # ---- the layer-by-layer path (three kernels through HBM, what round 1 shipped and what training uses), for comparison
    Ht, Wt, tiles = plan_tiles(FRAMES_PER_STEP, LR_H, LR_W, 4)
    panels = ops.make_panels([t.as_tuple() for t in tiles])
    t1, t2 = net._get_bufs(len(tiles), Ht, Wt)
    a = net.arena
    k_f1 = event_time(lambda: ops.conv_first_tc(lr, net.plan.views[net._i1], a.view("f1/bias:0"), 5, "SAME", "tanh", panels=panels, panel_hw=(Ht, Wt), out=t1))
    k_f2 = event_time(lambda: ops.conv_tc(t1, net.plan.views[net._i2], a.view("f2/bias:0"), 3, "tanh", out=t2))
    k_f3 = event_time(lambda: ops.conv_tc_last(t2, net.plan.views[net._i3], net.bias3, 3, net.cout3, None, shuffle_r=SCALE, panels=panels,
                                               frame_shape=(FRAMES_PER_STEP, LR_H, LR_W), out=out))
    lr_px = FRAMES_PER_STEP * LR_H * LR_W
    layered = {"f1_ms": round(k_f1, 4), "f2_ms": round(k_f2, 4), "f3_ms": round(k_f3, 4),
               "output_Mpix_per_s": round(out_pix / (k_f1 + k_f2 + k_f3) / 1e3, 1),
               "GBps": {"f1": round(lr_px * (4 * C + 128) / k_f1 / 1e6, 1), "f2": round(lr_px * 192 / k_f2 / 1e6, 1),
                        "f3": round(lr_px * (64 + 4 * C * SCALE * SCALE) / k_f3 / 1e6, 1)},
               "step_algorithmic_GB": round(lr_px * (4 * C + 128 + 192 + 64 + 4 * C * SCALE * SCALE) / 1e9, 3)}
    roofline["layered_path"] = layered

    # ---- end to end THROUGH THE DROP-IN SEAM: session.run(model[...], feed_dict={lr_source: host frames}) with page-locked
    # host arrays.  Every step moves its input host->device and its result device->host inside the timed region.  The
    # headline form returns what the reference's test driver writes to disk -- uint8 = saturate_cast(sr*127.5+127.5)
    # (espcn/espcn/experiment_test.py:179-184) -- and the fp32 form (the raw session.run fetch) is reported next to it.
    lr_host = pinned_empty(lr.shape)
    lr_host[...] = lr.cpu().numpy()
    out_u8, out_f32 = pinned_empty(out.shape, "uint8"), pinned_empty(out.shape, "float32")
    sess = Session()

    lr_raw = pinned_empty(lr.shape, "uint8")  # the image as the driver reads it from disk (experiment_test.py:154-159), before `/ 127.5 - 1.0`
    lr_raw[...] = np.clip(np.rint((lr_host + 1.0) * 127.5), 0, 255).astype(np.uint8)

    def e2e_raw():  # uint8 frames in (normalised on the device, bit-identical to the host arithmetic), uint8 frames out
        sess.run(model["hr_images_u8"], feed_dict={model["lr_source_u8"]: lr_raw}, out={model["hr_images_u8"]: out_u8})

    def e2e_u8():
        sess.run(model["hr_images_u8"], feed_dict={lr_ph: lr_host}, out={model["hr_images_u8"]: out_u8})

    def e2e_f32():
        sess.run(model["hr_images"], feed_dict={lr_ph: lr_host}, out={model["hr_images"]: out_f32})

    n_e2e = max(4, args.steps // 2)
    ms_raw, _ = timed_steps(e2e_raw, n_e2e, 2, world, None)
    ms_u8, _ = timed_steps(e2e_u8, n_e2e, 2, world, None)
    ms_f32, _ = timed_steps(e2e_f32, max(2, n_e2e // 4), 1, world, None)
    e2e = {"value": round(out_pix * world * n_e2e / ms_raw / 1e3, 1), "unit": "output Mpix/s", "h2d_bytes_per_step": lr_raw.size,
           "d2h_bytes_per_step": out_u8.size,
           "form": "uint8 frames in (the decoded image, normalised on the device as experiment_test.py:159 does on the host), uint8 frames out "
                   "(saturate_cast, the PNG hand-off of :179-184)",
           "fp32_in_u8_out_form": {"value": round(out_pix * world * n_e2e / ms_u8 / 1e3, 1), "h2d_bytes_per_step": lr_host.size * 4, "d2h_bytes_per_step": out_u8.size},
           "fp32_form": {"value": round(out_pix * world * max(2, n_e2e // 4) / ms_f32 / 1e3, 1), "h2d_bytes_per_step": lr_host.size * 4, "d2h_bytes_per_step": out_f32.size * 4},
           "note": "Session.run on pinned host arrays: per call, row bands of each frame pipeline H2D / fused kernel / D2H on three streams; the call returns when the last band is on the host"}
    cfg = {"workload": f"ESPCN 3x (5x5-64 tanh, 3x3-32 tanh, 3x3-9 + pixel shuffle) inference, {FRAMES_PER_STEP} synthetic 1920x1080 Y frames/step/GPU, ONE fused kernel per step",
           "frames_per_step_per_gpu": FRAMES_PER_STEP, "lr_shape": [LR_H, LR_W, C], "scale": SCALE,
           "l2_policy": "input + output per step (33 MB + 299 MB) > 126 MB L2; the activations never leave the SM", "parallelism": f"frames x{world}"}
    return dict(metric="ESPCN 3x output Mpix/s (fwd)", value=round(value, 1), unit="output Mpix/s", ms=ms, clocks=clocks, roofline=roofline,
                e2e=e2e, gpu_launches=n_launch, config=cfg, scaling="weak")


def espcn_cpu(steps: int, threads: int):
    """CPU restatement (oracle, torch-CPU fp32 oneDNN) of the same ESPCN forward + pixel shuffle: `steps` steps of
    FRAMES_PER_STEP 1920x1080 Y frames each, the GPU arm's step."""
    from oracle import models as OM
    from oracle import ops as O
    torch.set_num_threads(threads)
    p = OM.espcn_init(seed=42, scaling_factor=SCALE, channels=1)
    lr = OM.synthetic_images(1235, FRAMES_PER_STEP, LR_H, LR_W, 1)
    OM.espcn_forward(p, lr[:1, :64, :64], dtype=np.float32)  # warm
    t0 = time.perf_counter()
    for _ in range(steps):
        y = OM.espcn_forward(p, lr, dtype=np.float32)
        O.pixel_shuffle(y, SCALE)
    dt = time.perf_counter() - t0
    return steps * FRAMES_PER_STEP * LR_H * SCALE * LR_W * SCALE / dt / 1e6, dt


# ------------------------------------------------------------------------------------------------ VDSR training
def vdsr_train_workload(args, rank, world):
    from ml_super_resolution_b200.vdsr.model_vdsr import VdsrNet
    net = VdsrNet(None, VDSR_LAYERS, 3, seed=42)
    g = torch.Generator(device="cuda").manual_seed(1236 + rank)
    hd = torch.rand((TRAIN_BATCH, TRAIN_PATCH, TRAIN_PATCH, 3), device="cuda", generator=g) * 2 - 1
    from ml_super_resolution_b200 import ops
    scales = torch.tensor([2.0, 3.0, 4.0], device="cuda")[torch.arange(TRAIN_BATCH, device="cuda") % 3]
    sd = ((ops.degrade_gauss_bilinear(hd * 0.5 + 0.5, scales)) * 2 - 1).contiguous()

    gstep = net.make_graphed_step(sd, hd)  # CUDA-graph replay of the identical kernel sequence (fwd+loss+bwd | all-reduce | Adam+repack)

    def step():
        gstep(5e-5)

    ms, clocks = timed_steps(step, args.steps, args.warmup, world, ClockSampler(torch.cuda.current_device()) if rank == 0 else None)
    value = TRAIN_BATCH * world * args.steps / ms * 1e3
    flops_per_patch = 6.7217e9
    tf = value * flops_per_patch / 1e12 / world
    roofline = tensor_roofline("whole training step (18x fwd/dgrad/wgrad tcgen05 convs dominate)", tf, ms)
    n_launch = launches_of(lambda: net.train_step(sd, hd, lr=5e-5, use_adam=True)) * args.steps  # same kernels the graphs replay
    sd_h, hd_h = sd.cpu().pin_memory(), hd.cpu().pin_memory()
    loss_buf = net._train_bufs["loss"]
    ms_e = train_e2e_leg([sd, hd], [sd_h, hd_h], step, loss_buf, max(2, args.steps // 2), world)
    e2e = {"value": round(TRAIN_BATCH * world * max(2, args.steps // 2) / ms_e * 1e3, 1), "unit": "patches/s",
           "h2d_bytes_per_step": 2 * sd_h.numel() * 4, "d2h_bytes_per_step": 4}
    # ---- the same step fed by the device-resident input pipeline (crop + flip + degrade from a uint8 image
    # pool in HBM, prefetched on a side stream) instead of a fixed batch; and the pipeline alone
    from ml_super_resolution_b200.vdsr import dataset as D
    rs = np.random.RandomState(7 + rank)
    pool = D.DevicePool([rs.randint(0, 256, (256, 256, 3)).astype(np.uint8) for _ in range(64)])
    gen = D.image_batches(pool, [2.0, 3.0, 4.0], TRAIN_PATCH, TRAIN_BATCH, rng=rs)

    def fed_step():
        s_, h_ = next(gen)
        sd.copy_(s_)
        hd.copy_(h_)
        gstep(5e-5)

    def pipe_step():
        next(gen)

    n_fed = max(5, args.steps // 2)
    ms_fed, _ = timed_steps(fed_step, n_fed, 3, world, None)
    ms_pipe, _ = timed_steps(pipe_step, n_fed, 3, world, None)
    pipeline = {"train_patches_per_s_fed_by_device_pipeline": round(TRAIN_BATCH * world * n_fed / ms_fed * 1e3, 1),
                "pipeline_alone_patches_per_s": round(TRAIN_BATCH * world * n_fed / ms_pipe * 1e3, 1),
                "pool": "64 synthetic 256x256x3 uint8 images resident in HBM; host draws the reference's random crops / flips / scales"}
    cfg = {"workload": f"VDSR-20 3x3-64 residual training, {TRAIN_BATCH} synthetic 41x41x3 patches/GPU, blur+bilinear 2/3/4x degrade, Adam",
           "input_pipeline": pipeline,
           "global_batch": TRAIN_BATCH * world, "parallelism": f"dp{world}", "l2_policy": "saved activations 19 x 14.5 MB = 275 MB per step > 126 MB L2"}
    return dict(metric="VDSR-20 training patches/s", value=round(value, 1), unit="patches/s", ms=ms, clocks=clocks, roofline=roofline, e2e=e2e,
                gpu_launches=n_launch, config=cfg, scaling="weak")


def vdsr_train_cpu(steps: int, threads: int):
    from oracle import models as OM
    torch.set_num_threads(threads)
    p = OM.vdsr_init(seed=42)
    sd = OM.synthetic_images(1, TRAIN_BATCH, TRAIN_PATCH, TRAIN_PATCH, 3)
    hd = OM.synthetic_images(2, TRAIN_BATCH, TRAIN_PATCH, TRAIN_PATCH, 3)
    t0 = time.perf_counter()
    for _ in range(steps):
        OM.vdsr_loss_and_grads(p, sd, hd, dtype=np.float32)
    dt = time.perf_counter() - t0
    return steps * TRAIN_BATCH / dt, dt


# ------------------------------------------------------------------------------------------------ VDSR 4K inference
def vdsr_infer_workload(args, rank, world):
    from ml_super_resolution_b200.vdsr.model_vdsr import VdsrNet
    net = VdsrNet(None, VDSR_LAYERS, 3, seed=42)
    H, W = FRAME_4K
    g = torch.Generator(device="cuda").manual_seed(1237)
    sd = torch.rand((1, H, W, 3), device="cuda", generator=g) * 2 - 1
    out = torch.empty_like(sd)

    # with several GPUs the frame is cut into a rank grid (2 x 4 regions at 8 GPUs), every region extended by the 20-px
    # receptive-field halo on its interior sides (tiling.rank_region); inside a rank, 242-px column panels swap seam columns
    tile_rows = None
    def step():
        net.forward(sd, out=out, rank=rank, world=world, tile_rows=tile_rows)

    ms, clocks = timed_steps(step, args.steps, args.warmup, world, ClockSampler(torch.cuda.current_device()) if rank == 0 else None)
    value = H * W * args.steps / ms / 1e3  # one frame per step for the whole job (strong scaling over tiles)
    tf = value * 1e6 * 1334016 / 1e12 / world
    roofline = tensor_roofline("conv_strip_kernel 3x3 64->64 (18 of 20 layers, column-strip form)", tf, ms)
    n_launch = launches_of(step) * args.steps
    sd_h = sd.cpu().pin_memory()
    out_h = torch.empty(sd.shape).pin_memory()
    # every rank moves only its region: the pixels it reads (host -> device) and the pixels it owns (device -> host), each as one
    # strided DMA transfer (ops.copy_region); one rank = the whole frame
    from ml_super_resolution_b200 import ops
    from ml_super_resolution_b200.tiling import rank_region
    own_box, read_box = rank_region(world, rank, H, W, VDSR_LAYERS)
    area = lambda bx: (bx[1] - bx[0]) * (bx[3] - bx[2])  # noqa: E731
    h2d = sum(area(rank_region(world, r, H, W, VDSR_LAYERS)[1]) for r in range(world)) * 3 * 4
    d2h = sum(area(rank_region(world, r, H, W, VDSR_LAYERS)[0]) for r in range(world)) * 3 * 4
    # throughput metric: frames are double-buffered, H2D / compute / D2H run on three streams so the PCIe transfers of
    # neighbouring frames overlap the 20 layers (every byte still moves every step)
    sd_dev = [torch.empty_like(sd) for _ in range(2)]
    out_dev = [out, torch.empty_like(out)]
    s_in, s_out, s_c = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.current_stream()
    ev_in = [torch.cuda.Event() for _ in range(2)]
    ev_c = [torch.cuda.Event() for _ in range(2)]
    ev_out = [torch.cuda.Event() for _ in range(2)]
    counter = [0]

    def e2e_step():
        b = counter[0] & 1
        counter[0] += 1
        with torch.cuda.stream(s_in):
            s_in.wait_event(ev_c[b])  # the forward that last read sd_dev[b] is done
            ops.copy_region(sd_dev[b], sd_h, read_box)
            ev_in[b].record(s_in)
        s_c.wait_event(ev_in[b])
        s_c.wait_event(ev_out[b])     # the D2H that last read out_dev[b] is done
        net.forward(sd_dev[b], out=out_dev[b], rank=rank, world=world, tile_rows=tile_rows)
        ev_c[b].record(s_c)
        with torch.cuda.stream(s_out):
            s_out.wait_event(ev_c[b])
            ops.copy_region(out_h, out_dev[b], own_box)
            ev_out[b].record(s_out)

    def e2e_finalize():
        s_c.wait_event(ev_out[0])
        s_c.wait_event(ev_out[1])

    ms_e, _ = timed_steps(e2e_step, max(2, args.steps // 2), 2, world, None, finalize=e2e_finalize)
    e2e = {"value": round(H * W * max(2, args.steps // 2) / ms_e / 1e3, 1), "unit": "output Mpix/s", "h2d_bytes_per_step": h2d,
           "d2h_bytes_per_step": d2h}
    cfg = {"workload": "VDSR-20 3x tiled inference, one synthetic 3840x2160x3 frame/step, 242-px column panels exchanging seam columns per layer; with several GPUs one region of a rank grid per GPU (2 x 4 at 8), 20-px halo on its interior sides",
           "parallelism": f"tiles x{world}", "l2_policy": "activations 2 x 1.26 GB ping-pong >> 126 MB L2"}
    return dict(metric="VDSR 3x output Mpix/s (fwd)", value=round(value, 1), unit="output Mpix/s", ms=ms, clocks=clocks, roofline=roofline, e2e=e2e,
                gpu_launches=n_launch, config=cfg, scaling="strong")


def vdsr_infer_cpu(threads: int, rows: int = 270):
    """Bounded sample: a 270-row x 3840 band of the 4K frame (1/8 of the pixels) through all 20 layers."""
    from oracle import models as OM
    torch.set_num_threads(threads)
    p = OM.vdsr_init(seed=42)
    sd = OM.synthetic_images(3, 1, rows, FRAME_4K[1], 3)
    t0 = time.perf_counter()
    OM.vdsr_forward_t(OM._to_t(p, np.float32), OM._t(sd, np.float32))
    dt = time.perf_counter() - t0
    return rows * FRAME_4K[1] / dt / 1e6, dt


# ------------------------------------------------------------------------------------------------ SRCNN training (BASELINE configs[0])
SRCNN_BATCH, SRCNN_PATCH = 128, 33


def srcnn_train_workload(args, rank, world):
    """SRCNN 9-1-5 forward+backward+Adam on synthetic 33x33 Y patches, batch 128 per GPU (replicas; the reference has no DP for it)."""
    from ml_super_resolution_b200.srcnn.srcnn import SrcnnNet, FLAGS
    net = SrcnnNet(None, 1, seed=42, flags=FLAGS)
    net.arena.w.mul_(60.0)  # trained-like magnitudes (reference init is sigma=0.001: every ReLU would be numerically dead)
    net.repack()
    g = torch.Generator(device="cuda").manual_seed(1238 + rank)
    hi = torch.rand((SRCNN_BATCH, SRCNN_PATCH, SRCNN_PATCH, 1), device="cuda", generator=g) * 2 - 1

    gstep = net.make_graphed_step(hi)  # CUDA-graph replay of the identical kernel sequence

    def step():
        gstep(1e-3)

    ms, clocks = timed_steps(step, args.steps, args.warmup, world, ClockSampler(torch.cuda.current_device()) if rank == 0 else None)
    value = SRCNN_BATCH * world * args.steps / ms * 1e3
    # 2*MACs: conv1 81*64 @25x25, conv2 64*32 @25x25, conv3 25*32 @21x21 per patch; backward = wgrad for all + dgrad for conv2/conv3
    fwd = 2 * (625 * 81 * 64 + 625 * 64 * 32 + 441 * 25 * 32)
    flops_per_patch = fwd * 3 - 2 * 625 * 81 * 64
    tf = value * flops_per_patch / 1e12 / world
    roofline = tensor_roofline("whole SRCNN step (launch-latency bound: 29 MFLOP/patch, ~45 launches)", tf, ms, digits=5)
    n_launch = launches_of(lambda: net.train_step(hi, 1e-3)) * args.steps  # the same kernels the graphs replay
    hi_h = hi.cpu().pin_memory()
    loss_h = torch.zeros(1).pin_memory()

    loss_dev = gstep(1e-3)
    ne = max(2, args.steps // 2)
    ms_e = train_e2e_leg([hi], [hi_h], lambda: gstep(1e-3), loss_dev, ne, world)
    e2e = {"value": round(SRCNN_BATCH * world * ne / ms_e * 1e3, 1), "unit": "patches/s", "h2d_bytes_per_step": hi_h.numel() * 4, "d2h_bytes_per_step": 4}
    cfg = {"workload": f"SRCNN 9-1-5 3x training (bicubic degrade, VALID convs, row-L2 loss, Adam), {SRCNN_BATCH} synthetic 33x33 Y patches/GPU",
           "global_batch": SRCNN_BATCH * world, "parallelism": f"replicas x{world}", "l2_policy": "working set ~40 MB < L2: a 3.7 GFLOP step is launch-bound, not memory-bound"}
    return dict(metric="SRCNN training patches/s", value=round(value, 1), unit="patches/s", ms=ms, clocks=clocks, roofline=roofline, e2e=e2e,
                gpu_launches=n_launch, config=cfg, scaling="weak")


def srcnn_train_cpu(steps: int, threads: int):
    from oracle import models as OM
    from oracle import ops as O
    torch.set_num_threads(threads)
    p = OM.srcnn_init(seed=42, channels=1)
    hi = OM.synthetic_images(4, SRCNN_BATCH, SRCNN_PATCH, SRCNN_PATCH, 1)
    t0 = time.perf_counter()
    for _ in range(steps):
        lo = O.resize_bicubic_tf1(O.resize_bicubic_tf1(hi, SRCNN_PATCH // 3, SRCNN_PATCH // 3), SRCNN_PATCH, SRCNN_PATCH)
        OM.srcnn_loss_and_grads(p, lo, hi, dtype=np.float32)
    dt = time.perf_counter() - t0
    return steps * SRCNN_BATCH / dt, dt


# ------------------------------------------------------------------------------------------------ EnhanceNet generator fwd+bwd (BASELINE configs[4])
ENET_BATCH, ENET_LR = 64, 32


def enet_train_workload(args, rank, world):
    """EnhanceNet generator forward + backward (+ Adam) on synthetic 32x32 -> 128x128 patches, batch 64 per GPU; the upstream
    gradient is the MSE term d(mean((sr-hd)^2))/d(sr) (the VGG / texture / adversarial terms are outside the hot path)."""
    from ml_super_resolution_b200 import ops
    from ml_super_resolution_b200.enet.model_enet import EnetGenerator
    net = EnetGenerator(None, seed=42)
    net.arena.w.mul_(2.5)
    net.repack()
    g = torch.Generator(device="cuda").manual_seed(1239 + rank)
    sd = torch.rand((ENET_BATCH, ENET_LR, ENET_LR, 3), device="cuda", generator=g) * 2 - 1
    bq = torch.rand((ENET_BATCH, 4 * ENET_LR, 4 * ENET_LR, 3), device="cuda", generator=g) * 2 - 1
    hd = torch.rand((ENET_BATCH, 4 * ENET_LR, 4 * ENET_LR, 3), device="cuda", generator=g) * 2 - 1
    dsr = torch.empty_like(hd)
    loss = torch.zeros(1, device="cuda")
    a = net.arena
    t = [0]

    def loss_head(sr):  # runs between the generator's forward and backward passes
        loss.zero_()
        ops.mse_fwd_bwd(sr, hd, loss, dsr)
        return dsr

    def eager_step():
        net.forward_backward(sd, bq, loss_head)
        if world > 1:
            ops.allreduce_grads(a.g)
        t[0] += 1
        ops.adam_step(a.w, a.g, a.m, a.v, 1e-4, t[0])
        net._tb["plan"].run(a.w)
        net.repack()

    net.forward_backward(sd, bq, hd)  # allocate the training buffers outside the timed region
    n_launch_step = launches_of(eager_step)
    gstep = net.make_graphed_step(sd, bq, loss_head)  # the whole step (incl. the NCCL all-reduce at N > 1) as one CUDA graph

    def step():
        gstep(1e-4)

    ms, clocks = timed_steps(step, args.steps, args.warmup, world, ClockSampler(torch.cuda.current_device()) if rank == 0 else None)
    value = ENET_BATCH * world * args.steps / ms * 1e3
    flops_per_patch = 3.617e9 * 3  # fwd + dgrad + wgrad, SURVEY 8 row a5: 3.617 GFLOP fwd/patch
    tf = value * flops_per_patch / 1e12 / world
    roofline = tensor_roofline("whole generator step (25 fwd + 24 dgrad + 25 wgrad tcgen05 convs; one CUDA graph)", tf, ms)
    n_launch = n_launch_step * args.steps
    sd_h, bq_h, hd_h = sd.cpu().pin_memory(), bq.cpu().pin_memory(), hd.cpu().pin_memory()
    loss_h = torch.zeros(1).pin_memory()

    ne = max(2, args.steps // 2)
    ms_e = train_e2e_leg([sd, bq, hd], [sd_h, bq_h, hd_h], step, loss, ne, world)
    e2e = {"value": round(ENET_BATCH * world * ne / ms_e * 1e3, 1), "unit": "patches/s",
           "h2d_bytes_per_step": (sd_h.numel() + bq_h.numel() + hd_h.numel()) * 4, "d2h_bytes_per_step": 4}
    cfg = {"workload": f"EnhanceNet 4x generator forward+backward+Adam (MSE upstream gradient), {ENET_BATCH} synthetic 32x32->128x128 patches/GPU",
           "global_batch": ENET_BATCH * world, "parallelism": f"dp{world}", "l2_policy": "saved activations ~1.3 GB per step >> 126 MB L2"}
    return dict(metric="EnhanceNet generator training patches/s", value=round(value, 1), unit="patches/s", ms=ms, clocks=clocks, roofline=roofline,
                e2e=e2e, gpu_launches=n_launch, config=cfg, scaling="weak")


def enet_train_cpu(steps: int, threads: int, batch: int = 4):
    from oracle import models as OM
    torch.set_num_threads(threads)
    p = OM.enet_g_init(seed=42)
    sd = OM.synthetic_images(5, batch, ENET_LR, ENET_LR, 3)
    bq = OM.synthetic_images(6, batch, 4 * ENET_LR, 4 * ENET_LR, 3)
    dsr = OM.synthetic_images(7, batch, 4 * ENET_LR, 4 * ENET_LR, 3) * 1e-4
    t0 = time.perf_counter()
    for _ in range(steps):
        OM.enet_generator_grads(p, sd, bq, dsr, dtype=np.float32)
    dt = time.perf_counter() - t0
    return steps * batch / dt, dt


WORKLOADS = {"enet_train": enet_train_workload, "espcn": espcn_workload, "vdsr_train": vdsr_train_workload, "vdsr_infer": vdsr_infer_workload, "srcnn_train": srcnn_train_workload}


def cpu_baseline_for(workload: str, threads: int):
    """The oracle port of `workload` timed on the host cores, on a bounded sample of the same workload (a few seconds each)."""
    with torch.no_grad():
        if workload == "espcn":
            v, dt = espcn_cpu(2, threads)
            return {"value": round(v, 2), "unit": "output Mpix/s", "cores": threads, "kind": "port", "sample": f"2 steps of {FRAMES_PER_STEP} 1920x1080 Y frames, {dt:.1f} s"}
        if workload == "vdsr_infer":
            v, dt = vdsr_infer_cpu(threads)
            return {"value": round(v, 3), "unit": "output Mpix/s", "cores": threads, "kind": "port", "sample": f"one 270x3840 band of the 4K frame through all 20 layers, {dt:.1f} s"}
    if workload == "enet_train":
        v, dt = enet_train_cpu(2, threads)
        return {"value": round(v, 2), "unit": "patches/s", "cores": threads, "kind": "port", "sample": f"2 fwd+bwd steps of 4 patches, {dt:.1f} s"}
    if workload == "srcnn_train":
        v, dt = srcnn_train_cpu(5, threads)
        return {"value": round(v, 2), "unit": "patches/s", "cores": threads, "kind": "port", "sample": f"5 degrade+fwd+bwd steps of 128 patches, {dt:.1f} s"}
    if workload == "vdsr_train":
        v, dt = vdsr_train_cpu(3, threads)
        cpu = {"value": round(v, 2), "unit": "patches/s", "cores": threads, "kind": "port", "sample": f"3 fwd+bwd steps of 64 41x41x3 patches (20 layers, no optimiser step), {dt:.1f} s"}
        from oracle import ops as O  # the reference's python input generator (numpy restatement), single-threaded as in the reference
        rs = np.random.RandomState(7)
        imgs = [rs.randint(0, 256, (256, 256, 3)).astype(np.uint8) for _ in range(8)]
        g_cpu = O.vdsr_image_batches(imgs, [2.0, 3.0, 4.0], TRAIN_PATCH, TRAIN_BATCH, rs)
        t0 = time.perf_counter()
        next(g_cpu)
        cpu["input_generator_patches_per_s"] = round(TRAIN_BATCH / (time.perf_counter() - t0), 1)
        return cpu
    return None


def dp_check(rank: int, world: int) -> dict:
    """Multi-GPU correctness, run ONCE before the timed region of every N > 1 run so that the scaling records carry it:
      grad_rel_err    data-parallel VDSR gradients after the NCCL all-reduce vs the single-GPU gradients of the concatenated batch
      replica_spread  max |w_rank - w_rank'| after 3 DP Adam steps (must be exactly 0)
      tiled_equal     VDSR tile-sharded frame (ranks own disjoint tiles) == the single-GPU frame, bit for bit
      espcn_bands_equal  fused ESPCN row bands computed by different ranks == the single-GPU frame, bit for bit"""
    import torch.distributed as dist
    from ml_super_resolution_b200.espcn.model_espcn import EspcnNet
    from ml_super_resolution_b200.initializers import vdsr_params
    from ml_super_resolution_b200.vdsr.model_vdsr import VdsrNet
    L, per = 6, 8
    rng = np.random.default_rng(7)
    params = vdsr_params(3, L, 3)
    for k in params:
        if k.endswith("bias:0"):
            params[k] = (0.05 * rng.standard_normal(params[k].shape)).astype(np.float32)
    sd_all = torch.from_numpy(rng.uniform(-1, 1, (per * world, 41, 41, 3)).astype(np.float32)).cuda()
    hd_all = torch.from_numpy(rng.uniform(-1, 1, (per * world, 41, 41, 3)).astype(np.float32)).cuda()
    sd, hd = sd_all[rank * per:(rank + 1) * per].contiguous(), hd_all[rank * per:(rank + 1) * per].contiguous()
    net = VdsrNet(params, L)
    net.forward_backward(sd, hd, numel_total=float(sd_all.numel()))
    dist.all_reduce(net.arena.g)
    ref = VdsrNet(params, L)
    ref.forward_backward(sd_all, hd_all)
    err = float((net.arena.g - ref.arena.g).norm() / ref.arena.g.norm())
    loss_dp = net._train_bufs["loss"][0:1].clone()
    dist.all_reduce(loss_dp)
    lerr = abs(float(loss_dp) - float(ref._train_bufs["loss"][0])) / float(ref._train_bufs["loss"][0])
    net2 = VdsrNet(params, L)
    for _ in range(3):
        net2.train_step(sd, hd, lr=1e-3)
    w = net2.arena.w.clone()
    wmax, wmin = w.clone(), w.clone()
    dist.all_reduce(wmax, op=dist.ReduceOp.MAX)
    dist.all_reduce(wmin, op=dist.ReduceOp.MIN)
    spread = float((wmax - wmin).abs().max())
    frame = torch.from_numpy(rng.uniform(-1, 1, (1, 200, 600, 3)).astype(np.float32)).cuda()
    out = torch.zeros_like(frame)
    net.forward(frame, out=out, tile_rows=80, rank=rank, world=world)
    dist.all_reduce(out)  # disjoint ownership: the sum assembles the frame (check-only collective)
    same = bool(torch.equal(out, ref.forward(frame, tile_rows=80)))
    en = EspcnNet(None, 3, 1, seed=5)
    en.arena.w.mul_(5.0)
    en.repack()
    lrf = torch.from_numpy(rng.uniform(-1, 1, (2, 90, 300, 1)).astype(np.float32)).cuda()
    eo = torch.zeros((2, 270, 900, 1), device="cuda")
    en.forward_fused(lrf, out=eo, rank=rank, world=world)
    dist.all_reduce(eo)
    esame = bool(torch.equal(eo, en.forward_fused(lrf)))
    # the graphed step with the exchange fused into Adam over NVLink peer memory (srk_allreduce_adam_step_dev) against the graphed
    # step with the NCCL all-reduce: same trajectories (the two sum the ranks in different orders), replicas bit-identical
    traj = {}
    for fused in (True, False):
        n3 = VdsrNet(params, L)
        gs = n3.make_graphed_step(sd.clone(), hd.clone(), peer_exchange=fused)
        for _ in range(4):
            gs(1e-3)
        torch.cuda.synchronize()
        traj[fused] = (n3.arena.w.clone(), gs.fused_exchange)
    wf = traj[True][0]
    fmax, fmin = wf.clone(), wf.clone()
    dist.all_reduce(fmax, op=dist.ReduceOp.MAX)
    dist.all_reduce(fmin, op=dist.ReduceOp.MIN)
    fspread = float((fmax - fmin).abs().max())
    fdiff = float((wf - traj[False][0]).abs().mean())
    res = {"world": world, "grad_rel_err": err, "loss_rel_err": lerr, "replica_spread": spread, "tiled_equal": same, "espcn_bands_equal": esame,
           "fused_exchange": bool(traj[True][1]), "fused_replica_spread": fspread, "fused_vs_nccl_mean_abs": fdiff}
    assert err < 2e-3 and lerr < 1e-4 and spread == 0.0 and same and esame and fspread == 0.0 and fdiff < 2e-5, res
    return res


def run_reference(args, rank, world, out):
    """The reference's CPU path: oracle restatement of the TF-1.8 graph on torch-CPU fp32 with all host threads."""
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    t_all = time.perf_counter()
    if args.workload == "espcn":
        steps = max(1, min(args.steps, 4))
        with torch.no_grad():
            for _ in range(min(args.warmup, 1)):
                espcn_cpu(1, threads)
            v, dt = espcn_cpu(steps, threads)
        unit, metric, sample = "output Mpix/s", "ESPCN 3x output Mpix/s (fwd)", f"{steps} step(s) of {FRAMES_PER_STEP} 1920x1080 Y frames each (the GPU arm's step)"
        cfg = {"workload": f"ESPCN 3x (5x5-64 tanh, 3x3-32 tanh, 3x3-9 + pixel shuffle) inference, {FRAMES_PER_STEP} synthetic 1920x1080 Y frames/step (CPU restatement, torch-CPU fp32)",
               "frames_per_step_per_gpu": FRAMES_PER_STEP, "lr_shape": [LR_H, LR_W, 1], "scale": SCALE}
    elif args.workload == "enet_train":
        steps = max(1, min(args.steps, 3))
        v, dt = enet_train_cpu(steps, threads)
        unit, metric, sample = "patches/s", "EnhanceNet generator training patches/s", f"{steps} fwd+bwd step(s) of 4 32x32->128x128 patches (GPU arm: 64/step, + MSE head + Adam)"
        cfg = {"workload": "EnhanceNet generator forward+backward, synthetic 32x32->128x128 patches (CPU restatement, torch-CPU fp32 autograd)"}
    elif args.workload == "srcnn_train":
        steps = max(1, min(args.steps, 10))
        v, dt = srcnn_train_cpu(steps, threads)
        unit, metric, sample = "patches/s", "SRCNN training patches/s", f"{steps} degrade+fwd+bwd step(s) of 128 33x33 Y patches (no optimiser step)"
        cfg = {"workload": "SRCNN 9-1-5 training, 128 synthetic 33x33 Y patches (CPU restatement, torch-CPU fp32 autograd)"}
    elif args.workload == "vdsr_train":
        steps = max(1, min(args.steps, 5))
        v, dt = vdsr_train_cpu(steps, threads)
        unit, metric, sample = "patches/s", "VDSR-20 training patches/s", f"{steps} fwd+bwd step(s) of 64 41x41x3 patches (no optimiser step)"
        cfg = {"workload": "VDSR-20 training, 64 synthetic 41x41x3 patches (CPU restatement, torch-CPU fp32 autograd)"}
    else:
        with torch.no_grad():
            v, dt = vdsr_infer_cpu(threads)
        unit, metric, steps, sample = "output Mpix/s", "VDSR 3x output Mpix/s (fwd)", 1, "one 270x3840 band (1/8 of a 4K frame)"
        cfg = {"workload": "VDSR-20 inference on a synthetic 4K frame (CPU restatement, torch-CPU fp32)"}
    line = {"impl": "reference", "metric": metric, "value": round(v, 3), "unit": unit, "n_gpus": world, "steps": steps, "warmup": min(args.warmup, 1),
            "ms_per_step": round(dt * 1e3 / steps, 2), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": cfg,
            "cpu_baseline": {"value": round(v, 3), "unit": unit, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": round(v, 3), "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "wall_s": round(time.perf_counter() - t_all, 1)}
    out.emit(json.dumps(line))


class _JsonOnlyStdout:
    """Keeps stdout for the ONE JSON line: while the benchmark runs, file descriptor 1 points at stderr, so anything a library
    prints to stdout (NCCL's version banner, for one) cannot end up next to the record."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def emit(self, line: str):
        sys.stdout.flush()
        os.write(self.saved, (line + "\n").encode())

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)
        return False


def main():
    with _JsonOnlyStdout() as out:
        _main(out)


def _main(out):
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="espcn", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--no-also", action="store_true", help="skip the secondary workloads in the default run")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "native" else args.warmup
    rank, world, local = dist_info()
    if args.impl == "reference":
        run_reference(args, rank, world, out)
        return
    assert torch.cuda.is_available(), "bench.py (native arm) needs a B200; there is no CPU fallback"
    if world != args.gpus and world == 1 and args.gpus > 1:
        raise SystemExit("launch N>1 with: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 bench.py --gpus N ...")
    torch.cuda.set_device(local)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local))
    dpc = dp_check(rank, world) if world > 1 else None
    res = WORKLOADS[args.workload](args, rank, world)
    also = {}
    if args.workload == "espcn" and not args.no_also:
        sub = argparse.Namespace(**vars(args))
        sub.steps, sub.warmup = max(5, args.steps // 2), 3
        for name in ("vdsr_train", "vdsr_infer", "srcnn_train", "enet_train"):
            r = WORKLOADS[name](sub, rank, world)
            also[name] = {"metric": r["metric"], "value": r["value"], "unit": r["unit"], "ms_per_step": round(r["ms"] / sub.steps, 4),
                          "roofline": r["roofline"], "e2e": r["e2e"], "scaling": r["scaling"], "steps": sub.steps, "config": r["config"]}
    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            cpu = cpu_baseline_for(args.workload, threads)
            for name in also:  # every part of BASELINE.json's metric carries its own CPU baseline in the same run
                also[name]["cpu_baseline"] = cpu_baseline_for(name, threads)
        line = {"metric": res["metric"], "value": res["value"], "unit": res["unit"], "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": round(res["ms"] / args.steps, 4), "higher_is_better": True, "scaling": res["scaling"], "vs_baseline": None,
                "dtype": "bf16", "data": "synthetic", "config": res["config"], "clocks": res["clocks"], "e2e": res["e2e"],
                "gpu_launches": res["gpu_launches"], "roofline": res["roofline"], "cpu_baseline": cpu}
        if dpc is not None:
            line["dp_check"] = dpc
        if also:
            line["also"] = also
        out.emit(json.dumps(line))
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
