"""Model-level parity (through the reference-shaped builders and the C ABI) against the CPU oracle.
north_star tolerances: conv outputs max-abs <= 2e-2 on [0,1] images in bf16 (= 4e-2 on the reference's
[-1,1] range), PSNR within 0.02 dB, pixel-shuffle indexing bit-exact, tiled == untiled."""
import numpy as np
import pytest
import torch

from oracle import models as OM
from oracle import ops as O

pytestmark = pytest.mark.gpu

TOL_BF16 = 4e-2  # max-abs on [-1,1] data


def _trained_like(params, scale=1.0, seed=5):
    """Reference initialisers leave biases at zero; add non-zero biases so every epilogue path is exercised."""
    rng = np.random.default_rng(seed)
    out = {}
    for k, v in params.items():
        if k.endswith(("bias:0", "biases:0")):
            out[k] = (0.05 * rng.standard_normal(v.shape)).astype(np.float32)
        else:
            out[k] = (v * scale).astype(np.float32)
    return out


def _psnr(a, b, max_val=2.0):
    return float(np.mean(O.psnr(a, b, max_val)))


def test_vdsr_forward_matches_oracle(srk_ops):
    from ml_super_resolution_b200.vdsr.model_vdsr import build_model
    from ml_super_resolution_b200.session import Session, placeholder
    params = _trained_like(OM.vdsr_init(seed=42))
    hd = OM.synthetic_images(1236, 4, 41, 41, 3)
    sd = np.stack([O.hd_image_to_sd_image(h * 0.5 + 0.5, 3) * 2 - 1 for h in hd]).astype(np.float32)
    sd_ph, hd_ph = placeholder([None, None, None, 3], "sd_images"), placeholder([None, None, None, 3], "hd_images")
    model = build_model(sd_ph, hd_ph, num_layers=20, use_adam=True, params=params)
    with Session() as s:
        got = s.run({"sr": model["sr_images"], "c7": model["conv.7"], "c20": model["conv.20"], "loss": model["loss"]},
                    feed_dict={sd_ph: sd, hd_ph: hd})
    ref = OM.vdsr_forward(params, sd)
    assert np.abs(got["sr"] - ref["sr_images"]).max() <= TOL_BF16
    assert np.abs(got["c7"] - ref["conv.7"]).max() <= TOL_BF16 * max(1.0, np.abs(ref["conv.7"]).max())
    assert np.abs(got["c20"] - ref["conv.20"]).max() <= TOL_BF16
    # PSNR of each result against the ground truth agrees within 0.02 dB
    assert abs(_psnr(got["sr"], hd) - _psnr(ref["sr_images"], hd)) <= 0.02
    ref_loss, _, _, _ = OM.vdsr_loss_and_grads(params, sd, hd)
    assert abs(got["loss"] - ref_loss) <= 2e-3 * abs(ref_loss) + 1e-5


def test_vdsr_tiled_equals_untiled_and_oracle(srk_ops):
    from ml_super_resolution_b200.vdsr.model_vdsr import VdsrNet
    params = _trained_like(OM.vdsr_init(seed=1, num_layers=8))
    net = VdsrNet(params, num_layers=8)
    sd = OM.synthetic_images(77, 1, 96, 200, 3)
    x = torch.from_numpy(sd).cuda()
    full = net.forward(x).cpu().numpy()
    tiled = net.forward(x, tile_rows=40).cpu().numpy()          # row tiles with an 8-px halo
    assert np.array_equal(full, tiled), "tiled inference must be bit-identical to the un-tiled frame"
    # two-rank tile sharding without a collective: each rank fills only the pixels it owns
    out = torch.full_like(x, float("nan"))
    net.forward(x, out=out, tile_rows=40, rank=0, world=2)
    net.forward(x, out=out, tile_rows=40, rank=1, world=2)
    assert np.array_equal(out.cpu().numpy(), full)
    # column panels that swap their seam columns after every layer (no recomputed halo): still bit-identical, alone, combined
    # with row bands, and sharded band-wise over two ranks
    assert np.array_equal(net.forward(x, max_panel_w=80).cpu().numpy(), full)
    assert np.array_equal(net.forward(x, max_panel_w=72, tile_rows=56).cpu().numpy(), full)
    out = torch.full_like(x, float("nan"))
    net.forward(x, out=out, max_panel_w=72, tile_rows=56, rank=0, world=2)
    net.forward(x, out=out, max_panel_w=72, tile_rows=56, rank=1, world=2)
    assert np.array_equal(out.cpu().numpy(), full)
    # a shard that splits a band falls back to receptive-field halos
    out = torch.full_like(x, float("nan"))
    for r in range(3):
        net.forward(x, out=out, max_panel_w=80, rank=r, world=3)
    assert np.array_equal(out.cpu().numpy(), full)
    # wide frame (column panels) against the oracle
    sdw = OM.synthetic_images(78, 1, 24, 600, 3)
    got = net.forward(torch.from_numpy(sdw).cuda()).cpu().numpy()
    ref = OM.vdsr_forward(params, sdw, num_layers=8)["sr_images"]
    assert np.abs(got - ref).max() <= TOL_BF16


def test_vdsr_train_step_matches_oracle(srk_ops):
    from ml_super_resolution_b200.vdsr.model_vdsr import VdsrNet
    L = 6
    params = _trained_like(OM.vdsr_init(seed=3, num_layers=L))
    net = VdsrNet(params, num_layers=L)
    hd = OM.synthetic_images(11, 8, 41, 41, 3)
    sd = np.ascontiguousarray(np.stack([O.hd_image_to_sd_image(h * 0.5 + 0.5, 2) * 2 - 1 for h in hd]), dtype=np.float32)
    b = net.forward_backward(torch.from_numpy(sd).cuda(), torch.from_numpy(hd).cuda())
    loss = float(b["loss"].sum())
    ref_loss, ref_mse, ref_g, _ = OM.vdsr_loss_and_grads(params, sd, hd, num_layers=L)
    assert abs(loss - ref_loss) <= 2e-3 * ref_loss
    got_g = net.arena.to_numpy("g")
    for k, g in ref_g.items():
        if k.endswith("kernel:0"):
            g = g - 1e-4 * params[k]  # the oracle's gradient includes the l2 term; ours adds it inside Adam
        rel = np.linalg.norm(got_g[k] - g) / (np.linalg.norm(g) + 1e-30)
        assert rel <= 3e-2, f"{k}: relative gradient error {rel:.4f}"
    # ten Adam steps: loss trajectory follows the fp64 oracle
    from oracle.ops import adam_tf
    p64 = {k: v.astype(np.float64) for k, v in params.items()}
    m = {k: np.zeros_like(v) for k, v in p64.items()}
    v_ = {k: np.zeros_like(v) for k, v in p64.items()}
    ours, theirs = [], []
    sdt, hdt = torch.from_numpy(sd).cuda(), torch.from_numpy(hd).cuda()
    for t in range(1, 11):
        ours.append(float(net.train_step(sdt, hdt, lr=1e-3, use_adam=True)))
        l, _, g, _ = OM.vdsr_loss_and_grads(p64, sd, hd, num_layers=L)
        theirs.append(l)
        for k in p64:
            p64[k], m[k], v_[k] = adam_tf(p64[k], g[k], m[k], v_[k], t, 1e-3, dtype=np.float64)
    assert np.allclose(ours, theirs, rtol=2e-2), (ours, theirs)
    assert ours[-1] < ours[0]


@pytest.mark.parametrize("channels,r,shape", [(3, 3, (2, 17, 17)), (1, 3, (1, 36, 64)), (3, 2, (1, 20, 300)), (3, 4, (1, 9, 11))])
def test_espcn_forward_matches_oracle(srk_ops, channels, r, shape):
    from ml_super_resolution_b200.espcn.model_espcn import EspcnNet
    n, h, w = shape
    params = _trained_like(OM.espcn_init(seed=9, scaling_factor=r, channels=channels), scale=5.0)
    net = EspcnNet(params, r, channels)
    lr = OM.synthetic_images(5, n, h, w, channels)
    x = torch.from_numpy(lr).cuda()
    packed = net.forward(x, shuffle=False).cpu().numpy()
    shuffled = net.forward(x, shuffle=True).cpu().numpy()
    ref = OM.espcn_forward(params, lr)
    assert np.abs(packed - ref).max() <= TOL_BF16
    # the fused depth_to_space is index-exact: shuffling our own packed output reproduces it bit for bit
    assert np.array_equal(shuffled, O.pixel_shuffle(packed, r))
    assert abs(_psnr(shuffled, O.pixel_shuffle(ref, r) + 0.1) - _psnr(O.pixel_shuffle(ref, r), O.pixel_shuffle(ref, r) + 0.1)) <= 0.02


def test_vdsr_graphed_step_equals_eager(srk_ops):
    """The CUDA-graph replay of the training step performs the same arithmetic as the eager step."""
    from ml_super_resolution_b200.vdsr.model_vdsr import VdsrNet
    L = 5
    params = _trained_like(OM.vdsr_init(seed=4, num_layers=L))
    hd = OM.synthetic_images(21, 8, 41, 41, 3)
    sd = OM.synthetic_images(22, 8, 41, 41, 3)
    sdt, hdt = torch.from_numpy(sd).cuda(), torch.from_numpy(hd).cuda()
    eager, graphed = VdsrNet(params, num_layers=L), VdsrNet(params, num_layers=L)
    gstep = graphed.make_graphed_step(sdt, hdt)
    for _ in range(3):
        le = float(eager.train_step(sdt, hdt, lr=1e-3))
        lg = float(gstep(1e-3).sum())
        assert le == pytest.approx(lg, rel=2e-5)  # the MSE partial sums meet in an fp32 atomic: order-dependent last bits
    we, wg = eager.arena.to_numpy(), graphed.arena.to_numpy()
    for k in we:
        # first/last-layer wgrad and the MSE sum use fp32 atomics (order-dependent last bits); Adam's first steps act like
        # sign(g), so a gradient element near zero may move the other way: bounded by 2*lr per step, and rare
        d = np.abs(we[k] - wg[k])
        assert d.max() <= 3 * 2e-3 and d.mean() <= 2e-5, (k, d.max(), d.mean())


@pytest.mark.parametrize("channels,S,batch", [(1, 33, 8), (3, 45, 2)])
def test_srcnn_forward_degrade_and_loss(srk_ops, channels, S, batch):
    from ml_super_resolution_b200.srcnn.srcnn import SrcnnNet
    params = _trained_like(OM.srcnn_init(seed=6, channels=channels), scale=60.0)  # sigma 0.001 init -> visible activations
    net = SrcnnNet(params, channels)
    hi = OM.synthetic_images(31, batch, S, S, channels)
    hit = torch.from_numpy(hi).cuda()
    lo = net.degrade(hit)
    ref_lo = O.resize_bicubic_tf1(O.resize_bicubic_tf1(hi, S // 3, S // 3), S, S)
    assert np.array_equal(lo.cpu().numpy(), ref_lo)  # TF1 bicubic: bit exact
    sr = net.forward(lo)
    ref_sr = OM.srcnn_forward(params, ref_lo)
    assert sr.shape == ref_sr.shape == (batch, S - 12, S - 12, channels)
    assert np.abs(sr.cpu().numpy() - ref_sr).max() <= TOL_BF16
    loss, dsr = net.loss(sr, hit)
    bb = S - 12
    ref_loss, ref_grad = O.l2norm_rows_mean(sr.cpu().numpy(), hi[:, 6:6 + bb, 6:6 + bb, :], bb * bb)
    assert abs(float(loss) - ref_loss) <= 1e-4 * ref_loss
    np.testing.assert_allclose(dsr.cpu().numpy(), ref_grad, rtol=1e-4, atol=1e-8)


def test_enet_generator_forward(srk_ops):
    from ml_super_resolution_b200.enet.model_enet import EnetGenerator
    params = _trained_like(OM.enet_g_init(seed=8), scale=2.5)
    net = EnetGenerator(params)
    sd = OM.synthetic_images(41, 2, 32, 32, 3)
    bq = OM.synthetic_images(42, 2, 128, 128, 3)
    got = net.forward(torch.from_numpy(sd).cuda(), torch.from_numpy(bq).cuda()).cpu().numpy()
    ref = OM.enet_generator_forward(params, sd, bq)
    res_scale = np.abs(ref - bq).max()
    assert res_scale > 0.05, "test weights too small to exercise the generator"
    assert np.abs(got - ref).max() <= TOL_BF16 * max(1.0, res_scale)
    assert abs(_psnr(got, bq + 0.3) - _psnr(ref, bq + 0.3)) <= 0.02


def test_espcn_train_step_matches_oracle(srk_ops):
    """ESPCN training (MSE in packed space, tanh layers, Adam): loss, every gradient and a 5-step trajectory vs the oracle."""
    from ml_super_resolution_b200.espcn.model_espcn import EspcnNet
    from oracle.ops import adam_tf
    params = _trained_like(OM.espcn_init(seed=12, scaling_factor=3, channels=3), scale=4.0)
    net = EspcnNet(params, 3, 3)
    lr = OM.synthetic_images(51, 16, 17, 17, 3)
    hr = O.pixel_unshuffle(OM.synthetic_images(52, 16, 51, 51, 3), 3)
    lrt, hrt = torch.from_numpy(lr).cuda(), torch.from_numpy(np.ascontiguousarray(hr)).cuda()
    b = net.forward_backward(lrt, hrt)
    ref_loss, ref_g, _ = OM.espcn_loss_and_grads(params, lr, hr)
    assert abs(float(b["loss"]) - ref_loss) <= 2e-3 * ref_loss
    got = net.arena.to_numpy("g")
    for k, g in ref_g.items():
        rel = np.linalg.norm(got[k] - g) / (np.linalg.norm(g) + 1e-30)
        assert rel <= 3e-2, f"{k}: relative gradient error {rel:.4f}"
    p64 = {k: v.astype(np.float64) for k, v in params.items()}
    m = {k: np.zeros_like(v) for k, v in p64.items()}
    v_ = {k: np.zeros_like(v) for k, v in p64.items()}
    ours, theirs = [], []
    for t in range(1, 6):
        ours.append(float(net.train_step(lrt, hrt, 1e-3)))
        l, g, _ = OM.espcn_loss_and_grads(p64, lr, hr)
        theirs.append(l)
        for k in p64:
            p64[k], m[k], v_[k] = adam_tf(p64[k], g[k], m[k], v_[k], t, 1e-3, dtype=np.float64)
    assert np.allclose(ours, theirs, rtol=2e-2), (ours, theirs)
    assert ours[-1] < ours[0]
    # inference weights were re-packed too: the forward path sees the trained parameters
    packed = net.forward(lrt, shuffle=False).cpu().numpy()
    assert np.abs(packed - OM.espcn_forward(p64, lr)).max() <= TOL_BF16


@pytest.mark.parametrize("channels,S,batch", [(1, 33, 16), (3, 45, 4)])
def test_srcnn_train_step_matches_oracle(srk_ops, channels, S, batch):
    """SRCNN training (9-1-5 VALID, row-L2-norm loss through tanh, Adam(1e-3,.5,.9)): loss, every gradient and a 4-step
    trajectory vs the oracle (srcnn/srcnn.py:100-157)."""
    from ml_super_resolution_b200.srcnn.srcnn import SrcnnNet
    from oracle.ops import adam_tf
    params = _trained_like(OM.srcnn_init(seed=6, channels=channels), scale=60.0)
    net = SrcnnNet(params, channels)
    hi = OM.synthetic_images(33, batch, S, S, channels)
    hit = torch.from_numpy(hi).cuda()
    lo = O.resize_bicubic_tf1(O.resize_bicubic_tf1(hi, S // 3, S // 3), S, S)
    b = net.forward_backward(hit)
    ref_loss, ref_g, ref_sr = OM.srcnn_loss_and_grads(params, lo, hi)
    assert np.abs(b["sr"].cpu().numpy() - ref_sr).max() <= TOL_BF16
    assert abs(float(b["loss"]) - ref_loss) <= 5e-3 * ref_loss
    got = net.arena.to_numpy("g")
    for k, g in ref_g.items():
        rel = np.linalg.norm(got[k] - g) / (np.linalg.norm(g) + 1e-30)
        assert rel <= 3e-2, f"{k}: relative gradient error {rel:.4f}"
    p64 = {k: v.astype(np.float64) for k, v in params.items()}
    m = {k: np.zeros_like(v) for k, v in p64.items()}
    v_ = {k: np.zeros_like(v) for k, v in p64.items()}
    ours, theirs = [], []
    for t in range(1, 5):
        ours.append(float(net.train_step(hit, 1e-3)))
        l, g, _ = OM.srcnn_loss_and_grads(p64, lo, hi)
        theirs.append(l)
        for k in p64:
            p64[k], m[k], v_[k] = adam_tf(p64[k], g[k], m[k], v_[k], t, 1e-3, beta1=0.5, beta2=0.9, dtype=np.float64)
    assert np.allclose(ours, theirs, rtol=2e-2), (ours, theirs)
    # the inference path sees the trained parameters
    sr = net.forward(net.degrade(hit)).cpu().numpy()
    assert np.abs(sr - OM.srcnn_forward(p64, lo)).max() <= TOL_BF16
    # CUDA-graph replay of the step follows the eager trajectory (same kernels; only atomics order may differ)
    net2 = SrcnnNet(params, channels)
    gstep = net2.make_graphed_step(hit.clone())
    graphed = [float(gstep(1e-3)) for _ in range(4)]
    assert np.allclose(graphed, ours, rtol=2e-3), (graphed, ours)


def test_enet_generator_backward_matches_oracle(srk_ops):
    """EnhanceNet generator backward for a supplied d(sr): every kernel / bias gradient of the 25 layers vs autograd of the
    oracle (enet/enet/model_enet.py:8-31,44-115,336-341; BASELINE cfg5 geometry 32x32 -> 128x128)."""
    from ml_super_resolution_b200.enet.model_enet import EnetGenerator
    params = _trained_like(OM.enet_g_init(seed=8), scale=2.5)
    net = EnetGenerator(params)
    sd = OM.synthetic_images(41, 2, 32, 32, 3)
    bq = OM.synthetic_images(42, 2, 128, 128, 3)
    # upstream gradient of an MSE term against a synthetic HD batch: d(mean((sr-hd)^2))/d(sr).  (An i.i.d.-noise d(sr) makes
    # every weight gradient an incoherent sum in which the ~1 % of ReLU masks that bf16 activations flip shows up as
    # ~sqrt(1 %) relative error; a real loss gradient is spatially coherent.)
    hd = OM.synthetic_images(43, 2, 128, 128, 3)
    dsr = (2.0 * (OM.enet_generator_forward(params, sd, bq) - hd) / hd.size).astype(np.float32)
    sr = net.forward_backward(torch.from_numpy(sd).cuda(), torch.from_numpy(bq).cuda(), torch.from_numpy(dsr).cuda()).cpu().numpy()
    ref_g, ref_sr = OM.enet_generator_grads(params, sd, bq, dsr)
    assert np.abs(sr - ref_sr).max() <= TOL_BF16 * max(1.0, np.abs(ref_sr - bq).max())
    got = net.arena.to_numpy("g")
    worst = 0.0
    for k, g in ref_g.items():
        rel = np.linalg.norm(got[k] - g) / (np.linalg.norm(g) + 1e-30)
        worst = max(worst, rel)
        assert rel <= 6e-2, f"{k}: relative gradient error {rel:.4f}"
    assert worst > 0.0


def test_enet_generator_tiled_equals_untiled(srk_ops):
    """Column panels (seam exchange at 1x / 2x / 4x resolution), row bands (13-px LR halo) and band-wise rank shards of the
    EnhanceNet generator reproduce the un-tiled frame bit for bit; a frame wider than one panel matches the oracle."""
    from ml_super_resolution_b200.enet.model_enet import EnetGenerator
    params = _trained_like(OM.enet_g_init(seed=8), scale=2.5)
    net = EnetGenerator(params)
    sd = OM.synthetic_images(44, 1, 48, 56, 3)
    bq = OM.synthetic_images(45, 1, 192, 224, 3)
    sdt, bqt = torch.from_numpy(sd).cuda(), torch.from_numpy(bq).cuda()
    full = net.forward(sdt, bqt).cpu().numpy()
    assert np.array_equal(net.forward(sdt, bqt, max_panel_w=24).cpu().numpy(), full)
    assert np.array_equal(net.forward(sdt, bqt, max_panel_w=30, tile_rows=40).cpu().numpy(), full)
    out = torch.full_like(bqt, float("nan"))
    for r in range(2):
        net.forward(sdt, bqt, out=out, max_panel_w=30, tile_rows=40, rank=r, world=2)
    assert np.array_equal(out.cpu().numpy(), full)
    out = torch.full_like(bqt, float("nan"))  # a band split over three ranks: the column halo is recomputed instead of exchanged
    for r in range(3):
        net.forward(sdt, bqt, out=out, max_panel_w=40, rank=r, world=3)
    assert np.array_equal(out.cpu().numpy(), full)
    sdw = OM.synthetic_images(46, 1, 16, 100, 3)
    bqw = OM.synthetic_images(47, 1, 64, 400, 3)
    got = net.forward(torch.from_numpy(sdw).cuda(), torch.from_numpy(bqw).cuda()).cpu().numpy()
    ref = OM.enet_generator_forward(params, sdw, bqw)
    assert np.abs(got - ref).max() <= TOL_BF16 * max(1.0, np.abs(ref - bqw).max())


def test_espcn_trains_from_reference_tfrecords(srk_ops, tmp_path):
    """patch pairs written in the reference's TFRecord layout feed EspcnNet.train_step through the
    prefetching device iterator; the first step's loss equals the oracle's on the same (decoded) batch."""
    from ml_super_resolution_b200.espcn import dataset as ED
    from ml_super_resolution_b200.espcn.model_espcn import EspcnNet
    params = _trained_like(OM.espcn_init(seed=12, scaling_factor=3, channels=3), scale=4.0)
    lrs = OM.synthetic_images(61, 8, 17, 17, 3)
    hrs = O.pixel_unshuffle(OM.synthetic_images(62, 8, 51, 51, 3), 3)
    for i in range(8):
        ED.write_patch(str(tmp_path / f"{i:03d}.tfrecord"), lrs[i], hrs[i])
    it = ED.build_image_batch_iterator(str(tmp_path), batch_size=8, upscaling_factor=3, seed=0, device="cuda")
    lr_d, hr_d = next(it)
    assert lr_d.shape == (8, 17, 17, 3) and hr_d.shape == (8, 17, 17, 27)
    net = EspcnNet(params, 3, 3)
    loss = float(net.train_step(lr_d, hr_d, 1e-3))
    ref_loss, _, _ = OM.espcn_loss_and_grads(params, lr_d.cpu().numpy(), hr_d.cpu().numpy())
    assert abs(loss - ref_loss) <= 2e-3 * ref_loss
    lr2, _ = next(it)  # second epoch batch arrives through the other buffer set
    assert lr2.shape == (8, 17, 17, 3)
