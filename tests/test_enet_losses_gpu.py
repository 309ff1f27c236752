"""EnhanceNet's training losses against the torch-CPU fp64 restatement of enet/enet/model_enet.py:118-261 and
model_vgg.py (oracle/enet_losses.py): VGG-19 features, perceptual and texture-matching losses, discriminator, log losses, the
gradient handed to the generator and the discriminator's weight gradients.  bf16 activations through 16 + 10 layers: loss values
within 2 %, gradients within 5 % relative L2 (measured values are printed in the assertion messages)."""
import numpy as np
import pytest
import torch

from oracle import enet_losses as E

pytestmark = pytest.mark.gpu


def _rel(a, r):
    return float(np.linalg.norm(np.asarray(a, np.float64) - r) / (np.linalg.norm(r) + 1e-30))


@pytest.fixture(scope="module")
def setup():
    from ml_super_resolution_b200.enet import losses as L
    rng = np.random.default_rng(0)
    vw = L.vgg19_random_weights(1)
    dpar = L.discriminator_params(64, 2)
    for k in dpar:  # trained-like discriminator: larger kernels, non-zero biases (the reference initialises biases to zero)
        dpar[k] = (dpar[k] * 1.6).astype(np.float32) if k.endswith("kernel:0") else (0.05 * rng.standard_normal(dpar[k].shape)).astype(np.float32)
    # smooth-ish images: low-pass noise, as natural patches are
    def img(seed):
        r = np.random.default_rng(seed)
        a = r.standard_normal((2, 16, 16, 3))
        a = np.kron(a, np.ones((1, 4, 4, 1))) * 0.4 + 0.15 * r.standard_normal((2, 64, 64, 3))
        return np.clip(a, -1, 1).astype(np.float32)
    return L, vw, dpar, img(10), img(11)


def test_vgg19_features(srk_ops, setup):
    L, vw, _, sr, _ = setup
    vgg = L.Vgg19(vw)
    acts = vgg.forward(torch.from_numpy(sr).cuda())
    ref = E.vgg19(torch.from_numpy(sr).double(), vw)
    for name in ("block1_conv1", "block2_pool", "block3_conv1", "block5_pool"):
        r = ref[name].numpy()
        assert acts[name].shape == r.shape
        assert _rel(acts[name].float().cpu().numpy(), r) <= 2e-2, name


def test_vgg_gradient_precision_probe(srk_ops, setup):
    """Prints the perceptual-loss gradient error with bf16 and with fp32 (tf32 tensor core) VGG activations."""
    L, vw, dpar, sr, hd = setup
    ref = E.enet_losses(sr, hd, vw, dpar, "p")
    for dt, pr in ((torch.bfloat16, False), (torch.float32, False), (torch.float32, True)):
        vgg = L.Vgg19(vw, dtype=dt, precise=pr)
        x, h = torch.from_numpy(sr).cuda(), torch.from_numpy(hd).cuda()
        loss = torch.zeros(1, device="cuda")
        sa, ha = vgg.forward(x), vgg.forward(h)
        taps = L.perceptual_loss(vgg, sa, ha, loss)
        dsr = torch.zeros_like(x)
        vgg.backward(sa, taps, dsr, accumulate=False)
        print(dt, pr, "p_loss", float(loss), "ref", ref["p_loss"], "dsr rel", _rel(dsr.cpu().numpy(), ref["dsr"]))


def test_losses_and_generator_gradient(srk_ops, setup):
    L, vw, dpar, sr, hd = setup
    from ml_super_resolution_b200.enet.model_enet import EnetPat
    t = EnetPat("pat", vw, None, dpar, hd_size=64)
    dsr = t.losses_and_dsr(torch.from_numpy(sr).cuda(), torch.from_numpy(hd).cuda())
    ref = E.enet_losses(sr, hd, vw, dpar, "pat")
    for k in ("p_loss", "g_loss", "t_loss", "g_loss_all"):
        got = float(t.last[k])
        assert abs(got - ref[k]) <= 2e-2 * abs(ref[k]) + 1e-9, (k, got, ref[k])
    rel = _rel(dsr.cpu().numpy(), ref["dsr"])
    assert rel <= 5e-2, f"d(g_losses)/d(sr) relative L2 error {rel:.4f}"


@pytest.mark.parametrize("pat", ["p", "pa"])
def test_loss_subsets(srk_ops, setup, pat):
    L, vw, dpar, sr, hd = setup
    from ml_super_resolution_b200.enet.model_enet import EnetPat
    t = EnetPat(pat, vw, None, dpar, hd_size=64)
    dsr = t.losses_and_dsr(torch.from_numpy(sr).cuda(), torch.from_numpy(hd).cuda())
    ref = E.enet_losses(sr, hd, vw, dpar, pat)
    assert abs(float(t.last["g_loss_all"]) - ref["g_loss_all"]) <= 2e-2 * abs(ref["g_loss_all"])
    assert _rel(dsr.cpu().numpy(), ref["dsr"]) <= 5e-2


def test_discriminator_loss_and_weight_gradients(srk_ops, setup):
    L, vw, dpar, sr, hd = setup
    d = L.Discriminator(64, dpar)
    loss = torch.zeros(1, device="cuda")
    d.discriminator_loss_and_grads(torch.from_numpy(sr).cuda(), torch.from_numpy(hd).cuda(), loss)
    ref = E.enet_losses(sr, hd, vw, dpar, "pa")
    assert abs(float(loss) - ref["a_loss"]) <= 2e-2 * ref["a_loss"]
    got = d.arena.to_numpy("g")
    worst = max((_rel(got[k], ref["d_grads"][k]), k) for k in got)
    assert worst[0] <= 5e-2, worst
    w0 = d.arena.w.clone()
    d.adam_step(1e-4)
    assert float((d.arena.w - w0).abs().max()) > 0


def test_build_enet_session_training_step(srk_ops, setup):
    """`build_enet` keys and one g_trainer + d_trainer run through the session seam (enet/enet/experiment_train.py's loop)."""
    L, vw, dpar, sr, hd = setup
    from ml_super_resolution_b200.enet.model_enet import build_enet
    from ml_super_resolution_b200.session import Session, placeholder
    rng = np.random.default_rng(3)
    sd = rng.uniform(-1, 1, (2, 16, 16, 3)).astype(np.float32)
    bq = np.clip(hd + 0.05 * rng.standard_normal(hd.shape), -1, 1).astype(np.float32)
    phs = [placeholder([None, None, None, 3], n) for n in ("sd_images", "bq_images", "hd_images")]
    model = build_enet(*phs, "pat", None, d_params=dpar, hd_size=64)
    assert {"sd_images", "bq_images", "sr_images", "hd_images", "step", "p_loss", "g_loss_all", "g_trainer", "a_loss", "g_loss", "d_trainer", "t_loss"} <= set(model)
    with Session() as s:
        feeds = {phs[0]: sd, phs[1]: bq, phs[2]: hd}
        keys = ("g_trainer", "d_trainer", "g_loss_all", "a_loss", "p_loss", "t_loss", "g_loss", "step")
        first = s.run({k: model[k] for k in keys}, feed_dict=feeds)
        for _ in range(3):
            last = s.run({k: model[k] for k in keys}, feed_dict=feeds)
        assert last["step"] == first["step"] + 3 and all(np.isfinite(last[k]) for k in keys if k.endswith("loss") or k == "g_loss_all")
        # discriminator alone on a fixed pair of batches: its loss goes down
        d0 = s.run({"d_trainer": model["d_trainer"], "a_loss": model["a_loss"]}, feed_dict=feeds)["a_loss"]
        for _ in range(30):
            d1 = s.run({"d_trainer": model["d_trainer"], "a_loss": model["a_loss"]}, feed_dict=feeds)["a_loss"]
        assert d1 < d0, (d0, d1)
        sr_img = s.run(model["sr_images"], feed_dict={phs[0]: sd, phs[1]: bq})
    assert sr_img.shape == (2, 64, 64, 3)
