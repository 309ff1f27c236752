"""The generic tcgen05 GEMM (bf16 and tf32) and the general-shape layers built on it (im2col convolution with stride and
TF 'SAME' padding, dense, max-pool): forward, data gradient and weight gradient against torch-CPU fp64 (the oracle's conv
primitive, oracle/ops.py)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _bf16(a):
    return torch.from_numpy(np.ascontiguousarray(a, np.float32)).to(torch.bfloat16)


@pytest.mark.parametrize("M,N,K,batch", [(128, 64, 64, 1), (300, 200, 136, 1), (77, 19, 40, 3), (1000, 513, 72, 1), (256, 256, 1024, 2)])
@pytest.mark.parametrize("tf32", [False, True])
def test_gemm_matches_fp64(srk_ops, M, N, K, batch, tf32):
    from ml_super_resolution_b200 import nn
    rng = np.random.default_rng(M + N + K)
    a = rng.standard_normal((batch, M, K)).astype(np.float32)
    b = rng.standard_normal((batch, N, K)).astype(np.float32)
    bias = rng.standard_normal(N).astype(np.float32)
    if tf32:
        ta, tb = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
        ref_a, ref_b = a.astype(np.float64), b.astype(np.float64)
        tol = 1e-4 * np.sqrt(K)   # 3xTF32: fp32-level operands; the tensor core's accumulator truncates
    else:
        ta, tb = _bf16(a).cuda(), _bf16(b).cuda()
        ref_a, ref_b = ta.float().cpu().numpy().astype(np.float64), tb.float().cpu().numpy().astype(np.float64)
        tol = 1e-4 * np.sqrt(K)   # exact products of the rounded operands, fp32 accumulation
    if batch == 1:
        ta, tb = ta[0], tb[0]
    got = nn.gemm(ta, tb, torch.from_numpy(bias).cuda(), "leaky_relu", 0.2, out_dtype=torch.float32).cpu().numpy().reshape(batch, M, N)
    ref = np.einsum("bmk,bnk->bmn", ref_a, ref_b) + bias
    ref = np.where(ref > 0, ref, 0.2 * ref)
    assert np.abs(got - ref).max() <= tol, np.abs(got - ref).max()
    if tf32:  # one tf32 product (operands truncated to 10 mantissa bits by the tensor core): error ~ 1e-3 per operand
        one = nn.gemm(ta, tb, torch.from_numpy(bias).cuda(), "leaky_relu", 0.2, out_dtype=torch.float32, precise=False).cpu().numpy().reshape(batch, M, N)
        assert np.abs(one - ref).max() <= 5e-3 * np.sqrt(K), np.abs(one - ref).max()


def _conv_ref(x, w, b, stride, pad, act):
    """TF semantics on torch-CPU fp64: 'SAME' pads pad_total // 2 before and the rest after."""
    xt = torch.from_numpy(x).double().permute(0, 3, 1, 2).requires_grad_(True)
    wt = torch.from_numpy(w).double().permute(3, 2, 0, 1).requires_grad_(True)
    bt = torch.from_numpy(b).double().requires_grad_(True)
    k = w.shape[0]
    if pad == "SAME":
        def pads(n):
            o = -(-n // stride)
            t = max((o - 1) * stride + k - n, 0)
            return t // 2, t - t // 2
        (pt, pb), (pl, pr) = pads(x.shape[1]), pads(x.shape[2])
        xp = F.pad(xt, (pl, pr, pt, pb))
    else:
        xp = xt
    y = F.conv2d(xp, wt, bt, stride=stride)
    if act == "leaky_relu":
        y = F.leaky_relu(y, 0.2)
    elif act == "relu":
        y = F.relu(y)
    elif act == "tanh":
        y = torch.tanh(y)
    return xt, wt, bt, y


@pytest.mark.parametrize("shape,cout,k,stride,pad,act", [((2, 16, 16, 32), 64, 3, 1, "SAME", "leaky_relu"), ((2, 16, 16, 32), 32, 3, 2, "SAME", "leaky_relu"),
                                                          ((1, 15, 13, 8), 24, 3, 2, "SAME", "relu"), ((2, 12, 12, 3), 64, 3, 1, "SAME", "relu"),
                                                          ((1, 20, 20, 1), 64, 5, 1, "SAME", "tanh"), ((1, 17, 17, 3), 64, 9, 1, "VALID", "relu")])
@pytest.mark.parametrize("dtype", ["bf16", "tf32"])
def test_conv_forward_dgrad_wgrad(srk_ops, shape, cout, k, stride, pad, act, dtype):
    from ml_super_resolution_b200 import nn
    rng = np.random.default_rng(sum(shape) + cout + k)
    x = rng.uniform(-1, 1, shape).astype(np.float32)
    w = (rng.standard_normal((k, k, shape[3], cout)) / np.sqrt(k * k * shape[3])).astype(np.float32)
    b = (0.1 * rng.standard_normal(cout)).astype(np.float32)
    td = torch.bfloat16 if dtype == "bf16" else torch.float32
    layer = nn.Conv(torch.from_numpy(w).cuda(), torch.from_numpy(b).cuda(), stride, pad, act, 0.2, dtype=td)
    xd = torch.from_numpy(x).cuda().to(td)
    y = layer.forward(xd)
    xt, wt, bt, yref = _conv_ref(xd.float().cpu().numpy(), w, b, stride, pad, act)
    yr = yref.permute(0, 2, 3, 1).detach().numpy()
    tol = 3e-2 if dtype == "bf16" else 1e-4
    assert y.shape == yr.shape and np.abs(y.float().cpu().numpy() - yr).max() <= tol
    dy = rng.standard_normal(yr.shape).astype(np.float32)
    dyd = torch.from_numpy(dy).cuda().to(td)
    yref.backward(torch.from_numpy(dyd.float().cpu().numpy()).double().permute(0, 3, 1, 2))
    grads = {}
    dx = layer.backward(xd, y, dyd, True, grads, ("w", "b"))
    rel = lambda a, r: np.linalg.norm(a - r) / (np.linalg.norm(r) + 1e-30)  # noqa: E731
    # bf16: pre-activations within rounding distance of zero flip the (leaky-)ReLU mask -- relative L2 ~ 0.8 * sqrt(flipped fraction)
    gate = 4e-2 if dtype == "bf16" else 1e-4
    assert rel(dx.float().cpu().numpy(), xt.grad.permute(0, 2, 3, 1).numpy()) <= gate
    assert rel(grads["w"].cpu().numpy(), wt.grad.permute(2, 3, 1, 0).numpy()) <= gate
    assert rel(grads["b"].cpu().numpy(), bt.grad.numpy()) <= gate


def test_dense_and_single_unit_head(srk_ops):
    from ml_super_resolution_b200 import nn
    rng = np.random.default_rng(3)
    x = rng.standard_normal((10, 256)).astype(np.float32)
    for nout, act in ((64, "leaky_relu"), (1, "sigmoid")):
        w = (rng.standard_normal((256, nout)) / 16).astype(np.float32)
        b = (0.1 * rng.standard_normal(nout)).astype(np.float32)
        layer = nn.Dense(torch.from_numpy(w).cuda(), torch.from_numpy(b).cuda(), act)
        xd = _bf16(x).cuda()
        y = layer.forward(xd, out_dtype=torch.float32)
        xt = xd.float().cpu().double().requires_grad_(True)
        wt = torch.from_numpy(w).double().requires_grad_(True)
        z = xt @ wt + torch.from_numpy(b).double()
        yr = F.leaky_relu(z, 0.2) if act == "leaky_relu" else torch.sigmoid(z)
        assert y.shape == (10, nout) and np.abs(y.cpu().numpy() - yr.detach().numpy()).max() <= 2e-2
        dy = rng.standard_normal((10, nout)).astype(np.float32)
        yr.backward(torch.from_numpy(dy).double())
        grads = {}
        dx = layer.backward(xd, y, torch.from_numpy(dy).cuda(), grads, ("w", "b"))
        rel = lambda a, r: np.linalg.norm(a - r) / (np.linalg.norm(r) + 1e-30)  # noqa: E731
        assert rel(dx.float().cpu().numpy(), xt.grad.numpy()) <= 2e-2
        assert rel(grads["w"].cpu().numpy(), wt.grad.numpy()) <= 2e-2


@pytest.mark.parametrize("shape", [(2, 8, 8, 16), (1, 7, 5, 8)])
def test_maxpool_forward_backward(srk_ops, shape):
    from ml_super_resolution_b200 import nn
    rng = np.random.default_rng(5)
    x = rng.standard_normal(shape).astype(np.float32)
    xd = torch.from_numpy(x).cuda()
    y = nn.maxpool2x2(xd)
    xt = torch.from_numpy(x).double().permute(0, 3, 1, 2).requires_grad_(True)
    yr = F.max_pool2d(xt, 2, 2, ceil_mode=True)
    assert np.array_equal(y.cpu().numpy(), yr.permute(0, 2, 3, 1).detach().numpy().astype(np.float32))
    dy = rng.standard_normal(tuple(y.shape)).astype(np.float32)
    yr.backward(torch.from_numpy(dy).double().permute(0, 3, 1, 2))
    dx = nn.maxpool2x2_bwd(xd, torch.from_numpy(dy).cuda())
    assert np.allclose(dx.cpu().numpy(), xt.grad.permute(0, 2, 3, 1).numpy())


# ------------------------------------------------------------------------------------------------ tf32 model variants (north_star: 1e-3)
TOL_TF32 = 2e-3  # max-abs on the reference's [-1,1] range = north_star's 1e-3 on [0,1] images


def _trained_like(params, scale=1.0, seed=5):
    rng = np.random.default_rng(seed)
    return {k: ((0.05 * rng.standard_normal(v.shape)) if k.endswith(("bias:0", "biases:0")) else v * scale).astype(np.float32) for k, v in params.items()}


def test_tf32_vdsr20_forward(srk_ops):
    from ml_super_resolution_b200 import tf32
    from oracle import models as OM
    from oracle import ops as O
    params = _trained_like(OM.vdsr_init(seed=42))
    hd = OM.synthetic_images(1236, 4, 41, 41, 3)
    sd = np.ascontiguousarray(np.stack([O.hd_image_to_sd_image(h * 0.5 + 0.5, 3) * 2 - 1 for h in hd]), dtype=np.float32)
    got = tf32.vdsr_forward(params, torch.from_numpy(sd).cuda()).cpu().numpy()
    ref = OM.vdsr_forward(params, sd)["sr_images"]
    err = np.abs(got - ref).max()
    assert err <= TOL_TF32, f"max-abs {err:.2e}"


@pytest.mark.parametrize("channels,r", [(1, 3), (3, 3), (3, 4)])
def test_tf32_espcn_forward(srk_ops, channels, r):
    from ml_super_resolution_b200 import tf32
    from oracle import models as OM
    params = _trained_like(OM.espcn_init(seed=9, scaling_factor=r, channels=channels), scale=5.0)
    lr = OM.synthetic_images(5, 2, 40, 52, channels)
    got = tf32.espcn_forward(params, torch.from_numpy(lr).cuda()).cpu().numpy()
    ref = OM.espcn_forward(params, lr)
    err = np.abs(got - ref).max()
    assert err <= TOL_TF32, f"max-abs {err:.2e}"


def test_tf32_srcnn_forward(srk_ops):
    from ml_super_resolution_b200 import tf32
    from oracle import models as OM
    params = _trained_like(OM.srcnn_init(seed=6, channels=1), scale=60.0)
    lo = OM.synthetic_images(31, 4, 33, 33, 1)
    got = tf32.srcnn_forward(params, torch.from_numpy(lo).cuda()).cpu().numpy()
    with torch.no_grad():
        ref = OM.srcnn_forward_t(OM._to_t(params, np.float64), OM._t(lo, np.float64)).numpy()
    err = np.abs(got - ref).max()
    assert got.shape == ref.shape and err <= TOL_TF32, f"max-abs {err:.2e}"
