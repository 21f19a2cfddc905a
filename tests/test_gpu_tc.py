"""GPU parity for the bf16 tensor-core kernels (tcgen05 implicit-GEMM conv) through the C ABI.

Integer-valued data makes bf16 products and FP32 sums EXACT, so values, pool arg-max indices (including the
reference's first-max tie rule, which integer data hits constantly) and ReLU masks must match the oracle
bit for bit.  A second case uses real-valued data with the stated bf16 tolerance."""
import numpy as np
import pytest

from oracle import np_ref as R

pytestmark = pytest.mark.gpu


def to_bf16_bits(x):
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)
    return ((u + 0x7FFF + ((u >> 16) & 1)) >> 16).astype(np.uint16)


def from_bf16_bits(b):
    return (b.astype(np.uint32) << 16).view(np.float32)


def bf16_round(x):
    return from_bf16_bits(to_bf16_bits(x))


def pad_nhwc(x_nchw):
    """NCHW float -> zero-haloed NHWC bf16 bits [N,H+2,W+2,C]."""
    n, c, h, w = x_nchw.shape
    out = np.zeros((n, h + 2, w + 2, c), np.float32)
    out[:, 1:-1, 1:-1, :] = x_nchw.transpose(0, 2, 3, 1)
    return to_bf16_bits(out)


def unpack_nibbles(b):
    """uint8 [..., C/2] -> [..., C] nibbles (low nibble first)."""
    lo, hi = b & 0xF, b >> 4
    return np.stack([lo, hi], axis=-1).reshape(b.shape[:-1] + (b.shape[-1] * 2,))


def _setup(n, seed, integer):
    rng = np.random.default_rng(seed)
    if integer:
        x = rng.integers(-2, 3, (n, 32, 14, 14)).astype(np.float32)
        w = rng.integers(-1, 2, (64, 32, 3, 3)).astype(np.float32)
    else:
        x = bf16_round(rng.standard_normal((n, 32, 14, 14)).astype(np.float32))
        w = bf16_round((rng.standard_normal((64, 32, 3, 3)) * 0.1).astype(np.float32))
    return rng, x, w


def _pack(w):
    from neuromah_b200._lib import lib
    from neuromah_b200.device import DeviceArray, stream
    wd = DeviceArray.from_numpy(w)
    wf = DeviceArray((64, 9, 32), np.uint16)
    wg = DeviceArray((32, 9, 64), np.uint16)
    lib.nm_tc_pack_conv_weights_bf16(wd.p, wf.p, wg.p, 64, 32, stream())
    return wf, wg


@pytest.mark.parametrize("n,integer", [(3, True), (64, True), (37, False)])
def test_tc_conv_fprop_relu_pool(n, integer):
    from neuromah_b200._lib import lib
    from neuromah_b200.device import DeviceArray, stream
    rng, x, w = _setup(n, 10 + n, integer)
    wf, _ = _pack(w)
    np.testing.assert_array_equal(from_bf16_bits(wf.numpy()), w.transpose(0, 2, 3, 1).reshape(64, 9, 32))
    xp = DeviceArray.from_numpy(pad_nhwc(x))
    y = DeviceArray.zeros((n, 7, 7, 64), np.uint16)
    idx = DeviceArray.zeros((n, 7, 7, 32), np.uint8)
    lib.nm_tc_conv3x3_fprop_pool_bf16(xp.p, wf.p, y.p, idx.p, n, 14, 14, 32, 64, 0, stream())
    conv, _ = R.conv_layer_forward(x, w, None, padding="same", relu=True, dtype=np.float64)
    pooled, pidx = R.maxpool2d_forward_cpu(conv.astype(np.float32), 2, 2, 2, 2, "valid")
    got = from_bf16_bits(y.numpy()).transpose(0, 3, 1, 2)                     # -> NCHW
    nib = unpack_nibbles(idx.numpy()).transpose(0, 3, 1, 2)
    exp_pos = (pidx[..., 0] % 2) * 2 + (pidx[..., 1] % 2)
    if integer:
        np.testing.assert_array_equal(got, bf16_round(pooled))
        np.testing.assert_array_equal(nib & 3, exp_pos)
        np.testing.assert_array_equal(nib >> 2, (pooled > 0).astype(np.uint8))
    else:
        np.testing.assert_allclose(got, pooled, rtol=2 ** -7, atol=1e-3)      # bf16 output rounding
        assert np.mean((nib & 3) != exp_pos) < 2e-3                           # near-ties only
    # padded output variant (for a following conv): interior equals the dense result, halo stays zero
    yp = DeviceArray.zeros((n, 9, 9, 64), np.uint16)
    lib.nm_tc_conv3x3_fprop_pool_bf16(xp.p, wf.p, yp.p, idx.p, n, 14, 14, 32, 64, 1, stream())
    ypn = yp.numpy()
    np.testing.assert_array_equal(ypn[:, 1:-1, 1:-1, :], y.numpy())
    assert not ypn[:, 0].any() and not ypn[:, -1].any() and not ypn[:, :, 0].any() and not ypn[:, :, -1].any()


@pytest.mark.parametrize("n,integer", [(2, True), (50, True), (19, False)])
def test_tc_conv_dgrad(n, integer):
    from neuromah_b200._lib import lib
    from neuromah_b200.device import DeviceArray, stream
    rng, x, w = _setup(n, 20 + n, integer)
    _, wg = _pack(w)
    dy = (rng.integers(-2, 3, (n, 64, 14, 14)) if integer else bf16_round(rng.standard_normal((n, 64, 14, 14)))).astype(np.float32)
    dyp = DeviceArray.from_numpy(pad_nhwc(dy))
    dx = DeviceArray.zeros((n, 16, 16, 32), np.uint16)
    lib.nm_tc_conv3x3_dgrad_bf16(dyp.p, wg.p, dx.p, n, 14, 14, 32, 64, stream())
    ref, _, _ = R.conv_layer_backward(x, w, None, dy, padding="same", relu=False, mode="wired", dtype=np.float64)
    got = from_bf16_bits(dx.numpy())[:, :14, :14, :].transpose(0, 3, 1, 2)
    if integer:
        np.testing.assert_array_equal(got, bf16_round(ref.astype(np.float32)))
    else:
        np.testing.assert_allclose(got, ref, rtol=2 ** -7, atol=2e-2)


@pytest.mark.parametrize("n,integer", [(2, True), (300, True), (41, False)])
def test_tc_conv_wgrad(n, integer):
    from neuromah_b200._lib import lib
    from neuromah_b200.device import DeviceArray, Scratch, stream
    rng, x, w = _setup(n, 30 + n, integer)
    dy = (rng.integers(-2, 3, (n, 64, 14, 14)) if integer else bf16_round(rng.standard_normal((n, 64, 14, 14)))).astype(np.float32)
    xp, dyp = DeviceArray.from_numpy(pad_nhwc(x)), DeviceArray.from_numpy(pad_nhwc(dy))
    dw = DeviceArray.zeros((64, 32, 3, 3))
    ws = DeviceArray((lib.nm_tc_conv3x3_wgrad_workspace(32, 64),), np.uint8)
    lib.nm_tc_conv3x3_wgrad_bf16(xp.p, dyp.p, dw.p, ws.p, ws.nbytes, n, 14, 14, 32, 64, stream())
    _, ref, _ = R.conv_layer_backward(x, w, None, dy, padding="same", relu=False, mode="wired", dtype=np.float64)
    if integer:
        np.testing.assert_array_equal(dw.numpy(), ref.astype(np.float32))
    else:
        np.testing.assert_allclose(dw.numpy(), ref, rtol=1e-4, atol=1e-2)
    dw2 = DeviceArray.zeros((64, 32, 3, 3))                        # deterministic: bitwise repeatable
    lib.nm_tc_conv3x3_wgrad_bf16(xp.p, dyp.p, dw2.p, ws.p, ws.nbytes, n, 14, 14, 32, 64, stream())
    np.testing.assert_array_equal(dw.numpy(), dw2.numpy())


@pytest.mark.parametrize("n,integer", [(2, True), (300, True), (41, False)])
def test_tc_conv_fused_backward(n, integer):
    """dgrad + wgrad from one pass over dY (nm_tc_conv3x3_bwd_bf16) equals the two separate kernels' references."""
    from neuromah_b200._lib import lib
    from neuromah_b200.device import DeviceArray, stream
    rng, x, w = _setup(n, 40 + n, integer)
    _, wg = _pack(w)
    dy = (rng.integers(-2, 3, (n, 64, 14, 14)) if integer else bf16_round(rng.standard_normal((n, 64, 14, 14)))).astype(np.float32)
    xp, dyp = DeviceArray.from_numpy(pad_nhwc(x)), DeviceArray.from_numpy(pad_nhwc(dy))
    dw = DeviceArray.zeros((64, 32, 3, 3))
    dx = DeviceArray.zeros((n, 16, 16, 32), np.uint16)
    ws = DeviceArray((lib.nm_tc_conv3x3_bwd_workspace(32, 64),), np.uint8)
    lib.nm_tc_conv3x3_bwd_bf16(xp.p, dyp.p, wg.p, dw.p, dx.p, ws.p, ws.nbytes, n, 14, 14, 32, 64, stream())
    dx_ref, dw_ref, _ = R.conv_layer_backward(x, w, None, dy, padding="same", relu=False, mode="wired", dtype=np.float64)
    got = from_bf16_bits(dx.numpy())[:, :14, :14, :].transpose(0, 3, 1, 2)
    if integer:
        np.testing.assert_array_equal(got, bf16_round(dx_ref.astype(np.float32)))
        np.testing.assert_array_equal(dw.numpy(), dw_ref.astype(np.float32))
    else:
        np.testing.assert_allclose(got, dx_ref, rtol=2 ** -7, atol=2e-2)
        np.testing.assert_allclose(dw.numpy(), dw_ref, rtol=1e-4, atol=1e-2)
    dw2 = DeviceArray.zeros((64, 32, 3, 3))                        # deterministic: bitwise repeatable
    lib.nm_tc_conv3x3_bwd_bf16(xp.p, dyp.p, wg.p, dw2.p, dx.p, ws.p, ws.nbytes, n, 14, 14, 32, 64, stream())
    np.testing.assert_array_equal(dw.numpy(), dw2.numpy())


@pytest.mark.parametrize("n,integer", [(1, True), (67, True), (300, False)])
def test_tc_conv_fused_backward_from_pooled_gradient(n, integer):
    """nm_tc_conv3x3_bwd_pool_bf16: dY is rebuilt in shared memory from the pooled gradient + arg-max nibbles
    (MaxPooling2D.backward routing, maxpooling_backend.cpp:145-238) and never exists in HBM."""
    from neuromah_b200._lib import lib
    from neuromah_b200.device import DeviceArray, stream
    rng, x, w = _setup(n, 60 + n, integer)
    _, wg = _pack(w)
    nib = rng.integers(0, 4, (n, 7, 7, 64)).astype(np.uint8) | (rng.integers(0, 2, (n, 7, 7, 64)).astype(np.uint8) << 2)
    idx = (nib[..., 0::2] | (nib[..., 1::2] << 4)).astype(np.uint8)
    live = (nib >> 2).astype(bool)
    dp = (rng.integers(-2, 3, (n, 7, 7, 64)) if integer else bf16_round(rng.standard_normal((n, 7, 7, 64)))).astype(np.float32)
    dp = np.where(live, dp, 0.0).astype(np.float32)                 # the producer applies the ReLU mask
    dy = np.zeros((n, 14, 14, 64), np.float32)
    for pos in range(4):
        dy[:, (pos >> 1)::2, (pos & 1)::2, :] = np.where((nib & 3) == pos, dp, 0.0)
    dy = dy.transpose(0, 3, 1, 2).copy()
    xp = DeviceArray.from_numpy(pad_nhwc(x))
    dpd, idxd = DeviceArray.from_numpy(to_bf16_bits(dp)), DeviceArray.from_numpy(idx)
    dw = DeviceArray.zeros((64, 32, 3, 3))
    dx = DeviceArray.zeros((n, 16, 16, 32), np.uint16)
    ws = DeviceArray((lib.nm_tc_conv3x3_bwd_workspace(32, 64),), np.uint8)
    lib.nm_tc_conv3x3_bwd_pool_bf16(xp.p, dpd.p, idxd.p, wg.p, dw.p, dx.p, ws.p, ws.nbytes, n, 14, 14, 32, 64, stream())
    dx_ref, dw_ref, _ = R.conv_layer_backward(x, w, None, dy, padding="same", relu=False, mode="wired", dtype=np.float64)
    got = from_bf16_bits(dx.numpy())[:, :14, :14, :].transpose(0, 3, 1, 2)
    if integer:
        np.testing.assert_array_equal(got, bf16_round(dx_ref.astype(np.float32)))
        np.testing.assert_array_equal(dw.numpy(), dw_ref.astype(np.float32))
    else:
        np.testing.assert_allclose(got, dx_ref, rtol=2 ** -7, atol=2e-2)
        np.testing.assert_allclose(dw.numpy(), dw_ref, rtol=1e-4, atol=1e-2)


# ------------------------------------------------------------------ tcgen05 Dense head ----
def _dense_case(B, K, n_out, hw, c, seed, integer):
    rng = np.random.default_rng(seed)
    if integer:
        x = rng.integers(0, 3, (B, K)).astype(np.float32)                      # bf16-exact activations
        w = (rng.integers(-4, 5, (K, n_out)) / 64.0).astype(np.float32)        # exact in bf16 -> lo half is zero
    else:
        x = bf16_round(np.abs(rng.standard_normal((B, K))).astype(np.float32))
        w = (rng.standard_normal((K, n_out)) * 0.02).astype(np.float32)        # full FP32 weights: exercises hi + lo
    b = (rng.standard_normal((1, n_out)) * 0.1).astype(np.float32)
    y = rng.integers(0, n_out, B).astype(np.int32)
    return rng, x, w, b, y


def _hwc_from_chw(a_chw_cols, hw, c):
    """[.., K] with K in the reference (c,h,w) order -> the kernels' (h,w,c) order."""
    lead = a_chw_cols.shape[:-1]
    return np.ascontiguousarray(a_chw_cols.reshape(lead + (c, hw)).swapaxes(-1, -2).reshape(lead + (hw * c,)))


@pytest.mark.parametrize("B,K,n_out,hw,c,integer", [(4096, 3136, 10, 49, 64, False), (300, 3136, 10, 49, 64, True),
                                                    (129, 512, 16, 8, 64, False), (7, 64, 3, 1, 64, True)])
def test_tc_dense_head_forward_and_wgrad(B, K, n_out, hw, c, integer):
    """Dense.py:79-80 + Softmax.py:7-11 + categorical_crossentropy.py:8-25 + fused gradient, then Dense.py:96-105."""
    from neuromah_b200._lib import lib
    from neuromah_b200.device import DeviceArray, stream
    rng, x_ref, w, b, y = _dense_case(B, K, n_out, hw, c, 50 + B, integer)
    s = stream()
    x_k = _hwc_from_chw(x_ref, hw, c)                          # what the conv stack hands over: (h,w,c) order
    xd = DeviceArray.from_numpy(to_bf16_bits(x_k))
    wd, bd, yd = DeviceArray.from_numpy(w), DeviceArray.from_numpy(b), DeviceArray.from_numpy(y)
    wp = DeviceArray.zeros((n_out, K))
    whl = DeviceArray.zeros((32, K), np.uint16)
    lib.nm_tc_pack_dense_weights(wd.p, wp.p, whl.p, K, n_out, hw, c, s)
    np.testing.assert_array_equal(wp.numpy(), _hwc_from_chw(w.T.copy(), hw, c))
    hl = from_bf16_bits(whl.numpy())
    np.testing.assert_allclose(hl[:n_out] + hl[16:16 + n_out], wp.numpy(), rtol=2 ** -15, atol=1e-30)
    assert not hl[n_out:16].any() and not hl[16 + n_out:].any()
    probs, dl = DeviceArray.zeros((B, n_out)), DeviceArray.zeros((B, n_out))
    dlh = DeviceArray.zeros((B, 32), np.uint16)
    losses, correct = DeviceArray.zeros((B,)), DeviceArray.zeros((B,), np.int32)
    accum = DeviceArray.zeros((2,), np.float64)
    ws = DeviceArray.zeros((lib.nm_tc_dense_softmax_ce_workspace(B, K),), np.uint8)
    for _ in range(2):                                         # twice: the ticket counter must be left at zero
        lib.nm_tc_dense_softmax_ce_bf16(xd.p, whl.p, bd.p, yd.p, probs.p, dl.p, dlh.p, losses.p, correct.p, accum.p,
                                        ws.p, ws.nbytes, B, K, n_out, 1.0 / B, s)
    z = x_ref.astype(np.float64) @ w.astype(np.float64) + b
    p_ref = R.softmax_forward(z)
    tol = dict(rtol=1e-5, atol=1e-6) if integer else dict(rtol=2e-4, atol=2e-6)
    np.testing.assert_allclose(probs.numpy(), p_ref, **tol)
    loss_ref = R.cce_forward(p_ref, y)
    np.testing.assert_allclose(losses.numpy(), loss_ref, rtol=1e-4, atol=1e-5)
    d_ref = R.softmax_cce_backward(p_ref, y)
    np.testing.assert_allclose(dl.numpy(), d_ref, rtol=1e-4, atol=1e-5 / B)
    pred = np.argmax(probs.numpy(), axis=1)
    np.testing.assert_array_equal(correct.numpy(), (pred == y).astype(np.int32))
    acc = accum.numpy()
    np.testing.assert_allclose(acc[0], 2 * losses.numpy().astype(np.float64).sum(), rtol=1e-9)
    assert acc[1] == 2 * correct.numpy().sum()
    dh = from_bf16_bits(dlh.numpy())
    np.testing.assert_allclose(dh[:, :n_out] + dh[:, 16:16 + n_out], dl.numpy(), rtol=2 ** -15, atol=1e-30)

    # wgrad from the same tensors
    dw, db = DeviceArray.zeros((K, n_out)), DeviceArray.zeros((1, n_out))
    ws2 = DeviceArray.zeros((lib.nm_tc_dense_wgrad_workspace(B, K),), np.uint8)
    lib.nm_tc_dense_wgrad_bf16(xd.p, dlh.p, dl.p, wd.p, dw.p, db.p, ws2.p, ws2.nbytes, B, K, n_out, hw, c,
                               1.0 / B, 0.0, 0.0, s)
    dlr = dl.numpy().astype(np.float64)
    dw_ref = x_ref.astype(np.float64).T @ dlr / B
    db_ref = dlr.sum(axis=0, keepdims=True) / B
    np.testing.assert_allclose(dw.numpy(), dw_ref, rtol=1e-4, atol=1e-5 * np.abs(dw_ref).max())
    np.testing.assert_allclose(db.numpy(), db_ref, rtol=1e-4, atol=1e-5 * np.abs(db_ref).max())
    dw2 = DeviceArray.zeros((K, n_out))                        # deterministic
    lib.nm_tc_dense_wgrad_bf16(xd.p, dlh.p, dl.p, wd.p, dw2.p, db.p, ws2.p, ws2.nbytes, B, K, n_out, hw, c,
                               1.0 / B, 0.0, 0.0, s)
    np.testing.assert_array_equal(dw.numpy(), dw2.numpy())
    # a dlogits tensor that did not come from the fused kernel: pack it, L1/L2 terms on
    lib.nm_tc_pack_dlogits_bf16(dl.p, dlh.p, B, n_out, s)
    lib.nm_tc_dense_wgrad_bf16(xd.p, dlh.p, dl.p, wd.p, dw2.p, db.p, ws2.p, ws2.nbytes, B, K, n_out, hw, c,
                               1.0 / B, 0.01, 0.02, s)
    reg = 0.01 * np.sign(w) + 2 * 0.02 * w
    np.testing.assert_allclose(dw2.numpy(), dw_ref + reg, rtol=1e-4, atol=1e-5 * np.abs(dw_ref + reg).max())


@pytest.mark.parametrize("B", [5, 700])
def test_tc_dense_bwd_unpool(B):
    """Dense.py:108 + maxpooling_backend.cpp:145-238 + Relu.py:12-14 fused into the last conv's dY."""
    from neuromah_b200._lib import lib
    from neuromah_b200.device import DeviceArray, stream
    rng = np.random.default_rng(70 + B)
    n_out, PH, PW, Cc = 10, 7, 7, 64
    K = PH * PW * Cc
    w = (rng.standard_normal((K, n_out)) * 0.05).astype(np.float32)          # reference (c,h,w) rows
    dl = (rng.standard_normal((B, n_out)) / B).astype(np.float32)
    nib = rng.integers(0, 8, (B, PH, PW, Cc)).astype(np.uint8)               # position | relu-bit << 2
    idx = (nib[..., 0::2] | (nib[..., 1::2] << 4)).astype(np.uint8)
    s = stream()
    wd = DeviceArray.from_numpy(w)
    wp = DeviceArray.zeros((n_out, K))
    lib.nm_tc_pack_dense_weights(wd.p, wp.p, None, K, n_out, PH * PW, Cc, s)
    dy = DeviceArray.zeros((B, 2 * PH + 2, 2 * PW + 2, Cc), np.uint16)
    dbias = DeviceArray.zeros((Cc,))
    ws = DeviceArray.zeros((lib.nm_tc_dense_bwd_unpool_workspace(Cc),), np.uint8)
    dld, idxd = DeviceArray.from_numpy(dl), DeviceArray.from_numpy(idx)     # keep the device buffers alive across the launch
    lib.nm_tc_dense_bwd_unpool_bf16(dld.p, wp.p, idxd.p, dy.p, dbias.p, ws.p, ws.nbytes, B, PH, PW, Cc, n_out, s)
    dflat = dl.astype(np.float64) @ w.astype(np.float64).T                    # (B, K) in (c,h,w) order
    dpool = dflat.reshape(B, Cc, PH, PW).transpose(0, 2, 3, 1)                # -> (B, PH, PW, C)
    live = (nib >> 2).astype(bool)
    ref = np.zeros((B, 2 * PH + 2, 2 * PW + 2, Cc))
    for pos in range(4):
        sel = live & ((nib & 3) == pos)
        ref[:, 1 + (pos >> 1):1 + 2 * PH:2, 1 + (pos & 1):1 + 2 * PW:2, :] = np.where(sel, dpool, 0.0)
    got = from_bf16_bits(dy.numpy())
    np.testing.assert_allclose(got, ref, rtol=2 ** -7, atol=1e-12)
    assert not got[:, 0].any() and not got[:, -1].any() and not got[:, :, 0].any() and not got[:, :, -1].any()
    np.testing.assert_allclose(dbias.numpy(), np.where(live, dpool, 0.0).sum(axis=(0, 1, 2)), rtol=1e-4, atol=1e-6)
    # pooled variant: ReLU-masked gradient of the pooled tensor only (the routing is left to the conv backward kernel)
    dpl = DeviceArray.zeros((B, PH * PW, Cc), np.uint16)
    dbias2 = DeviceArray.zeros((Cc,))
    lib.nm_tc_dense_bwd_pool_bf16(dld.p, wp.p, idxd.p, dpl.p, dbias2.p, ws.p, ws.nbytes, B, PH, PW, Cc, n_out, s)
    np.testing.assert_allclose(from_bf16_bits(dpl.numpy()).reshape(B, PH, PW, Cc), np.where(live, dpool, 0.0), rtol=2 ** -7, atol=1e-12)
    np.testing.assert_array_equal(dbias2.numpy(), dbias.numpy())


# ------------------------------------------------------------------ tcgen05 conv1 (C_in = 1) ----
@pytest.mark.parametrize("n,integer", [(1, True), (37, True), (300, False)])
def test_tc_conv1_forward_backward(n, integer):
    """First conv (1->32, 3x3 same, ReLU) + 2x2 pool on tensor cores via the 4x4-patch formulation, fwd and bwd."""
    from neuromah_b200._lib import lib
    from neuromah_b200.device import DeviceArray, stream
    rng = np.random.default_rng(90 + n)
    if integer:
        x = rng.integers(-3, 4, (n, 1, 28, 28)).astype(np.float32)
        w = rng.integers(-2, 3, (32, 1, 3, 3)).astype(np.float32)
    else:
        x = bf16_round(rng.random((n, 1, 28, 28), dtype=np.float32))
        w = bf16_round((rng.standard_normal((32, 1, 3, 3)) * 0.3).astype(np.float32))
    s = stream()
    xd, wd = DeviceArray.from_numpy(x), DeviceArray.from_numpy(w)
    we = DeviceArray.zeros((128, 16), np.uint16)
    lib.nm_tc_pack_conv1_weights_bf16(wd.p, we.p, 32, s)
    y = DeviceArray.zeros((n, 16, 16, 32), np.uint16)
    idx = DeviceArray.zeros((n, 14, 14, 16), np.uint8)
    lib.nm_tc_conv1_fwd_pool_bf16(xd.p, we.p, y.p, idx.p, n, 28, 28, 32, s)
    conv, _ = R.conv_layer_forward(x, w, None, padding="same", relu=True, dtype=np.float64)
    pooled, pidx = R.maxpool2d_forward_cpu(conv.astype(np.float32), 2, 2, 2, 2, "valid")
    yn = from_bf16_bits(y.numpy())
    got = yn[:, 1:-1, 1:-1, :].transpose(0, 3, 1, 2)
    assert not yn[:, 0].any() and not yn[:, -1].any() and not yn[:, :, 0].any() and not yn[:, :, -1].any()
    nib = unpack_nibbles(idx.numpy()).transpose(0, 3, 1, 2)
    exp_pos = (pidx[..., 0] % 2) * 2 + (pidx[..., 1] % 2)
    if integer:
        np.testing.assert_array_equal(got, bf16_round(pooled))
        np.testing.assert_array_equal(nib & 3, exp_pos)
        np.testing.assert_array_equal(nib >> 2, (pooled > 0).astype(np.uint8))
    else:
        np.testing.assert_allclose(got, pooled, rtol=2 ** -7, atol=1e-3)
        # the kernel takes the first maximum of the bf16-ROUNDED conv values (the values the rest of the plan sees): where
        # the oracle's float maximum is unique after rounding the two agree, elsewhere the earlier position wins
        cb = bf16_round(np.maximum(conv, 0).astype(np.float32))
        win = np.stack([cb[:, :, dy::2, dx::2] for dy in (0, 1) for dx in (0, 1)], axis=-1)
        exp_bf16 = np.argmax(win, axis=-1)
        np.testing.assert_array_equal(nib & 3, exp_bf16)
        np.testing.assert_array_equal(nib >> 2, (win.max(axis=-1) > 0).astype(np.uint8))
        assert np.mean((nib & 3) != exp_pos) < 2e-2                           # near-ties only

    # backward: gradient of the pooled output on the un-shifted grid, routed by the kernel's own nibbles
    dp = (rng.integers(-2, 3, (n, 32, 14, 14)) if integer else bf16_round(rng.standard_normal((n, 32, 14, 14)))).astype(np.float32)
    grid = np.zeros((n, 16, 16, 32), np.float32)
    grid[:, :14, :14, :] = dp.transpose(0, 2, 3, 1)
    grid[:, 14:, :, :] = 7.0                                                  # don't-care rows/cols must be ignored
    grid[:, :, 14:, :] = 7.0
    dw, db = DeviceArray.zeros((32, 1, 3, 3)), DeviceArray.zeros((32, 1))
    ws = DeviceArray.zeros((lib.nm_tc_conv1_bwd_workspace(32),), np.uint8)
    gd = DeviceArray.from_numpy(to_bf16_bits(grid))
    lib.nm_tc_conv1_bwd_bf16(xd.p, gd.p, idx.p, dw.p, db.p, ws.p, ws.nbytes, n, 28, 28, 32, s)
    pos = nib & 3
    live = (nib >> 2).astype(bool)
    dconv = np.zeros((n, 32, 28, 28))
    for p4 in range(4):
        dconv[:, :, (p4 >> 1)::2, (p4 & 1)::2] = np.where(live & (pos == p4), dp, 0.0)
    _, dw_ref, db_ref = R.conv_layer_backward(x, w, None, dconv.astype(np.float32), padding="same", relu=False, mode="wired",
                                              dtype=np.float64)
    if integer:
        np.testing.assert_array_equal(dw.numpy(), dw_ref.astype(np.float32))
        np.testing.assert_array_equal(db.numpy().ravel(), db_ref.astype(np.float32).ravel())
    else:
        np.testing.assert_allclose(dw.numpy(), dw_ref, rtol=1e-4, atol=1e-3 * np.abs(dw_ref).max())
        np.testing.assert_allclose(db.numpy().ravel(), db_ref.ravel(), rtol=1e-4, atol=1e-3 * np.abs(db_ref).max())
    dw2 = DeviceArray.zeros((32, 1, 3, 3))                                    # deterministic
    lib.nm_tc_conv1_bwd_bf16(xd.p, gd.p, idx.p, dw2.p, db.p, ws.p, ws.nbytes, n, 28, 28, 32, s)
    np.testing.assert_array_equal(dw.numpy(), dw2.numpy())


def test_conv1_kernels_are_repeatable_at_batch_4096():
    """Regression (round 2): the conv1 builders handed their staged-input slot back (mbarrier arrive) right behind the
    shared-memory loads that read it; the arrive can overtake loads that have been issued but have not returned, and the
    bulk copy of a later tile then overwrote pixels still to be read -- about one launch in ten at batch 4096 produced
    ~30 wrong pool windows once the index math in between got cheaper.  The slot is now released only after the loaded
    values have been consumed into registers.  Many launches into fresh NaN-filled outputs must agree bit for bit."""
    from neuromah_b200._lib import lib
    from neuromah_b200.device import DeviceArray, stream
    n = 4096
    rng = np.random.default_rng(0)
    x = rng.random((n, 1, 28, 28), dtype=np.float32)
    w = (rng.standard_normal((32, 1, 3, 3)) * 0.3).astype(np.float32)
    wext = DeviceArray((128, 16), np.uint16)
    lib.nm_tc_pack_conv1_weights_bf16(DeviceArray.from_numpy(w).p, wext.p, 32, stream())
    xd = DeviceArray.from_numpy(x)
    dp = DeviceArray.from_numpy(to_bf16_bits(rng.standard_normal((n, 16, 16, 32)).astype(np.float32)))
    ws = DeviceArray((lib.nm_tc_conv1_bwd_workspace(32),), np.uint8)
    good = None
    for rep in range(24):
        y = DeviceArray.from_numpy(np.full((n, 16, 16, 32), 0x7FC0, np.uint16))
        idx = DeviceArray.from_numpy(np.full((n, 14, 14, 16), 0xEE, np.uint8))
        lib.nm_tc_conv1_fwd_pool_bf16(xd.p, wext.p, y.p, idx.p, n, 28, 28, 32, stream())
        dw, db = DeviceArray.zeros((32, 1, 3, 3)), DeviceArray.zeros((32,))
        lib.nm_tc_conv1_bwd_bf16(xd.p, dp.p, idx.p, dw.p, db.p, ws.p, ws.nbytes, n, 28, 28, 32, stream())
        got = (y.numpy()[:, 1:15, 1:15].copy(), idx.numpy().copy(), dw.numpy().copy(), db.numpy().copy())
        if good is None:
            good = got
            assert not (good[0] == 0x7FC0).any() and not (good[1] == 0xEE).any()      # every window was written
        else:
            for a, b, what in zip(got, good, ("y", "idx", "dw", "db")):
                np.testing.assert_array_equal(a, b, err_msg=f"launch {rep}: {what}")
