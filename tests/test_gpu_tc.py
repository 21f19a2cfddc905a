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
