"""GPU parity for the general-shape bf16 tensor-core 3x3 convolution (nm_tc_conv_gen.cu: streamed filter bank,
channel counts padded to multiples of 64 -- the VGG-style stack of BASELINE.json config 4) through the C ABI.

Integer-valued data makes bf16 products and FP32 sums EXACT (|x| <= 2, |w| <= 1, at most 256*9 terms), so the
results must match the oracle bit for bit after the one bf16 rounding of the stored activation; a second case uses
real-valued data with the stated bf16 tolerance."""
import numpy as np
import pytest

from oracle import np_ref as R
from test_gpu_tc import bf16_round, from_bf16_bits, pad_nhwc, to_bf16_bits

pytestmark = pytest.mark.gpu


def _rup(a, b=64):
    return (a + b - 1) // b * b


def _data(rng, shape, integer, lo=-2, hi=3, scale=1.0):
    if integer:
        return rng.integers(lo, hi, shape).astype(np.float32)
    return bf16_round((rng.standard_normal(shape) * scale).astype(np.float32))


def _pack_weights(w, cp=None):
    from neuromah_b200._lib import lib
    from neuromah_b200.device import DeviceArray, stream
    oc, c = w.shape[:2]
    ocp, cp = _rup(oc), (cp or _rup(c))
    wf = DeviceArray((ocp, 9, cp), np.uint16)
    wd = DeviceArray((cp, 9, ocp), np.uint16)
    w_dev = DeviceArray.from_numpy(w)                  # keep the source alive until the (asynchronous) pack has been enqueued
    lib.nm_tc_pack_conv_weights_gen_bf16(w_dev.p, wf.p, wd.p, oc, c, ocp, cp, stream())
    wf.numpy()                                         # synchronises: w_dev may go
    return wf, wd, ocp, cp


def _pad_channels(x_nchw, cp):
    n, c, h, w = x_nchw.shape
    out = np.zeros((n, cp, h, w), np.float32)
    out[:, :c] = x_nchw
    return out


def _halo_is_zero(a):
    return not (a[:, 0].any() or a[:, -1].any() or a[:, :, 0].any() or a[:, :, -1].any())


FPROP_CASES = [  # n, h, w, c, oc, integer
    (3, 8, 8, 128, 256, True),      # two K chunks, two N tiles of 128
    (2, 32, 32, 64, 64, True),      # N tile of 64, activation tile split over two TMA boxes (Wp = 34)
    (5, 16, 16, 64, 128, True),
    (2, 8, 8, 256, 256, True),      # four K chunks
    (7, 16, 16, 128, 128, False),
    (1, 8, 8, 64, 64, True),        # fewer rows than one 256-pixel tile
]


@pytest.mark.parametrize("n,h,w,c,oc,integer", FPROP_CASES)
def test_tc_gen_fprop_bias_relu(n, h, w, c, oc, integer):
    from neuromah_b200._lib import lib
    from neuromah_b200.device import DeviceArray, stream
    rng = np.random.default_rng(100 + n + c + oc)
    x = _data(rng, (n, c, h, w), integer)
    wt = _data(rng, (oc, c, 3, 3), integer, -1, 2, 0.05)
    wf, _, ocp, cp = _pack_weights(wt)
    np.testing.assert_array_equal(from_bf16_bits(wf.numpy()), bf16_round(wt).transpose(0, 2, 3, 1).reshape(oc, 9, c))
    xp = DeviceArray.from_numpy(pad_nhwc(x))
    y = DeviceArray.from_numpy(np.full((n, h + 2, w + 2, oc), 0x7FC0, np.uint16))       # NaN-filled: every element must be written
    lib.nm_tc_conv3x3_gen_bf16(xp.p, wf.p, None, None, y.p, n, h, w, c, oc, 1, stream())
    ref, _ = R.conv_layer_forward(x, wt, None, padding="same", relu=True, dtype=np.float64)
    got = y.numpy()
    assert _halo_is_zero(got)
    gi = from_bf16_bits(got[:, 1:-1, 1:-1, :]).transpose(0, 3, 1, 2)
    if integer:
        np.testing.assert_array_equal(gi, bf16_round(ref.astype(np.float32)))
    else:
        np.testing.assert_allclose(gi, ref, rtol=2 ** -7, atol=2e-3)
    # bias switch (config.conv_bias), no ReLU
    b = _data(rng, (oc,), integer)
    b_dev = DeviceArray.from_numpy(b)
    lib.nm_tc_conv3x3_gen_bf16(xp.p, wf.p, b_dev.p, None, y.p, n, h, w, c, oc, 0, stream())
    ref2, _ = R.conv_layer_forward(x, wt, None, padding="same", relu=False, dtype=np.float64)
    ref2 = ref2 + b.reshape(1, oc, 1, 1)
    got = y.numpy()
    assert _halo_is_zero(got)
    gi = from_bf16_bits(got[:, 1:-1, 1:-1, :]).transpose(0, 3, 1, 2)
    if integer:
        np.testing.assert_array_equal(gi, bf16_round(ref2.astype(np.float32)))
    else:
        np.testing.assert_allclose(gi, ref2, rtol=2 ** -7, atol=2e-3)


@pytest.mark.parametrize("cpad,u8", [(64, False), (32, False), (32, True)])
def test_tc_gen_first_layer_from_nchw_input(cpad, u8):
    """C(3 -> 64) of the VGG-style stack: f32 (or raw uint8) NCHW input packed to a haloed, channel-padded bf16 tensor on
    the device; input channels padded to 64 (128-byte K chunks) or to 32 (the 64-byte K-chunk variant the plan uses)."""
    from neuromah_b200._lib import lib
    from neuromah_b200.device import DeviceArray, stream
    rng = np.random.default_rng(7)
    n, c, h, w, oc = 3, 3, 32, 32, 64
    x = _data(rng, (n, c, h, w), True, 0, 4)
    wt = _data(rng, (oc, c, 3, 3), True, -1, 2)
    wf, _, ocp, cp = _pack_weights(wt, cpad)
    xp = DeviceArray.from_numpy(np.full((n, h + 2, w + 2, cp), 0x7FC0, np.uint16))
    if u8:
        x_dev = DeviceArray.from_numpy(x.astype(np.uint8), dtype=np.uint8)
        lib.nm_tc_pack_input_nhwc_u8_bf16(x_dev.p, 1.0, xp.p, n, c, h, w, cp, stream())
    else:
        x_dev = DeviceArray.from_numpy(x)
        lib.nm_tc_pack_input_nhwc_bf16(x_dev.p, xp.p, n, c, h, w, cp, stream())
    np.testing.assert_array_equal(xp.numpy(), pad_nhwc(_pad_channels(x, cp)))
    y = DeviceArray.zeros((n, h + 2, w + 2, oc), np.uint16)
    lib.nm_tc_conv3x3_gen_bf16(xp.p, wf.p, None, None, y.p, n, h, w, cp, oc, 1, stream())
    ref, _ = R.conv_layer_forward(x, wt, None, padding="same", relu=True, dtype=np.float64)
    np.testing.assert_array_equal(from_bf16_bits(y.numpy()[:, 1:-1, 1:-1, :]).transpose(0, 3, 1, 2), bf16_round(ref.astype(np.float32)))


@pytest.mark.parametrize("n,h,w,c,oc,integer", [(3, 8, 8, 128, 256, True), (2, 32, 32, 64, 64, True), (4, 16, 16, 128, 64, True),
                                                (6, 16, 16, 64, 128, False)])
def test_tc_gen_dgrad_with_relu_mask(n, h, w, c, oc, integer):
    """dX = col2im(W^T dY) (Conv2D.cpp:218-222) then the ReLU backward of the layer below (Relu.py:12-14), one launch."""
    from neuromah_b200._lib import lib
    from neuromah_b200.device import DeviceArray, stream
    rng = np.random.default_rng(200 + n + c + oc)
    x = _data(rng, (n, c, h, w), integer)
    wt = _data(rng, (oc, c, 3, 3), integer, -1, 2, 0.05)
    dy = _data(rng, (n, oc, h, w), integer)
    act = np.maximum(_data(rng, (n, c, h, w), integer), 0)                       # the layer below's post-ReLU output
    _, wd, ocp, cp = _pack_weights(wt)
    dyp, actp = DeviceArray.from_numpy(pad_nhwc(dy)), DeviceArray.from_numpy(pad_nhwc(act))
    dx = DeviceArray.from_numpy(np.full((n, h + 2, w + 2, c), 0x7FC0, np.uint16))
    lib.nm_tc_conv3x3_gen_bf16(dyp.p, wd.p, None, None, dx.p, n, h, w, oc, c, 0, stream())
    ref, _, _ = R.conv_layer_backward(x, wt, None, dy, padding="same", relu=False, mode="wired", dtype=np.float64)
    got = dx.numpy()
    assert _halo_is_zero(got)
    gi = from_bf16_bits(got[:, 1:-1, 1:-1, :]).transpose(0, 3, 1, 2)
    if integer:
        np.testing.assert_array_equal(gi, bf16_round(ref.astype(np.float32)))
    else:
        np.testing.assert_allclose(gi, ref, rtol=2 ** -7, atol=2e-2)
    lib.nm_tc_conv3x3_gen_bf16(dyp.p, wd.p, None, actp.p, dx.p, n, h, w, oc, c, 0, stream())
    refm = R.relu_backward(ref, act)
    got = dx.numpy()
    assert _halo_is_zero(got)
    gi = from_bf16_bits(got[:, 1:-1, 1:-1, :]).transpose(0, 3, 1, 2)
    if integer:
        np.testing.assert_array_equal(gi, bf16_round(refm.astype(np.float32)))
    else:
        np.testing.assert_allclose(gi, refm, rtol=2 ** -7, atol=2e-2)


@pytest.mark.parametrize("n,h,w,c,oc,integer,cpad", [(3, 8, 8, 128, 256, True, None), (2, 32, 32, 64, 64, True, None),
                                                     (3, 32, 32, 3, 64, True, None), (3, 32, 32, 3, 64, True, 32),
                                                     (9, 16, 16, 64, 128, True, None), (5, 8, 8, 256, 256, False, None),
                                                     (1, 8, 8, 64, 64, True, None), (40, 16, 16, 3, 128, True, 32)])
def test_tc_gen_wgrad_and_bias_gradient(n, h, w, c, oc, integer, cpad):
    """dW and the bias gradient (column sums of dY formed by an extra MMA against a tile of ones inside the wgrad kernel,
    spread over the CTAs that share a pixel range): bit-exact on integer data for one to eight input-channel chunks,
    fewer pixel blocks than CTAs sharing them (n = 1), and input channels padded to 32."""
    from neuromah_b200._lib import lib
    from neuromah_b200.device import DeviceArray, stream
    rng = np.random.default_rng(300 + n + c + oc)
    x = _data(rng, (n, c, h, w), integer)
    wt = _data(rng, (oc, c, 3, 3), integer, -1, 2, 0.05)
    dy = _data(rng, (n, oc, h, w), integer)
    ocp, cp = _rup(oc), (cpad or _rup(c))
    xp, dyp = DeviceArray.from_numpy(pad_nhwc(_pad_channels(x, cp))), DeviceArray.from_numpy(pad_nhwc(dy))
    dw = DeviceArray.from_numpy(np.full((oc, c, 3, 3), np.nan, np.float32))
    db = DeviceArray.from_numpy(np.full((oc,), np.nan, np.float32))
    ws = DeviceArray((lib.nm_tc_conv3x3_wgrad_gen_workspace(n, h, w, c, cp, ocp),), np.uint8)
    lib.nm_tc_conv3x3_wgrad_gen_bf16(xp.p, dyp.p, dw.p, db.p, ws.p, ws.nbytes, n, h, w, c, oc, cp, ocp, stream())
    _, ref, refb = R.conv_layer_backward(x, wt, None, dy, padding="same", relu=False, mode="wired", dtype=np.float64)
    if integer:
        np.testing.assert_array_equal(dw.numpy(), ref.astype(np.float32))
        np.testing.assert_array_equal(db.numpy(), refb.reshape(-1).astype(np.float32))
    else:
        np.testing.assert_allclose(dw.numpy(), ref, rtol=1e-4, atol=1e-2)
        np.testing.assert_allclose(db.numpy(), refb.reshape(-1), rtol=1e-4, atol=1e-2)
    dw2 = DeviceArray.zeros((oc, c, 3, 3))                                       # deterministic: bitwise repeatable
    lib.nm_tc_conv3x3_wgrad_gen_bf16(xp.p, dyp.p, dw2.p, None, ws.p, ws.nbytes, n, h, w, c, oc, cp, ocp, stream())
    np.testing.assert_array_equal(dw.numpy(), dw2.numpy())


def test_tc_gen_rejects_unpadded_channel_counts():
    from neuromah_b200._lib import lib
    from neuromah_b200.device import DeviceArray, stream
    a = DeviceArray.zeros((1024,), np.uint16)
    with pytest.raises(ValueError):
        lib.nm_tc_conv3x3_gen_bf16(a.p, a.p, None, None, a.p, 1, 8, 8, 48, 64, 1, stream())
    with pytest.raises(ValueError):
        lib.nm_tc_conv3x3_wgrad_gen_bf16(a.p, a.p, a.p, None, a.p, 16, 1, 8, 8, 64, 64, 64, 96, stream())
