"""GPU parity of the FP32-exact tensor-core convolution kernels (mode "bf16x3": nm_tc_conv3x3_gen_x3,
nm_tc_conv3x3_wgrad_gen_x3, nm_tc_maxpool2x2_fwd_x3, the plane packs and the NHWC-planes <-> NCHW-float32 transposes)
through the C ABI, on real-valued float32 data against the float64 oracle.

Tolerance: 8e-6 (measured <= 2.6e-6) of the largest output magnitude per tensor (FP32 accumulation order; the stated tolerance of the mode is
rtol 1e-4 / atol 1e-5 at the model level, tests/test_gpu_vgg_tc_x3_plan.py).  Planes are allocated for a LARGER batch than
the one in flight and NaN-filled past it: nothing outside the current batch may be read or written.
Reference arithmetic: Conv2D.cpp:83-146 / :150-240, Conv_wrapepr.py:78-109, maxpooling_backend.cpp:94-108."""
import ctypes as C

import numpy as np
import pytest

from oracle import np_ref as R
from test_gpu_tc import from_bf16_bits, to_bf16_bits, unpack_nibbles

pytestmark = pytest.mark.gpu
NAN = 0x7FC0


def _rup(a, b=64):
    return (a + b - 1) // b * b


def split3_host(a):
    """float32 -> the three bf16 planes (as float32) the device forms: h = bf16(v), m = bf16(v - h), l = bf16(v - h - m)."""
    a = np.ascontiguousarray(a, np.float32)
    h = from_bf16_bits(to_bf16_bits(a))
    r = a - h
    m = from_bf16_bits(to_bf16_bits(r))
    l = from_bf16_bits(to_bf16_bits(r - m))
    return h, m, l


def planes_nhwc(x_nchw, cp, cap):
    """NCHW float32 -> bf16 bit planes [3][cap][H+2][W+2][cp] (zero halo, zero padded channels, NaN past the batch)."""
    n, c, h, w = x_nchw.shape
    out = np.full((3, cap, h + 2, w + 2, cp), NAN, np.uint16)
    for pl, part in enumerate(split3_host(x_nchw)):
        t = np.zeros((n, h + 2, w + 2, cp), np.float32)
        t[:, 1:-1, 1:-1, :c] = part.transpose(0, 2, 3, 1)
        out[pl, :n] = to_bf16_bits(t)
    return out


def merge(planes_bits):
    p = from_bf16_bits(planes_bits).astype(np.float64)
    return p[0] + p[1] + p[2]


def interior_nchw(planes_bits, n, c):
    return merge(planes_bits[:, :n, 1:-1, 1:-1, :c]).transpose(0, 3, 1, 2)


def halo_zero_and_tail_untouched(bits, n):
    live = bits[:, :n]
    ok = not (live[:, :, 0].any() or live[:, :, -1].any() or live[:, :, :, 0].any() or live[:, :, :, -1].any())
    return ok and (bits[:, n:] == NAN).all()


def pack_weights_x3(w, cp=None):
    from neuromah_b200._lib import lib
    from neuromah_b200.device import DeviceArray, stream
    oc, c = w.shape[:2]
    ocp, cp = _rup(oc), (cp or _rup(c))
    wf, wd = DeviceArray((3, ocp, 9, cp), np.uint16), DeviceArray((3, cp, 9, ocp), np.uint16)
    w_dev = DeviceArray.from_numpy(w)
    lib.nm_tc_pack_conv_weights_gen_x3(w_dev.p, wf.p, wd.p, oc, c, ocp, cp, stream())
    wf.numpy()
    return wf, wd, ocp, cp


CASES = [  # n, h, w, c, oc
    (3, 8, 8, 128, 256),       # two K chunks, two N tiles of 128
    (2, 32, 32, 64, 64),       # N tile of 64 (a filter row per ring slot), activation tile split over two TMA boxes
    (5, 16, 16, 64, 128),
    (1, 8, 8, 64, 64),         # fewer rows than one tile
    (4, 16, 16, 40, 24),       # widths that are zero-padded to 64
]


@pytest.mark.parametrize("n,h,w,c,oc", CASES)
def test_x3_conv_fprop_dgrad_wgrad(n, h, w, c, oc):
    from neuromah_b200._lib import lib
    from neuromah_b200.device import DeviceArray, stream
    rng = np.random.default_rng(500 + n + c + oc)
    cap = n + 2
    x = rng.standard_normal((n, c, h, w)).astype(np.float32)
    wt = (rng.standard_normal((oc, c, 3, 3)) * 0.05).astype(np.float32)
    b = rng.standard_normal((oc,)).astype(np.float32)
    dy = rng.standard_normal((n, oc, h, w)).astype(np.float32)
    wf, wd, ocp, cp = pack_weights_x3(wt)
    assert np.abs(merge(wf.numpy())[:oc, :, :c] - wt.transpose(0, 2, 3, 1).reshape(oc, 9, c)).max() <= 2.0 ** -24 * np.abs(wt).max()
    assert not merge(wf.numpy())[oc:].any() and not merge(wf.numpy())[:, :, c:].any()
    xs, ys = cap * (h + 2) * (w + 2) * cp, cap * (h + 2) * (w + 2) * ocp
    xp = DeviceArray.from_numpy(planes_nhwc(x, cp, cap))
    # ---- fprop + ReLU (Q4: no bias), then the bias switch without ReLU
    y = DeviceArray.from_numpy(np.full((3, cap, h + 2, w + 2, ocp), NAN, np.uint16))
    lib.nm_tc_conv3x3_gen_x3(xp.p, xs, wf.p, None, None, y.p, ys, n, h, w, cp, ocp, 1, stream())
    ref, _ = R.conv_layer_forward(x, wt, None, padding="same", relu=True, dtype=np.float64)
    got = y.numpy()
    assert halo_zero_and_tail_untouched(got, n)
    e = np.abs(interior_nchw(got, n, oc) - ref).max() / np.abs(ref).max()
    print("fprop", e)
    assert e <= 8e-6
    assert not got[:, :n, :, :, oc:].any()                                 # padded output channels stay zero
    bp = np.zeros(ocp, np.float32); bp[:oc] = b
    lib.nm_tc_conv3x3_gen_x3(xp.p, xs, wf.p, DeviceArray.from_numpy(bp).p, None, y.p, ys, n, h, w, cp, ocp, 0, stream())
    ref2, _ = R.conv_layer_forward(x, wt, None, padding="same", relu=False, dtype=np.float64)
    ref2 = ref2 + b.reshape(1, oc, 1, 1)
    got = y.numpy()
    assert halo_zero_and_tail_untouched(got, n)
    assert np.abs(interior_nchw(got, n, oc) - ref2).max() <= 4e-6 * np.abs(ref2).max()
    # ---- dgrad + the ReLU mask of the layer below (the h plane of its activation)
    act = np.maximum(rng.standard_normal((n, c, h, w)), 0).astype(np.float32)
    dyp, actp = DeviceArray.from_numpy(planes_nhwc(dy, ocp, cap)), DeviceArray.from_numpy(planes_nhwc(act, cp, cap))
    dx = DeviceArray.from_numpy(np.full((3, cap, h + 2, w + 2, cp), NAN, np.uint16))
    lib.nm_tc_conv3x3_gen_x3(dyp.p, ys, wd.p, None, actp.p, dx.p, xs, n, h, w, ocp, cp, 0, stream())
    refd, refw, refb = R.conv_layer_backward(x, wt, None, dy, padding="same", relu=False, mode="wired", dtype=np.float64)
    refm = np.where(from_bf16_bits(to_bf16_bits(act)) > 0, refd, 0.0)
    got = dx.numpy()
    assert halo_zero_and_tail_untouched(got, n)
    e = np.abs(interior_nchw(got, n, c) - refm).max() / np.abs(refm).max()
    print("dgrad", e)
    assert e <= 8e-6
    # ---- wgrad + bias gradient
    dw = DeviceArray.from_numpy(np.full((oc, c, 3, 3), np.nan, np.float32))
    db = DeviceArray.from_numpy(np.full((oc,), np.nan, np.float32))
    ws = DeviceArray((lib.nm_tc_conv3x3_wgrad_gen_workspace(n, h, w, c, cp, ocp),), np.uint8)
    lib.nm_tc_conv3x3_wgrad_gen_x3(xp.p, xs, dyp.p, ys, dw.p, db.p, ws.p, ws.nbytes, n, h, w, c, oc, cp, ocp, stream())
    e = np.abs(dw.numpy() - refw).max() / np.abs(refw).max()
    eb = np.abs(db.numpy() - refb.reshape(-1)).max() / max(np.abs(refb).max(), np.sqrt(n * h * w))
    print("wgrad", e, "dbias", eb)
    assert e <= 4e-6 and eb <= 4e-6
    dw2 = DeviceArray.zeros((oc, c, 3, 3))
    lib.nm_tc_conv3x3_wgrad_gen_x3(xp.p, xs, dyp.p, ys, dw2.p, None, ws.p, ws.nbytes, n, h, w, c, oc, cp, ocp, stream())
    np.testing.assert_array_equal(dw.numpy(), dw2.numpy())                 # deterministic


@pytest.mark.parametrize("u8", [False, True])
def test_x3_first_layer_from_nchw_input(u8):
    """C(3 -> 64) with the input channels padded to 32: float32 (or raw uint8 / 255) NCHW input packed into planes on the device."""
    from neuromah_b200._lib import lib
    from neuromah_b200.device import DeviceArray, stream
    rng = np.random.default_rng(7)
    n, c, h, w, oc, cp, cap = 3, 3, 32, 32, 64, 32, 4
    wt = (rng.standard_normal((oc, c, 3, 3)) * 0.2).astype(np.float32)
    wf, _, ocp, _ = pack_weights_x3(wt, cp)
    xs = cap * (h + 2) * (w + 2) * cp
    xp = DeviceArray.from_numpy(np.full((3, cap, h + 2, w + 2, cp), NAN, np.uint16))
    if u8:
        xu = rng.integers(0, 256, (n, c, h, w)).astype(np.uint8)
        x = xu.astype(np.float32) / np.float32(255.0)
        lib.nm_tc_pack_input_nhwc_u8_x3(DeviceArray.from_numpy(xu, dtype=np.uint8).p, 255.0, xp.p, xs, n, c, h, w, cp, stream())
    else:
        x = rng.standard_normal((n, c, h, w)).astype(np.float32)
        lib.nm_tc_pack_input_nhwc_x3(DeviceArray.from_numpy(x).p, xp.p, xs, n, c, h, w, cp, stream())
    np.testing.assert_array_equal(xp.numpy(), planes_nhwc(x, cp, cap))
    ys = cap * (h + 2) * (w + 2) * ocp
    y = DeviceArray.from_numpy(np.full((3, cap, h + 2, w + 2, ocp), NAN, np.uint16))
    lib.nm_tc_conv3x3_gen_x3(xp.p, xs, wf.p, None, None, y.p, ys, n, h, w, cp, ocp, 1, stream())
    ref, _ = R.conv_layer_forward(x, wt, None, padding="same", relu=True, dtype=np.float64)
    assert np.abs(interior_nchw(y.numpy(), n, oc) - ref).max() <= 4e-6 * np.abs(ref).max()


@pytest.mark.parametrize("padded", [0, 1])
def test_x3_maxpool_forward_backward_and_head_transposes(padded):
    from neuromah_b200._lib import lib
    from neuromah_b200.device import DeviceArray, stream
    rng = np.random.default_rng(11 + padded)
    n, c, h, w, cap = 3, 64, 8, 8, 5
    x = np.maximum(rng.standard_normal((n, c, h, w)), 0).astype(np.float32)       # post-ReLU: zeros tie, the first one must win
    x[0, :, 0:2, 0:2] = x[0, :, 0:1, 0:1]                                         # a window of four equal values
    xs = cap * (h + 2) * (w + 2) * c
    xp = DeviceArray.from_numpy(planes_nhwc(x, c, cap))
    oh, ow = (h // 2 + 2, w // 2 + 2) if padded else (h // 2, w // 2)
    ys = cap * oh * ow * c
    y = DeviceArray.from_numpy(np.full((3, cap, oh, ow, c), NAN, np.uint16))
    idx = DeviceArray.zeros((cap, h // 2, w // 2, c // 2), np.uint8)
    lib.nm_tc_maxpool2x2_fwd_x3(xp.p, xs, y.p, ys, idx.p, n, h, w, c, padded, stream())
    xe = interior_nchw(planes_nhwc(x, c, cap), n, c)                              # the values the planes carry (== x to 2^-24)
    assert np.abs(xe - x).max() <= 2.0 ** -24 * np.abs(x).max()
    ref = xe.reshape(n, c, h // 2, 2, w // 2, 2).max(axis=(3, 5))                 # MaxPooling2D(2, 2), 'valid'
    np.testing.assert_array_equal(ref.astype(np.float32), R.maxpool2d_forward_cpu(xe.astype(np.float32), 2, 2, 2, 2, "valid")[0])
    got = y.numpy()
    inner = got[:, :n, 1:-1, 1:-1] if padded else got[:, :n]
    np.testing.assert_array_equal(merge(inner).transpose(0, 3, 1, 2), ref)         # exact: the winner's planes are copied
    assert (got[:, n:] == NAN).all()
    nib = unpack_nibbles(idx.numpy()[:n]).transpose(0, 3, 1, 2)
    xw = xe.reshape(n, c, h // 2, 2, w // 2, 2).transpose(0, 1, 2, 4, 3, 5).reshape(n, c, h // 2, w // 2, 4)
    np.testing.assert_array_equal(nib & 3, xw.argmax(-1))                          # first maximum in the row-major window scan
    np.testing.assert_array_equal((nib >> 2) & 1, (ref > 0).astype(np.uint8))
    # backward: nm_tc_maxpool2x2_bwd_bf16 once per plane routes the planes of the pooled gradient
    dp = rng.standard_normal((n, c, h // 2, w // 2)).astype(np.float32)
    dpl = np.full((3, cap, oh, ow, c), NAN, np.uint16)
    for pl, part in enumerate(split3_host(dp)):
        t = np.zeros((n, oh, ow, c), np.float32)
        if padded:
            t[:, 1:-1, 1:-1] = part.transpose(0, 2, 3, 1)
        else:
            t[:] = part.transpose(0, 2, 3, 1)
        dpl[pl, :n] = to_bf16_bits(t)
    dpool = DeviceArray.from_numpy(dpl)
    dyv = DeviceArray.from_numpy(np.full((3, cap, h + 2, w + 2, c), NAN, np.uint16))
    for pl in range(3):
        lib.nm_tc_maxpool2x2_bwd_bf16(C.c_void_p(dpool.ptr + 2 * pl * ys), idx.p, C.c_void_p(dyv.ptr + 2 * pl * xs), n, h, w, c, padded, stream())
    want = np.zeros((n, c, h // 2, 2, w // 2, 2))
    pos = xw.argmax(-1)
    routed = np.where(ref > 0, merge(dpl[:, :n, 1:-1, 1:-1] if padded else dpl[:, :n]).transpose(0, 3, 1, 2), 0.0)
    for q in range(4):
        want[:, :, :, q // 2, :, q % 2] = np.where(pos == q, routed, 0.0)
    got = dyv.numpy()
    assert halo_zero_and_tail_untouched(got, n)
    np.testing.assert_array_equal(interior_nchw(got, n, c), want.reshape(n, c, h, w))
    if not padded:
        # planes [3][N][HW][C] -> float32 [N][C*HW] (Flatten's order) and back
        hw = (h // 2) * (w // 2)
        flat = DeviceArray.from_numpy(np.full((n, c * hw), np.nan, np.float32))
        lib.nm_tc_merge3_nhwc_to_nchw_f32(y.p, ys, flat.p, n, hw, c, c, stream())
        np.testing.assert_array_equal(flat.numpy(), ref.reshape(n, -1).astype(np.float32))
        g = rng.standard_normal((n, c * hw)).astype(np.float32)
        back = DeviceArray.from_numpy(np.full((3, cap, hw, c), NAN, np.uint16))
        lib.nm_tc_split3_nchw_to_nhwc(DeviceArray.from_numpy(g).p, back.p, cap * hw * c, n, hw, c, c, stream())
        gb = back.numpy()
        assert (gb[:, n:] == NAN).all()
        for pl, part in enumerate(split3_host(g.reshape(n, c, hw))):
            np.testing.assert_array_equal(gb[pl, :n], to_bf16_bits(part.transpose(0, 2, 1)))
        # a last conv narrower than its padded tensor: only the first cr channels are real
        cr = 40
        flat2 = DeviceArray.from_numpy(np.full((n, cr * hw), np.nan, np.float32))
        lib.nm_tc_merge3_nhwc_to_nchw_f32(y.p, ys, flat2.p, n, hw, cr, c, stream())
        np.testing.assert_array_equal(flat2.numpy(), ref[:, :cr].reshape(n, -1).astype(np.float32))
        g2 = rng.standard_normal((n, cr * hw)).astype(np.float32)
        lib.nm_tc_split3_nchw_to_nhwc(DeviceArray.from_numpy(g2).p, back.p, cap * hw * c, n, hw, cr, c, stream())
        gb = back.numpy()
        assert not gb[:, :n, :, cr:].any()
        np.testing.assert_array_equal(gb[0, :n, :, :cr], to_bf16_bits(g2.reshape(n, cr, hw).transpose(0, 2, 1)))
