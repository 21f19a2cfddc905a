"""CPU check of the DATA-LAYOUT CONTRACT of the tensor-core conv kernels (no GPU, no product code executed):
a NumPy restatement of the addressing the kernels use -- zero-haloed NHWC grid flattened to rows, a 3x3 tap = a flat row
shift of r*Wp + s, the packed operand orders [OC][tap][C] (fprop) and [C][8-tap][OC] (dgrad), the wgrad shift, the
pool nibble -- against the oracle's conv / pool.  It pins the index algebra documented in DESIGN.md section 4 and
include/neuromah_b200.h (nm_tc_conv3x3_gen_bf16, nm_tc_conv3x3_wgrad_gen_bf16, nm_tc_maxpool2x2_*_bf16): if this
algebra were wrong no kernel following it could match the oracle, and a maintainer changing a layout can re-derive the
expected operand here before touching CUDA."""
import numpy as np
import pytest

from oracle import np_ref as R


def halo_nhwc(x_nchw):
    n, c, h, w = x_nchw.shape
    out = np.zeros((n, h + 2, w + 2, c))
    out[:, 1:-1, 1:-1, :] = x_nchw.transpose(0, 2, 3, 1)
    return out


def pack_fprop(w):      # OIHW -> [OC][tap][C]                      (pack_conv_weights_gen_kernel)
    oc, c = w.shape[:2]
    return w.reshape(oc, c, 9).transpose(0, 2, 1).copy()


def pack_dgrad(w):      # OIHW -> [C][8 - tap][OC]
    oc, c = w.shape[:2]
    return w.reshape(oc, c, 9)[:, :, ::-1].transpose(1, 2, 0).copy()


def flat_conv(x_pad, w_packed, h, w):
    """E[p] = sum_tap X[p - (Wp+1) + r*Wp + s] . B[tap]^T over the flattened haloed grid; halo rows of the result zeroed."""
    n, hp, wp, c = x_pad.shape
    rows = n * hp * wp
    x = np.zeros((rows + 2 * (wp + 1), c))
    x[wp + 1:wp + 1 + rows] = x_pad.reshape(rows, c)          # rows before / after the tensor read as zero (TMA fill)
    out = np.zeros((rows, w_packed.shape[0]))
    for tap in range(9):
        r, s = divmod(tap, 3)
        out += x[r * wp + s:r * wp + s + rows] @ w_packed[:, tap, :].T
    out = out.reshape(n, hp, wp, -1)
    out[:, 0] = out[:, -1] = 0
    out[:, :, 0] = out[:, :, -1] = 0
    return out


@pytest.mark.parametrize("n,c,oc,h,w", [(2, 3, 4, 6, 6), (1, 5, 2, 4, 8), (3, 2, 3, 8, 4)])
def test_flat_shift_fprop_dgrad_wgrad_equal_the_oracle(n, c, oc, h, w):
    rng = np.random.default_rng(n * 100 + c * 10 + oc)
    x = rng.standard_normal((n, c, h, w))
    wt = rng.standard_normal((oc, c, 3, 3))
    dy = rng.standard_normal((n, oc, h, w))
    ref_y, _ = R.conv_layer_forward(x, wt, None, padding="same", relu=False, dtype=np.float64)
    ref_dx, ref_dw, ref_db = R.conv_layer_backward(x, wt, None, dy, padding="same", relu=False, mode="wired", dtype=np.float64)
    xp, dyp = halo_nhwc(x), halo_nhwc(dy)
    # fprop: y_pad = flat_conv(x_pad, [OC][tap][C])
    y = flat_conv(xp, pack_fprop(wt), h, w)
    np.testing.assert_allclose(y[:, 1:-1, 1:-1, :].transpose(0, 3, 1, 2), ref_y, rtol=1e-10, atol=1e-10)
    # dgrad: the SAME routine on dY with the flipped / transposed operand
    dx = flat_conv(dyp, pack_dgrad(wt), h, w)
    np.testing.assert_allclose(dx[:, 1:-1, 1:-1, :].transpose(0, 3, 1, 2), ref_dx, rtol=1e-10, atol=1e-10)
    # wgrad: dW[oc, tap, c] = sum_p dY[p, oc] * X[p + r*Wp + s - (Wp+1), c]; bias gradient = column sums of the haloed dY
    hp, wp = h + 2, w + 2
    rows = n * hp * wp
    xz = np.zeros((rows + 2 * (wp + 1), c))
    xz[wp + 1:wp + 1 + rows] = xp.reshape(rows, c)
    dyf = dyp.reshape(rows, oc)
    dw = np.stack([dyf.T @ xz[r * wp + s:r * wp + s + rows] for r in range(3) for s in range(3)], axis=1)   # [oc][tap][c]
    np.testing.assert_allclose(dw.transpose(0, 2, 1).reshape(oc, c, 3, 3), ref_dw, rtol=1e-10, atol=1e-10)
    np.testing.assert_allclose(dyf.sum(axis=0), ref_db.reshape(-1), rtol=1e-10, atol=1e-10)


def test_channel_padding_is_neutral():
    """Zero-padded input channels (first layer: 3 -> 64) and their zero weights change nothing."""
    rng = np.random.default_rng(5)
    x, wt = rng.standard_normal((2, 3, 4, 4)), rng.standard_normal((4, 3, 3, 3))
    xpad = np.concatenate([x, np.zeros((2, 5, 4, 4))], axis=1)
    wpad = np.concatenate([wt, np.zeros((4, 5, 3, 3))], axis=1)
    a = flat_conv(halo_nhwc(x), pack_fprop(wt), 4, 4)
    b = flat_conv(halo_nhwc(xpad), pack_fprop(wpad), 4, 4)
    np.testing.assert_array_equal(a, b)


def test_pool_nibble_is_the_reference_arg_max():
    """nibble = first-max position (dy*2 + dx, row-major window scan, strict >) | (max > 0) << 2, and routing the pooled
    gradient by it (masked by bit 2) is MaxPooling2D.backward followed by the ReLU backward of the conv below."""
    rng = np.random.default_rng(9)
    pre = rng.integers(-2, 3, (3, 4, 6, 8)).astype(np.float64)          # integer data: ties everywhere
    act = np.maximum(pre, 0)
    pooled, pidx = R.maxpool2d_forward_cpu(act.astype(np.float32), 2, 2, 2, 2, "valid")
    win = act.reshape(3, 4, 3, 2, 4, 2).transpose(0, 1, 2, 4, 3, 5).reshape(3, 4, 3, 4, 4)      # [n][c][ph][pw][dy*2+dx]
    pos = np.zeros(win.shape[:-1], np.int64)
    best = win[..., 0].copy()
    for q in range(1, 4):
        better = win[..., q] > best
        pos[better], best[better] = q, win[..., q][better]
    np.testing.assert_array_equal(best, pooled)
    np.testing.assert_array_equal(pos, (pidx[..., 0] % 2) * 2 + (pidx[..., 1] % 2))
    nib = pos | ((best > 0).astype(np.int64) << 2)
    dpool = rng.integers(-3, 4, pooled.shape).astype(np.float64)
    ref = R.relu_backward(R.maxpool2d_backward_cpu(dpool.astype(np.float32), pidx, act.astype(np.float32), 2, 2, 2, 2, "valid"), pre)
    dy = np.zeros_like(act)
    for q in range(4):
        sel = nib == (q | 4)
        dy[:, :, q // 2::2, q % 2::2] = np.where(sel, dpool, 0.0)
    np.testing.assert_array_equal(dy, ref)


def test_config4_stack_matches_the_survey_figures():
    """the VGG-style stack of config 4 has 1 186 378 parameters and 913.29 MFLOP per sample
    (fwd 305.61 + bwd 607.68, the first conv's dgrad excluded) -- the figures DESIGN.md's TFLOP/s are quoted on."""
    from oracle import ref_net
    spec = ref_net.vgg_cifar_spec()
    params = ref_net.init_params(spec, np.random.RandomState(0), init="he")
    assert sum(p["weights"].size + p["biases"].size for p in params if p is not None) == 1_186_378
    h = w = 32
    fwd = bwd = 0.0
    first = True
    for L in spec:
        if L["kind"] == "conv":
            f = 2.0 * L["oc"] * L["ic"] * 9 * h * w
            fwd += f
            bwd += f if first else 2 * f            # wgrad always, dgrad except for the first layer
            first = False
        elif L["kind"] == "pool":
            h, w = h // 2, w // 2
        elif L["kind"] == "dense":
            f = 2.0 * L["n_in"] * L["n_out"]
            fwd += f
            bwd += 2 * f
    assert abs(fwd / 1e6 - 305.61) < 0.01 and abs(bwd / 1e6 - 607.68) < 0.01 and abs((fwd + bwd) / 1e6 - 913.29) < 0.01


def test_uint8_pixel_normalisation_is_exact_after_bf16_rounding():
    """The bf16 plans read raw uint8 pixels and compute float(u) * (1/255) on the device (nm_tc_conv1_*_u8_bf16,
    nm_tc_pack_input_nhwc_u8_bf16).  The reference's preprocessing is the float32 DIVISION u / 255.0
    (examples/test_Conv2D_MNIST.py:27); the two differ in the last float32 bit for 126 of the 256 pixel values, but
    never after rounding to the bf16 MMA operand -- so uint8 input is bit-identical to float32 input in that mode."""
    u = np.arange(256, dtype=np.float32)
    div = (u / np.float32(255.0)).astype(np.float32)
    mul = (u * (np.float32(1.0) / np.float32(255.0))).astype(np.float32)

    def bf16_rn(x):
        i = x.view(np.uint32).astype(np.uint64)
        return ((i + 0x7FFF + ((i >> 16) & 1)) >> 16).astype(np.uint16)
    assert np.sum(div != mul) > 0                      # the float32 values do differ ...
    np.testing.assert_array_equal(bf16_rn(div), bf16_rn(mul))      # ... the operands never do


# ---------------------------------------------------------------- mode "bf16x3": three bf16 planes, six products ----
def _bf16(a):
    u = np.ascontiguousarray(a, np.float32).view(np.uint32).astype(np.uint64)
    return (((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16).astype(np.uint32).view(np.float32)


def _split3(a):
    a = np.ascontiguousarray(a, np.float32)
    h = _bf16(a)
    m = _bf16(a - h)
    return h, m, _bf16(a - h - m)


def test_three_plane_split_and_six_products_are_fp32_exact():
    """The arithmetic contract of the FP32-exact tensor-core mode (nm_tc_common.cuh: split3, split_plane_a / split_plane_b),
    restated in NumPy: the planes carry 24 mantissa bits, the pass table in the header decodes to the six plane pairs whose
    weight is above 2^-24 ordered smallest first, their sum with FP32 accumulation matches the float64 product to ~1e-6 of
    the output scale, and the three dropped pairs are below 2^-23 of it -- while the single bf16 product of mode "bf16" is
    ~1e-3 off, which is why that mode states a looser tolerance."""
    import pathlib
    import re
    src = (pathlib.Path(__file__).resolve().parents[1] / "neuromah_b200" / "csrc" / "nm_tc_common.cuh").read_text()
    ta = int(re.search(r"split_plane_a\(int pr\) \{ return \((0x[0-9a-fA-F]+) >>", src).group(1), 16)
    tb = int(re.search(r"split_plane_b\(int pr\) \{ return \((0x[0-9a-fA-F]+) >>", src).group(1), 16)
    pairs = [((ta >> (4 * pr)) & 0xF, (tb >> (4 * pr)) & 0xF) for pr in range(6)]
    assert sorted(pairs) == sorted([(0, 0), (0, 1), (1, 0), (1, 1), (0, 2), (2, 0)])          # every pair with a + b <= 2
    assert [a + b for a, b in pairs] == sorted((a + b for a, b in pairs), reverse=True)       # smallest products first
    assert sorted(a for a, _ in pairs[:3]) == [0, 1, 2]      # passes 0-2 see each plane of the first operand once (wgrad's bias MMA)
    rng = np.random.default_rng(0)
    x = rng.standard_normal((64, 784)).astype(np.float32)
    w = (rng.standard_normal((784, 128)) * 0.05).astype(np.float32)
    xs, ws = _split3(x), _split3(w)
    for v, parts in ((x, xs), (w, ws)):
        assert np.abs((parts[0].astype(np.float64) + parts[1] + parts[2]) - v).max() <= 2.0 ** -24 * np.abs(v).max()
    ref = x.astype(np.float64) @ w.astype(np.float64)
    scale = np.abs(ref).max()
    acc = np.zeros(ref.shape, np.float32)
    for a, b in pairs:
        acc = acc + (xs[a] @ ws[b]).astype(np.float32)                                         # FP32 accumulator
    assert np.abs(acc - ref).max() <= 2e-6 * scale
    dropped = sum(np.abs(xs[a].astype(np.float64) @ ws[b]) for a, b in ((1, 2), (2, 1), (2, 2)))
    assert dropped.max() <= 2.0 ** -23 * scale
    one = (xs[0].astype(np.float64) @ ws[0])
    assert 1e-4 * scale < np.abs(one - ref).max() < 1e-2 * scale
