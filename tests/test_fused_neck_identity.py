"""CPU check (float64, no GPU) of the algebra behind the BF16 graph's fused neck (DESIGN section 4,
detector.cu prep_fused_fpn_level / prep_fused_bin_p3 / upload_pair_conv3): the weight transformations restated
here in numpy must reproduce the reference's literal graph (model.rs:113-143) exactly."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
F = torch.nn.functional


def up2(t):
    return t.repeat_interleave(2, dim=2).repeat_interleave(2, dim=3)


def class_kernel(w, a, b):
    """3x3 kernel over the LOW-resolution map that equals conv3x3(w) of the 2x nearest-upsampled map at output
    parity (a, b): full-res tap dy lands on low-res row offset ((a + dy) >> 1) (floor), same for columns."""
    k = torch.zeros_like(w)
    for dy in (-1, 0, 1):
        for dx in (-1, 0, 1):
            rr, ss = ((a + dy) >> 1) + 1, ((b + dx) >> 1) + 1
            k[:, :, rr, ss] += w[:, :, dy + 1, dx + 1]
    return k


def shuffle(classes, h, w):
    """classes[a][b]: [B, C, h, w] -> [B, C, 2h, 2w] with class (a, b) of pixel (Y, X) at (2Y + a, 2X + b)."""
    out = torch.zeros(classes[0][0].shape[0], classes[0][0].shape[1], 2 * h, 2 * w, dtype=classes[0][0].dtype)
    for a in (0, 1):
        for b in (0, 1):
            out[:, :, a::2, b::2] = classes[a][b]
    return out


@pytest.mark.parametrize("c_l,c_u", [(64, 128), (128, 256)])
def test_fpn_level_identity(c_l, c_u):
    g = torch.Generator().manual_seed(c_l)
    rnd = lambda *s: torch.randn(*s, generator=g, dtype=torch.float64)
    x_l, x_u = rnd(2, c_l, 8, 12), rnd(2, c_u, 4, 6)
    w_in_l, w_in_u, w_out = rnd(256, c_l, 1, 1), rnd(256, c_u, 1, 1), rnd(64, 256, 3, 3)
    ref = F.conv2d(F.conv2d(x_l, w_in_l) + up2(F.conv2d(x_u, w_in_u)), w_out, padding=1)
    # (1) composed 3x3 on x_l
    wc = torch.einsum("omrs,mi->oirs", w_out, w_in_l[:, :, 0, 0])
    main = F.conv2d(x_l, wc, padding=1)
    # (2) four parity-class kernels, composed with the upper lateral, at low resolution, pixel-shuffled
    classes = [[None, None], [None, None]]
    taps = {}
    for a in (0, 1):
        for b in (0, 1):
            k = torch.einsum("omrs,mi->oirs", class_kernel(w_out, a, b), w_in_u[:, :, 0, 0])
            taps[(a, b)] = {(r, s) for r in range(3) for s in range(3) if k[:, :, r, s].abs().max() > 0}
            classes[a][b] = F.conv2d(x_u, k, padding=1)
    # every class uses 4 of the 9 taps: rows {0,1} / {1,2} for a = 0 / 1, same for columns
    for (a, b), t in taps.items():
        assert t == {(r, s) for r in (a, a + 1) for s in (b, b + 1)}
    got = main + shuffle(classes, 4, 6)
    assert torch.allclose(got, ref, rtol=1e-10, atol=1e-9)
    # (3) pair form: class (a,1) of column X and class (a,0) of column X + 1 read the same patch (columns X, X + 1):
    #     one 128-row kernel with taps rows(a) x {1, 2}, evaluated on output columns -1 .. W - 1
    for a in (0, 1):
        k1 = torch.einsum("omrs,mi->oirs", class_kernel(w_out, a, 1), w_in_u[:, :, 0, 0])
        k0 = torch.einsum("omrs,mi->oirs", class_kernel(w_out, a, 0), w_in_u[:, :, 0, 0])
        kp = torch.zeros(128, c_u, 3, 3, dtype=torch.float64)
        kp[:64, :, :, 1:] = k1[:, :, :, 1:]
        kp[64:, :, :, 1:] = k0[:, :, :, :2]
        # output index i = low-res column + 1: pad one extra input column on the left, evaluate W + 1 columns
        xp = F.pad(x_u, (2, 1, 1, 1))  # left 2 (= conv pad 1 + shift 1), right 1, top / bottom 1
        pair = F.conv2d(xp, kp)        # [B, 128, 4, 6 + 1]
        assert torch.allclose(pair[:, :64, :, 1:], classes[a][1], rtol=1e-10, atol=1e-9)   # half 0: column i - 1
        assert torch.allclose(pair[:, 64:, :, :-1], classes[a][0], rtol=1e-10, atol=1e-9)  # half 1: column i


def test_bin_conv1_nested_concat_identity():
    g = torch.Generator().manual_seed(7)
    rnd = lambda *s: torch.randn(*s, generator=g, dtype=torch.float64)
    p5, p4, p3, p2 = rnd(1, 64, 2, 3), rnd(1, 64, 4, 6), rnd(1, 64, 8, 12), rnd(1, 64, 16, 24)
    w = rnd(64, 256, 3, 3)
    scale, shift = rnd(64).abs() + 0.1, rnd(64)
    up = lambda t, r: t.repeat_interleave(r, dim=2).repeat_interleave(r, dim=3)
    fuse = torch.cat([up(p5, 8), up(p4, 4), up(p3, 2), p2], 1)  # model.rs:140
    ref = F.conv2d(fuse, w, padding=1) * scale[None, :, None, None] + shift[None, :, None, None]
    cat3 = torch.cat([up(p5, 4), up(p4, 2), p3], 1)  # nearest upsampling composes: up2(cat3) == fuse[:, :192]
    assert torch.equal(up2(cat3), fuse[:, :192])
    classes = [[F.conv2d(cat3, class_kernel(w[:, :192], a, b) * scale[:, None, None, None], padding=1) for b in (0, 1)] for a in (0, 1)]
    main = F.conv2d(p2, w[:, 192:], padding=1) * scale[None, :, None, None] + shift[None, :, None, None]
    got = main + shuffle(classes, 8, 12)  # the class sum joins after the main convolution's scale / shift
    assert torch.allclose(got, ref, rtol=1e-10, atol=1e-9)
