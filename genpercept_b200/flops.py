"""Algorithmic work of the hot path: FLOPs = 2*MACs of conv / linear /
QK^T / PV only, 2-token text context.  Pure arithmetic on the SD-2.1 topology — shared by both
bench arms, independent of any kernel."""


def _conv(cin, cout, k, hw):
    return 2.0 * cin * cout * k * k * hw


def _res(cin, cout, hw):
    f = _conv(cin, cout, 3, hw) + _conv(cout, cout, 3, hw)
    if cin != cout:
        f += _conv(cin, cout, 1, hw)
    return f


def _attn_sdpa(tokens, c):
    return 4.0 * tokens * tokens * c          # QK^T + PV


def _vae_mid(hw):
    return 2 * _res(512, 512, hw) + 4 * _conv(512, 512, 1, hw) + _attn_sdpa(hw, 512)


def vae_encoder_flops(H, W):
    hw = H * W
    f = _conv(3, 128, 3, hw)
    ch = (128, 128, 256, 512, 512)
    for i in range(4):
        f += _res(ch[i], ch[i + 1], hw) + _res(ch[i + 1], ch[i + 1], hw)
        if i < 3:
            hw //= 4
            f += _conv(ch[i + 1], ch[i + 1], 3, hw)
    f += _vae_mid(hw) + _conv(512, 8, 3, hw) + _conv(8, 8, 1, hw)
    return f


def vae_decoder_flops(H, W):
    hw = (H // 8) * (W // 8)
    f = _conv(4, 4, 1, hw) + _conv(4, 512, 3, hw) + _vae_mid(hw)
    prev, outc = (512, 512, 512, 256), (512, 512, 256, 128)
    for i in range(4):
        f += _res(prev[i], outc[i], hw) + 2 * _res(outc[i], outc[i], hw)
        if i < 3:
            hw *= 4
            f += _conv(outc[i], outc[i], 3, hw)
    return f + _conv(128, 3, 3, hw)


def _transformer(c, hw):
    lin = 2.0 * hw * (c * c * 2          # proj_in, proj_out
                      + 4 * c * c        # attn1 q,k,v,out
                      + 2 * c * c        # attn2 q,out (k,v on 2 tokens ~ 0)
                      + c * 8 * c + 4 * c * c)   # GEGLU proj, ff out
    return lin + _attn_sdpa(hw, c) + 4.0 * hw * 2 * c   # + 2-token cross attention


def unet_flops(H, W, with_out=True):
    hw = (H // 8) * (W // 8)
    f = _conv(4, 320, 3, hw)
    cin = 320
    outs = (320, 640, 1280, 1280)
    for i, c in enumerate(outs):
        for j in range(2):
            f += _res(cin if j == 0 else c, c, hw)
            if i < 3:
                f += _transformer(c, hw)
        if i < 3:
            hw //= 4
            f += _conv(c, c, 3, hw)
        cin = c
    f += 2 * _res(1280, 1280, hw) + _transformer(1280, hw)
    up = [(1280, 1280, (1280, 1280, 1280), False), (1280, 1280, (1280, 1280, 640), True),
          (1280, 640, (640, 640, 320), True), (640, 320, (320, 320, 320), True)]
    for i, (cprev, c, skips, attn) in enumerate(up):
        for j in range(3):
            f += _res((cprev if j == 0 else c) + skips[j], c, hw)
            if attn:
                f += _transformer(c, hw)
        if i < 3:
            hw *= 4
            f += _conv(c, c, 3, hw)
    if with_out:
        f += _conv(320, 4, 3, hw)
    return f


def dpt_head_flops(H, W):
    h = (H // 8) * (W // 8)
    f = _conv(320, 320, 3, 4 * h)
    sizes = (4 * h, h, h // 4, h // 16)
    for c, hw in zip((320, 640, 1280, 1280), sizes):
        f += _conv(c, 256, 3, hw)
    for li, hw in enumerate(sizes[::-1]):
        f += (2 if li == 0 else 4) * _conv(256, 256, 3, hw) + _conv(256, 256, 1, 4 * hw)
    hw = 16 * h
    f += _conv(256, 256, 3, hw) + _conv(256, 128, 3, hw) + _conv(128, 32, 3, 4 * hw) + _conv(32, 1, 1, 4 * hw)
    return f


def single_infer_flops(H, W, readout="vae"):
    if readout == "dpt":
        return vae_encoder_flops(H, W) + unet_flops(H, W, with_out=False) + dpt_head_flops(H, W)
    return vae_encoder_flops(H, W) + unet_flops(H, W) + vae_decoder_flops(H, W)
