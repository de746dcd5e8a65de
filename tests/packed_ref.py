"""CPU emulation of the CUDA data flow from the PACKED weights (test helper, not product code).

Walks the same launch plan as s3od_b200/csrc/engine.cu with torch CPU ops: im2col patch GEMM, fused QKV with
half-table RoPE, tap-major implicit-GEMM convolutions, sub-pixel transposed convolutions, out_conv commuted in
front of the bilinear up-sampling, merged mask heads.  With `emulate_bf16=True` activations are rounded to bf16 at
the points where the kernels store bf16, which predicts the numerical gap to the fp32 oracle without a GPU.
"""
import torch
import torch.nn.functional as F


def _r(t, on):
    return t.to(torch.bfloat16).float() if on else t


def _conv3x3_tapmajor(x_nhwc, w, bias=None):
    """x (B,H,W,C) ; w [Cout, 9*Cin] tap-major."""
    B, H, W, C = x_nhwc.shape
    co = w.shape[0]
    wk = w.float().reshape(co, 3, 3, C).permute(0, 3, 1, 2)
    y = F.conv2d(x_nhwc.permute(0, 3, 1, 2), wk, bias, padding=1)
    return y.permute(0, 2, 3, 1)


def forward_packed(pw, x, arch, emulate_bf16=False, stages=None):
    """pw: dict from weights.pack_weights; x: (B,3,S,S) fp32.  Returns pred_masks (B,K,S,S), pred_iou (B,K)."""
    bf = emulate_bf16
    B, _, S, _ = x.shape
    g = S // 16
    P, D, H, K = g * g, arch.hidden, arch.heads, arch.num_outputs
    ntok = P + 5
    # patches: row (b, py, px), col c*256 + ky*16 + kx
    patches = x.reshape(B, 3, g, 16, g, 16).permute(0, 2, 4, 1, 3, 5).reshape(B * P, 768)
    patches = _r(patches, True)                                         # the kernel input is bf16 regardless
    tok = patches @ pw["patch.w"].float().t() + pw["patch.b"]
    xs = torch.cat([pw["prefix"][None].expand(B, -1, -1), tok.reshape(B, P, D)], 1)      # fp32 residual
    cos, sin = pw["rope.cos"], pw["rope.sin"]                           # (P, 32)
    taps = []
    for l in range(arch.layers_needed):
        o = f"enc.{l}."
        xn = _r(F.layer_norm(xs, (D,), pw[o + "ln1.w"], pw[o + "ln1.b"], arch.ln_eps), bf)
        qkv = xn @ pw[o + "qkv.w"].float().t() + pw[o + "qkv.b"]
        q, k, v = qkv.split(D, dim=-1)
        q = q.reshape(B, ntok, H, 64).transpose(1, 2).clone()
        k = k.reshape(B, ntok, H, 64).transpose(1, 2).clone()
        v = v.reshape(B, ntok, H, 64).transpose(1, 2)
        for t in (q, k):
            lo, hi = t[:, :, 5:, :32].clone(), t[:, :, 5:, 32:].clone()
            t[:, :, 5:, :32] = lo * cos - hi * sin
            t[:, :, 5:, 32:] = hi * cos + lo * sin
        q = _r(q * (0.125 * 1.4426950408889634), bf)
        k, v = _r(k, bf), _r(v, bf)
        s = q @ k.transpose(2, 3)                                       # log2 domain
        p = torch.exp2(s - s.amax(-1, keepdim=True))
        ctx = (_r(p, bf) @ v) / p.sum(-1, keepdim=True)
        ctx = _r(ctx.transpose(1, 2).reshape(B, ntok, D), bf)
        xs = xs + (ctx @ pw[o + "o.w"].float().t() + pw[o + "o.b"]) * pw[o + "ls1"]
        xn = _r(F.layer_norm(xs, (D,), pw[o + "ln2.w"], pw[o + "ln2.b"], arch.ln_eps), bf)
        hm = _r(F.gelu(xn @ pw[o + "up.w"].float().t() + pw[o + "up.b"]), bf)
        xs = xs + (hm @ pw[o + "down.w"].float().t() + pw[o + "down.b"]) * pw[o + "ls2"]
        if stages is not None and l == 0:
            stages["layer0"] = xs.clone()
        if (l + 1) in arch.taps:
            taps.append(_r(xs[:, 5:].reshape(B, g, g, D), bf))
    if stages is not None:
        for i, t in enumerate(taps):
            stages[f"tap{i}"] = t.reshape(B, P, D)
    oc = arch.out_channels
    f = [_r(taps[j] @ pw[f"head.proj{j}.w"].float().t() + pw[f"head.proj{j}.b"], bf) for j in range(4)]

    def convt_eq_stride(xin, w, b, k, co):
        y = xin @ w.float().t()                                         # (B,g,g,k*k*co)
        y = y.reshape(B, g, g, k, k, co).permute(0, 1, 3, 2, 4, 5).reshape(B, g * k, g * k, co)
        return y + b

    r0 = _r(convt_eq_stride(f[0], pw["head.rs0.w"], pw["head.rs0.b"], 4, oc[0]), bf)
    r1 = _r(convt_eq_stride(f[1], pw["head.rs1.w"], pw["head.rs1.b"], 2, oc[1]), bf)
    w3 = pw["head.rs3.w"].float().reshape(oc[3], 3, 3, oc[3]).permute(0, 3, 1, 2)
    r3 = _r(F.conv2d(f[3].permute(0, 3, 1, 2), w3, pw["head.rs3.b"], stride=2, padding=1).permute(0, 2, 3, 1), bf)
    rn_in = [r0, r1, f[2], r3]
    lay = [_r(_conv3x3_tapmajor(rn_in[j], pw[f"head.rn{j + 1}.w"]), bf) for j in range(4)]

    def rcu(prefix, xin, extra=None):
        a = _r(F.relu(_conv3x3_tapmajor(F.relu(xin), pw[prefix + "c1.w"], pw[prefix + "c1.b"])), bf)
        y = _conv3x3_tapmajor(a, pw[prefix + "c2.w"], pw[prefix + "c2.b"]) + xin
        if extra is not None:
            y = y + extra
        return _r(y, bf)

    def up2(t):
        return F.interpolate(t.permute(0, 3, 1, 2), scale_factor=2, mode="bilinear", align_corners=False).permute(0, 2, 3, 1)

    path = None
    pooled = None
    for kk in (4, 3, 2, 1):
        pre = f"head.ref{kk}."
        s_in = lay[kk - 1] if kk == 4 else rcu(pre + "rcu1.", lay[kk - 1], path)
        u = rcu(pre + "rcu2.", s_in)
        qk = _r(u @ pw[pre + "out.w"].float().t() + pw[pre + "out.b"], bf)
        up = up2(qk)
        if kk == 1:
            pooled = up.mean(dim=(1, 2))
        path = _r(up, bf)
        if stages is not None:
            stages[f"p{kk}"] = path
    hid = F.relu(pooled @ pw["head.cls.w1"].t() + pw["head.cls.b1"])
    iou = hid @ pw["head.cls.w2"].t() + pw["head.cls.b2"]
    m1 = _r(_conv3x3_tapmajor(path, pw["head.mh.c1.w"], pw["head.mh.c1.b"]), bf)
    Hh = m1.shape[1]
    feat0 = torch.zeros(B, 2 * Hh, 2 * Hh, 64)
    wup = pw["head.mh.up.w"].float()
    off = {0: (0, -1), 1: (1, 0)}
    mp = F.pad(m1, (0, 0, 1, 1, 1, 1))                                  # zero pad H and W by 1
    for a in (0, 1):
        for b in (0, 1):
            wp = wup[(a * 2 + b) * 64:(a * 2 + b + 1) * 64]             # (64, 4*128)
            acc = 0
            for t in range(4):
                dh, dw = off[a][t // 2], off[b][t % 2]
                sl = mp[:, 1 + dh:1 + dh + Hh, 1 + dw:1 + dw + Hh, :]
                acc = acc + sl @ wp[:, t * 128:(t + 1) * 128].t()
            feat0[:, a::2, b::2, :] = acc
    feat0 = _r(F.relu(feat0 + pw["head.mh.up.b"]), bf)
    feat = _r(F.relu(_conv3x3_tapmajor(feat0, pw["head.mh.c2.w"], pw["head.mh.c2.b"])), bf)
    hh = F.relu(_conv3x3_tapmajor(feat, pw["head.mh.heads.w"], pw["head.mh.heads.b"]))   # (B,S,S,32K)
    masks = (hh.reshape(B, S, S, K, 32) * pw["head.mh.heads.w2"]).sum(-1) + pw["head.mh.heads.b2"]
    if stages is not None:
        stages["feat"] = feat
    return masks.permute(0, 3, 1, 2).contiguous(), iou
