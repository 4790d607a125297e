"""CPU oracle (TEST INFRASTRUCTURE ONLY) of the companion operators of SURVEY.md §8f rank 4: a functional torch-CPU
restatement over a plain state_dict.  Pinned against outputs of the unmodified reference class run in the build container
(tests/golden/make_golden_companions.py -> tests/golden/freprocess_c8.npz, tests/test_oracle_companions.py).
Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this."""
import torch
import torch.nn.functional as F


def freprocess_forward(sd, msf, panf):
    """SFIIN.Freprocess.forward, models/SFIIN.py:221-236.  sd keys: pre1, pre2, amp_fuse.{0,2}, pha_fuse.{0,2}, post
    (.weight [Cout,Cin,1,1], .bias)."""
    def conv(name, x):
        return F.conv2d(x, sd[name + ".weight"], sd[name + ".bias"])

    def fuse(name, x):                                                          # :215-218
        return conv(name + ".2", F.leaky_relu(conv(name + ".0", x), 0.1))

    H, W = msf.shape[-2:]
    msF = torch.fft.rfft2(conv("pre1", msf) + 1e-8, norm="backward")            # :223
    panF = torch.fft.rfft2(conv("pre2", panf) + 1e-8, norm="backward")          # :224
    amp = fuse("amp_fuse", torch.cat([torch.abs(msF), torch.abs(panF)], 1))     # :225-229
    pha = fuse("pha_fuse", torch.cat([torch.angle(msF), torch.angle(panF)], 1))  # :226-230
    real = amp * torch.cos(pha) + 1e-8                                          # :232
    imag = amp * torch.sin(pha) + 1e-8                                          # :233
    out = torch.complex(real, imag) + 1e-8                                      # :234
    out = torch.abs(torch.fft.irfft2(out, s=(H, W), norm="backward"))           # :235
    return conv("post", out)                                                    # :236


def window_attention_forward(sd, x, y, heads, head_dim, window_size, shifted, relative_pos_embedding):
    """PanFormer WindowAttention.forward, models/common/modules.py:387-437, over a plain state_dict (to_qkv.weight or
    to_q.weight + to_kv.weight, pos_embedding, to_out.weight/.bias, upper_lower_mask/left_right_mask when shifted).
    x (and y for cross attention): [b, n_h, n_w, dim]."""
    ws, d = window_size, window_size // 2
    if shifted:                                                                     # :388-391
        x = torch.roll(x, shifts=(-d, -d), dims=(1, 2))
        if y is not None:
            y = torch.roll(y, shifts=(-d, -d), dims=(1, 2))
    b, n_h, n_w, _ = x.shape
    if y is None:                                                                   # :396-402
        q, k, v = F.linear(x, sd["to_qkv.weight"]).chunk(3, dim=-1)
    else:
        k, v = F.linear(x, sd["to_kv.weight"]).chunk(2, dim=-1)
        q = F.linear(y, sd["to_q.weight"])
    nw_h, nw_w = n_h // ws, n_w // ws

    def windows(t):                                                                 # :407-410 'b (nw_h w_h) (nw_w w_w) (h d) -> b h (nw_h nw_w) (w_h w_w) d'
        t = t.reshape(b, nw_h, ws, nw_w, ws, heads, head_dim)
        return t.permute(0, 5, 1, 3, 2, 4, 6).reshape(b, heads, nw_h * nw_w, ws * ws, head_dim)

    q, k, v = windows(q), windows(k), windows(v)
    dots = torch.einsum("bhwid,bhwjd->bhwij", q, k) * (head_dim ** -0.5)            # :416
    if relative_pos_embedding:                                                      # :336-339, :418-419
        idx = torch.arange(ws * ws)
        coords = torch.stack([idx // ws, idx % ws], dim=1)
        rel = coords[None, :, :] - coords[:, None, :] + ws - 1
        dots = dots + sd["pos_embedding"][rel[:, :, 0], rel[:, :, 1]]
    else:
        dots = dots + sd["pos_embedding"]
    if shifted:                                                                     # :423-425
        dots[:, :, -nw_w:] += sd["upper_lower_mask"]
        dots[:, :, nw_w - 1::nw_w] += sd["left_right_mask"]
    out = torch.einsum("bhwij,bhwjd->bhwid", dots.softmax(dim=-1), v)               # :427-429
    out = out.reshape(b, heads, nw_h, nw_w, ws, ws, head_dim).permute(0, 2, 4, 3, 5, 1, 6).reshape(b, n_h, n_w, heads * head_dim)
    out = F.linear(out, sd["to_out.weight"], sd["to_out.bias"])                     # :433
    if shifted:                                                                     # :436-437
        out = torch.roll(out, shifts=(d, d), dims=(1, 2))
    return out
