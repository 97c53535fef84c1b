"""CPU interpreter of the packed arena (test infrastructure).

Walks (header, table, arena) exactly like csrc/flow.cu / csrc/embed.cu do and evaluates the same
sequence of fused ops with torch on the CPU.  It exists so the pack-time algebra in
flowcompare_b200/packing.py (LayerNorm fold, attention out-projection fold, ActNorm+LinearLU fold, row
interleaving, BatchNorm fold, per-cloud bias) can be checked against the oracle WITHOUT a GPU.  It is
not a fallback: nothing in flowcompare_b200/ imports it.
"""
import math
import struct

import torch
import torch.nn.functional as F

from flowcompare_b200.packing import gemm_kpad, gemm_ldw


class Cursor:
    def __init__(self, table, arena):
        self.table, self.arena, self.pos = [int(v) for v in table], arena, 0

    def next(self):
        v = self.table[self.pos]
        self.pos += 1
        return v

    def vec(self, n):
        off = self.next()
        return self.arena[off:off + n]

    def linear(self, K1, K2, N, has_bias=True):
        ldw = gemm_ldw(N)
        kp1 = gemm_kpad(K1)
        kp = kp1 + (gemm_kpad(K2) if K2 else 0)
        off = self.next()
        Wt = self.arena[off:off + kp * ldw].view(kp, ldw)
        boff = self.next()
        b = self.arena[boff:boff + N] if has_bias else None
        hoff, loff = self.next(), self.next()
        tc = None
        if hoff >= 0:
            # tcgen05 copies: hi + lo must reproduce the fp32 weight (to ~2^-22), in [rows][ldk] layout
            from flowcompare_b200.packing import tc_bn, tc_kpad, tc_n_tiles
            t1 = tc_kpad(K1)
            ldk = t1 + (tc_kpad(K2) if K2 else 0)
            rows = tc_n_tiles(N) * tc_bn(N)
            hi = self.arena[hoff:hoff + rows * ldk].view(rows, ldk)
            lo = self.arena[loff:loff + rows * ldk].view(rows, ldk)
            tc = dict(hi=hi, lo=lo, t1=t1, ldk=ldk)
        return dict(Wt=Wt, b=b, K1=K1, K2=K2, N=N, kp1=kp1, tc=tc)

    def mlp(self, K1, K2, hid, n_hidden, N_out):
        return dict(inp=self.linear(K1, K2, hid), hidden=[self.linear(hid, 0, hid) for _ in range(n_hidden)],
                    out=self.linear(hid, 0, N_out))


USE_TC = False   # True: evaluate every linear that has tcgen05 copies the way csrc/gemm_tc.cu does (3xTF32)


def _lin_3xtf32(l, A1, A2):
    from flowcompare_b200.packing import tf32_round
    tc = l["tc"]
    A = torch.zeros(A1.shape[0], tc["ldk"])
    A[:, :l["K1"]] = A1[:, :l["K1"]]
    if l["K2"]:
        A[:, tc["t1"]:tc["t1"] + l["K2"]] = A2[:, :l["K2"]]
    a_hi = tf32_round(A)
    a_lo = tf32_round(A - a_hi)
    whi, wlo = tc["hi"].double(), tc["lo"].double()
    y = a_lo.double() @ whi.t() + a_hi.double() @ wlo.t() + a_hi.double() @ whi.t()
    return y[:, :l["N"]].float()


def lin(l, A1, A2=None, bias=None):
    if USE_TC and l["tc"] is not None:
        y = _lin_3xtf32(l, A1, A2)
        b = bias if bias is not None else l["b"]
        return y + b if b is not None else y
    y = A1[:, :l["K1"]] @ l["Wt"][:l["K1"], :l["N"]]
    if l["K2"]:
        y = y + A2[:, :l["K2"]] @ l["Wt"][l["kp1"]:l["kp1"] + l["K2"], :l["N"]]
    b = bias if bias is not None else l["b"]
    return y + b if b is not None else y


def mlp_hidden(m, A1, A2=None, bias=None, act=F.gelu):
    h = act(lin(m["inp"], A1, A2, bias))
    res = None
    for i, l in enumerate(m["hidden"]):
        if i % 2 == 0:
            res, h = h, act(lin(l, h))
        else:
            h = act(res + lin(l, h))
    return h


def read_attn(c, attn_in, E, inner):
    return dict(csum=c.vec(inner), qbias=c.vec(inner), q=c.linear(attn_in, 0, inner, False),
                kv=c.linear(E, 0, 2 * inner, False))


def attention_block(pre, at, lat_cols, context, B, N, inner, act=F.gelu):
    h = mlp_hidden(pre, lat_cols, act=act)
    h4 = lin(pre["out"], h)
    mu = h4.mean(-1, keepdim=True)
    rstd = torch.rsqrt(((h4 - mu) ** 2).mean(-1, keepdim=True) + 1e-5)
    acc = lin(at["q"], h4)
    q = rstd * (acc - mu * at["csum"]) + at["qbias"]
    kv = lin(at["kv"], context.reshape(-1, context.shape[-1]))
    q = q.view(B, N, inner)
    kv = kv.view(B, -1, 2 * inner)
    w = torch.softmax(q @ kv[..., :inner].transpose(1, 2) * (inner ** -0.5), dim=-1)
    return (w @ kv[..., inner:]).reshape(B * N, inner)


def flow_log_prob(packed, x, context, extra, eps, eps_cif=None):
    """Mirrors csrc/flow.cu: flow_forward (header words 19.. select the coupling kind, the conditioner non-linearity, the
    identity augmenter and the CIF block; see packing.pack_flow)."""
    from oracle import port
    header, table, arena = packed
    hd = [int(v) for v in header]
    (_, _, L, D, d_in, half, ex, is_global, E, inner, attn_in, hid, n_hid, pre_hid, n_pre, aug_hid, n_aug,
     augpre_hid, n_augpre) = hd[:19]
    cpl_kind, nb, act_code, has_aug, D2, cif_hid, n_cif, affcif_hid, n_affcif = hd[19:28] if len(hd) >= 29 else (0, 0, 1, 1, 0, 0, 0, 0, 0)
    clamp = struct.unpack("<f", struct.pack("<i", hd[28]))[0] if len(hd) >= 29 else 0.0
    act = {1: F.gelu, 3: F.relu}[act_code]
    n2 = D - half
    S = D2 - D if D2 else 0
    cpl_out = {0: 2 * n2, 1: (3 * nb + 1) * half, 2: n2 * n2 + n2}[cpl_kind]
    c = Cursor(table, arena)
    ldj_const = struct.unpack("<d", struct.pack("<q", c.next()))[0]
    B, N = x.shape[0], x.shape[1]
    M = B * N
    k2 = 0 if is_global else inner
    if has_aug:
        if not is_global:
            augpre = c.mlp(d_in, 0, augpre_hid, n_augpre, attn_in)
            augattn = read_attn(c, attn_in, E, inner)
        aug = c.mlp(d_in, k2, aug_hid, n_aug, 2 * (D - d_in))
    has_cb = bool(ex) or bool(is_global)
    cb = None
    if has_cb:
        cbl = c.linear(ex + (E if is_global else 0), 0, (L + 1) * hid)
        parts = []
        if ex:
            parts.append(extra.reshape(B, 1))
        if is_global:
            parts.append(context.reshape(B, E))
        cb = lin(cbl, torch.cat(parts, dim=1))  # [B, (L+1)*hid]

    def cloud_bias(slot):
        if cb is None:
            return None
        return cb[:, slot * hid:(slot + 1) * hid].repeat_interleave(N, dim=0)

    lat = torch.zeros(M, max(D, D2))
    lat[:, :d_in] = x.reshape(M, d_in)
    o = None
    logp = torch.zeros(M)
    if has_aug:
        if not is_global:
            o = attention_block(augpre, augattn, lat, context, B, N, inner, act)
        h = mlp_hidden(aug, lat, o, cloud_bias(0), act)
        st = lin(aug["out"], h)
        mean, log_std = st[:, 0::2], st[:, 1::2]
        e = eps.reshape(M, D - d_in)
        lat[:, d_in:D] = mean + torch.exp(log_std) * e
        logp = (0.5 * e * e + log_std + 0.5 * math.log(2 * math.pi)).sum(-1)
    for l in range(L):
        if D2:
            cifnet = c.mlp(D, 0, cif_hid, n_cif, 2 * S)
            affcif = c.mlp(S, 0, affcif_hid, n_affcif, 2 * D)
            sc, bi = c.vec(D2), c.vec(D2)

            def cond_normal():
                pr = lin(cifnet["out"], mlp_hidden(cifnet, lat))
                sigma = torch.exp(pr[:, S:])
                return pr[:, :S], (sigma.clamp_max(clamp) if clamp > 0 else sigma)
            mean, sigma = cond_normal()
            e = eps_cif[l].reshape(M, S)
            lat[:, D:D2] = mean + sigma * e
            logp = logp + (0.5 * e * e + torch.log(sigma) + 0.5 * math.log(2 * math.pi)).sum(-1)
            st = lin(affcif["out"], mlp_hidden(affcif, lat[:, D:D2]))
            s = (2 * torch.sigmoid(st[:, 0::2]) - 1) + 1
            lat[:, :D] = lat[:, :D] * s + st[:, 1::2]
            logp = logp + torch.log(s).sum(-1)
            lat = lat * sc + bi
            mean, sigma = cond_normal()
            e = (lat[:, D:D2] - mean) / sigma
            logp = logp - (0.5 * e * e + torch.log(sigma) + 0.5 * math.log(2 * math.pi)).sum(-1)
        if not is_global:
            pre = c.mlp(half, 0, pre_hid, n_pre, attn_in)
            at = read_attn(c, attn_in, E, inner)
            o = attention_block(pre, at, lat, context, B, N, inner, F.gelu if D2 else act)
        sq = c.vec(4) if cpl_kind == 2 else None
        cpl = c.mlp(half, k2, hid, n_hid, cpl_out)
        h = mlp_hidden(cpl, lat, o, cloud_bias(l + 1), act)
        st = lin(cpl["out"], h)
        if cpl_kind == 0:
            s_raw, t = st[:, 0::2], st[:, 1::2]
            s = (2 * torch.sigmoid(s_raw) - 1) + 1
            y2, ldj = lat[:, half:D] * s + t, torch.log(s).sum(-1)
        elif cpl_kind == 1:
            pr = st.reshape(M, half, 3 * nb + 1)
            y2, lad = port.rq_spline(lat[:, half:D], pr[..., :nb], pr[..., nb:2 * nb], pr[..., 2 * nb:])
            ldj = lad.sum(-1)
        else:
            W = (sq[2] * torch.tanh(sq[0] * st[:, :n2 * n2] + sq[1]) + sq[3] + 1e-8).reshape(M, n2, n2)
            y2 = torch.matmul(torch.matrix_exp(W), lat[:, half:D].unsqueeze(-1)).squeeze(-1) + st[:, n2 * n2:]
            ldj = W.diagonal(dim1=-2, dim2=-1).sum(-1)
        lat = torch.cat((lat[:, :half], y2, lat[:, D:]), dim=1)
        logp = logp + ldj
        if l != L - 1:
            lu = c.linear(D, 0, D)
            lat = torch.cat((lin(lu, lat[:, :D]) + c.vec(D) * lat[:, :D], lat[:, D:]), dim=1)
    assert c.pos == len(c.table), (c.pos, len(c.table))
    logp = logp + ldj_const + (-0.5 * math.log(2 * math.pi) - 0.5 * lat[:, :D] ** 2).sum(-1)
    return logp.view(B, N)


def dgcnn_embed(packed, pts, knn_fn):
    """knn_fn(x [B,N,C], k) -> idx [B,N,k] (injected so the test can use the canonical-arithmetic oracle)."""
    header, table, arena = packed
    _, _, kind, d_in, k, E, out_hid, n_out = [int(v) for v in header]
    c = Cursor(table, arena)
    B, N = pts.shape[0], pts.shape[1]
    cin, cout = [d_in, 64, 64, 128], [64, 64, 128, 256]
    feats = []
    x = pts
    idxs = []
    for i in range(4):
        pq = c.linear(cin[i], 0, 2 * cout[i])
        idx = knn_fn(x, k)
        idxs.append(idx)
        PQ = lin(pq, x.reshape(B * N, -1)).view(B, N, 2 * cout[i])
        P, Q = PQ[..., :cout[i]], PQ[..., cout[i]:]
        gathered = torch.stack([P[b][idx[b]] for b in range(B)])  # [B,N,k,Cout]
        x = F.leaky_relu(gathered.max(dim=2)[0] + Q, 0.2)
        feats.append(x)
    conv5 = c.linear(512, 0, 512)
    y = F.leaky_relu(lin(conv5, torch.cat(feats, dim=-1).reshape(B * N, 512)), 0.2)
    out_mlp = c.mlp(1024 if kind == 1 else 512, 0, out_hid, n_out, E)
    assert c.pos == len(c.table)
    if kind == 0:
        return lin(out_mlp["out"], mlp_hidden(out_mlp, y)).view(B, N, E), idxs
    y = y.view(B, N, 512)
    pooled = torch.cat((y.max(dim=1)[0], y.mean(dim=1)), dim=1)
    return lin(out_mlp["out"], mlp_hidden(out_mlp, pooled)), idxs
