"""Jet (R-op / R^2-op) oracle on the CPU, fp64 -- TEST INFRASTRUCTURE ONLY.

Restates ``rop.py:69-164`` (the reference's closed-form forward / backward
recurrences for value, R{.} and R^2{.}) generalised from sigmoid-MLP + MSE to the
layer set of the BASELINE configs, and executes it over the SAME tape
(``optwboundeigenval_b200.tracer``) the CUDA library interprets.  Its purpose is
to pin the *algorithm* the kernels implement -- tape lowering, concatenation
aliasing, overwrite/accumulate flags, BatchNorm and loss-head jets -- against the
autograd oracle on the CPU, so that a GPU mismatch can only be a kernel bug.
``tests/test_jet_oracle.py`` checks it against ``autograd_oracle`` (itself pinned
to the reference by the golden vectors).

Conventions as in the library: order k = k-th derivative in t of the quantity
evaluated at w + t v (rop.py: rx/ry = order 1, r2x/r2y = order 2).
"""
from __future__ import annotations

from typing import List

import torch
import torch.nn.functional as F

from optwboundeigenval_b200 import tracer as T

DT = torch.float64


# ---- jets as python lists [a0, a1, a2] of broadcastable tensors ---------------------------------
def jmul(a, b, K):
    r = [a[0] * b[0]]
    if K >= 1:
        r.append(a[1] * b[0] + a[0] * b[1])
    if K >= 2:
        r.append(a[2] * b[0] + 2 * a[1] * b[1] + a[0] * b[2])
    return r


def jadd(a, b, K):
    return [a[k] + b[k] for k in range(K + 1)]


def jsub(a, b, K):
    return [a[k] - b[k] for k in range(K + 1)]


def jexp(a, K):
    e = torch.exp(a[0])
    r = [e]
    if K >= 1:
        r.append(e * a[1])
    if K >= 2:
        r.append(e * (a[2] + a[1] * a[1]))
    return r


def jrecip(a, K):
    q = 1.0 / a[0]
    r = [q]
    if K >= 1:
        r.append(-a[1] * q * q)
    if K >= 2:
        r.append((2 * a[1] * a[1] * q - a[2]) * q * q)
    return r


def jrsqrt(s, K):
    r0 = s[0] ** -0.5
    r = [r0]
    if K >= 1:
        r.append(-0.5 * r0 ** 3 * s[1])
    if K >= 2:
        r.append(0.75 * r0 ** 5 * s[1] ** 2 - 0.5 * r0 ** 3 * s[2])
    return r


def jsigmoid(u, K):
    e = jexp([-x for x in u], K)
    e[0] = e[0] + 1.0
    return jrecip(e, K)


def jsoftmax(u, K):
    m = u[0].max(dim=1, keepdim=True).values
    e = jexp([u[0] - m] + list(u[1:]), K)
    S = [x.sum(dim=1, keepdim=True) for x in e]
    return jmul(e, jrecip(S, K), K)


class JetTapeOracle:
    def __init__(self, tape: T.Tape, params: torch.Tensor, x: torch.Tensor, target, coef=None, loss_scale=None):
        self.tape = tape
        self.B = x.shape[0]
        self.w = params.to(DT)
        self.P = tape.n_params
        self.target = target
        self.coef = coef
        self.loss_scale = loss_scale if loss_scale is not None else 1.0 / self.B
        nb = len(tape.buf_elems)
        self.fw = [[torch.zeros(self.B, tape.buf_elems[b], dtype=DT) for b in range(nb)] for _ in range(3)]
        self.bw = [[torch.zeros(self.B, tape.buf_elems[b], dtype=DT) for b in range(nb)] for _ in range(3)]
        self.out = [torch.zeros(self.P, dtype=DT) for _ in range(3)]
        self.v = torch.zeros(self.P, dtype=DT)
        self.stats = {}
        self.argmax = {}
        self.loss = None
        # ReLU decisions: op index -> bool tensor, recorded by the order-0 forward.  `mask_override`
        # (same keys) substitutes given decisions for the natural `pre-activation > 0`: used by the
        # parity tests to condition the oracle on the decisions another implementation took where the
        # pre-activation is within fp32 rounding of zero (tests/kinks.py).
        self.masks = {}
        self.preact = {}
        self.mask_override = {}
        # the same for max pooling: op index -> int64 index tensor (as F.max_pool2d returns) substituted for the natural
        # arg-max where two window elements are within fp32 rounding of each other
        self.argmax_override = {}
        t0 = tape.tensors[0]
        self.view(self.fw, 0, 0).copy_(x.to(DT).reshape(self.B, *t0.shape))

    def view(self, arena, k, t):
        vt = self.tape.tensors[t]
        c, h, w = vt.shape
        return arena[k][vt.buf][:, vt.offset:vt.offset + vt.numel].view(self.B, c, h, w)

    def wslice(self, vec, off, shape):
        n = 1
        for s in shape:
            n *= s
        return vec[off:off + n].view(*shape)

    def decide(self, oi, pre):
        """order-0 ReLU of op `oi`: records the pre-activation and the on/off decisions, returns the output"""
        self.preact[oi] = pre.clone()
        m = self.mask_override.get(oi)
        if m is None:
            m = pre > 0
        self.masks[oi] = m
        return pre * m

    # ---- passes --------------------------------------------------------------------------------
    def run(self, K, v=None):
        if v is not None:
            self.v = v.to(torch.float32).to(DT)        # the reference's cast rounds v to fp32
        self.forward(K)
        self.backward(K)
        return self.out[K]

    def forward(self, K):
        tp = self.tape
        for oi, op in enumerate(tp.ops):
            first = bool(op.flags & T.F_FIRST)
            relu = bool(op.flags & T.F_RELU)
            xin = [self.view(self.fw, k, op.inp) for k in range(3)]
            yout = self.view(self.fw, K, op.out)
            y0 = self.view(self.fw, 0, op.out)
            if op.kind == T.OP_CONV:
                cin = tp.tensors[op.inp].shape[0]
                cout = tp.tensors[op.out].shape[0]
                kh, kw, sh, sw, ph, pw = op.geom
                W = self.wslice(self.w, op.w_off, (cout, cin, kh, kw))
                V = self.wslice(self.v, op.w_off, (cout, cin, kh, kw))
                conv = lambda a, wt: F.conv2d(a, wt, None, (sh, sw), (ph, pw))   # noqa: E731
                if K == 0:
                    y = conv(xin[0], W)
                    if op.b_off >= 0:
                        y = y + self.w[op.b_off:op.b_off + cout].view(1, -1, 1, 1)
                elif K == 1:
                    y = conv(xin[0], V)
                    if not first:
                        y = y + conv(xin[1], W)
                    if op.b_off >= 0:
                        y = y + self.v[op.b_off:op.b_off + cout].view(1, -1, 1, 1)
                else:
                    y = torch.zeros_like(yout) if first else conv(xin[2], W) + 2 * conv(xin[1], V)
                if relu:
                    y = self.decide(oi, y) if K == 0 else y * self.masks[oi]
                yout.copy_(y)
            elif op.kind == T.OP_BN:
                self.bn_forward(oi, op, K, first, relu)
            elif op.kind == T.OP_RELU:
                yout.copy_(self.decide(oi, xin[0]) if K == 0 else xin[K] * self.masks[oi])
            elif op.kind == T.OP_MAXPOOL:
                kh, kw, sh, sw, ph, pw = op.geom
                if K == 0:
                    y, idx = F.max_pool2d(xin[0], (kh, kw), (sh, sw), (ph, pw), return_indices=True)
                    if oi in self.argmax_override:
                        idx = self.argmax_override[oi]
                        y = xin[0].flatten(2).gather(2, idx.flatten(2)).view_as(y)
                    self.argmax[oi] = idx
                    yout.copy_(y)
                else:
                    idx = self.argmax[oi]
                    flat = xin[K].flatten(2)
                    yout.copy_(flat.gather(2, idx.flatten(2)).view_as(yout))
            elif op.kind == T.OP_AVGPOOL:
                yout.copy_(F.avg_pool2d(xin[K], op.geom[0]))
            elif op.kind == T.OP_COPY:
                yout.copy_(xin[K])
            elif op.kind == T.OP_ADD:
                y = xin[K] + self.view(self.fw, K, op.inp2)
                if relu:
                    y = self.decide(oi, y) if K == 0 else y * self.masks[oi]
                yout.copy_(y)
            else:
                raise RuntimeError("op kind %d" % op.kind)
        self.head(K)

    def bn_stats(self, oi, op, K, first):
        """channel jets mu, r exactly as bn.cu reduces them (new order only, top mean dropped)."""
        x = [self.view(self.fw, k, op.inp) for k in range(3)]
        st = self.stats.setdefault(oi, {"T": [None] * 3, "Q": [None] * 3})
        N = x[0].shape[0] * x[0].shape[2] * x[0].shape[3]
        dims = (0, 2, 3)
        if K == 0:
            st["T"][0] = x[0].sum(dims)
            st["Q"][0] = (x[0] * x[0]).sum(dims)
        elif first:
            st["T"][K] = torch.zeros_like(st["T"][0])
            st["Q"][K] = torch.zeros_like(st["T"][0])
        elif K == 1:
            mu0 = (st["T"][0] / N).view(1, -1, 1, 1)
            st["T"][1] = x[1].sum(dims)
            st["Q"][1] = (2 * (x[0] - mu0) * x[1]).sum(dims)
        else:
            mu0 = (st["T"][0] / N).view(1, -1, 1, 1)
            mu1 = (st["T"][1] / N).view(1, -1, 1, 1)
            c1 = x[1] - mu1
            st["T"][2] = x[2].sum(dims)
            st["Q"][2] = (2 * (c1 * c1 + (x[0] - mu0) * x[2])).sum(dims)
        return N

    def bn_channel(self, oi, op, K, N):
        st = self.stats[oi]
        mu = [(st["T"][k] / N).view(1, -1, 1, 1) for k in range(K + 1)]
        s = [(st["Q"][k] / N).view(1, -1, 1, 1) for k in range(K + 1)]
        s[0] = s[0] - mu[0] * mu[0] + op.eps
        r = jrsqrt(s, K)
        C = mu[0].numel()
        z = torch.zeros(1, C, 1, 1, dtype=DT)
        gam = [self.w[op.w_off:op.w_off + C].view(1, -1, 1, 1), self.v[op.w_off:op.w_off + C].view(1, -1, 1, 1), z]
        bet = [self.w[op.b_off:op.b_off + C].view(1, -1, 1, 1), self.v[op.b_off:op.b_off + C].view(1, -1, 1, 1), z]
        return mu, r, gam[:K + 1] + [z] * (2 - K), bet[:K + 1] + [z] * (2 - K)

    def bn_forward(self, oi, op, K, first, relu):
        N = self.bn_stats(oi, op, K, first)
        mu, r, gam, bet = self.bn_channel(oi, op, K, N)
        x = [self.view(self.fw, k, op.inp) for k in range(K + 1)]
        xh = jmul(jsub(x, mu, K), r, K)
        y = jadd(jmul(gam, xh, K), bet, K)[K]
        if relu:
            y = self.decide(oi, y) if K == 0 else y * self.masks[oi]
        self.view(self.fw, K, op.out).copy_(y)
        if K == 0:
            var = self.stats[oi]["Q"][0] / N - (self.stats[oi]["T"][0] / N) ** 2
            self.stats[oi]["batch_mean"] = self.stats[oi]["T"][0] / N
            self.stats[oi]["batch_var_unbiased"] = var * N / max(N - 1, 1)

    def head(self, K):
        tp = self.tape
        z = [self.view(self.fw, k, tp.logits).reshape(self.B, -1) for k in range(K + 1)]
        C = z[0].shape[1]
        kind = tp.head
        if kind in (T.HEAD_CE, T.HEAD_SOFTMAX_CE):
            onehot = F.one_hot(self.target.long(), C).to(DT)
            if kind == T.HEAD_CE:
                p = jsoftmax(z, K)
                zb = p[K] - (onehot if K == 0 else 0)
                if K == 0:
                    self.loss = -(torch.log_softmax(z[0], 1) * onehot).sum() * self.loss_scale
            else:
                p = jsoftmax(z, K)
                q = jsoftmax(p, K)
                q = [q[0] - onehot] + list(q[1:])
                dot = [x.sum(1, keepdim=True) for x in jmul(q, p, K)]
                zb = jmul(p, jsub(q, dot, K), K)[K]
                if K == 0:
                    self.loss = -(torch.log_softmax(p[0], 1) * onehot).sum() * self.loss_scale
            zb = zb * self.loss_scale
        else:
            t = self.target.to(DT)
            cf = self.coef.to(DT)
            if kind == T.HEAD_WBCE:
                s = jsigmoid(z, K)
                zb = cf * (s[K] - (t if K == 0 else 0))
                if K == 0:
                    self.loss = (cf * F.binary_cross_entropy_with_logits(z[0], t, reduction="none")).sum()
            else:
                s = jsigmoid(z, K)
                q = jsigmoid(s, K)
                q = [q[0] - t] + list(q[1:])
                one_minus = [1.0 - s[0]] + [-x for x in s[1:]]
                zb = cf * jmul(jmul(q, s, K), one_minus, K)[K]
                if K == 0:
                    self.loss = (cf * F.binary_cross_entropy_with_logits(s[0], t, reduction="none")).sum()
        self.view(self.bw, K, tp.logits).copy_(zb.view(self.B, *tp.tensors[tp.logits].shape))

    def backward(self, K):
        tp = self.tape
        out = self.out[K]
        out.zero_()
        binom = {0: [1], 1: [1, 1], 2: [1, 2, 1]}[K]
        for oi in range(len(tp.ops) - 1, -1, -1):
            op = tp.ops[oi]
            first = bool(op.flags & T.F_FIRST)
            relu = bool(op.flags & T.F_RELU)
            acc = bool(op.flags & T.F_BWD_ACC)
            g = [self.view(self.bw, k, op.out) for k in range(3)]
            x = [self.view(self.fw, k, op.inp) for k in range(3)]
            xbar = self.view(self.bw, K, op.inp)

            def emit(val):
                if acc:
                    xbar.add_(val)
                else:
                    xbar.copy_(val)

            if op.kind == T.OP_CONV:
                cin = tp.tensors[op.inp].shape[0]
                cout = tp.tensors[op.out].shape[0]
                kh, kw, sh, sw, ph, pw = op.geom
                if relu:
                    g[K].mul_(self.masks[oi])
                W = self.wslice(self.w, op.w_off, (cout, cin, kh, kw))
                V = self.wslice(self.v, op.w_off, (cout, cin, kh, kw))
                wshape = (cout, cin, kh, kw)
                wg = torch.zeros(wshape, dtype=DT)
                for j in range(K + 1):        # sum_j binom(K,j) corr(x_j, g_{K-j})
                    if j > 0 and first:
                        continue
                    wg += binom[j] * torch.nn.grad.conv2d_weight(x[j], wshape, g[K - j], (sh, sw), (ph, pw))
                out[op.w_off:op.w_off + wg.numel()] += wg.reshape(-1)
                if op.b_off >= 0:
                    out[op.b_off:op.b_off + cout] += g[K].sum((0, 2, 3))
                if not first:
                    ishape = x[0].shape
                    dg = torch.nn.grad.conv2d_input(ishape, W, g[K], (sh, sw), (ph, pw))
                    if K >= 1:
                        dg = dg + binom[1] * torch.nn.grad.conv2d_input(ishape, V, g[K - 1], (sh, sw), (ph, pw))
                    emit(dg)
            elif op.kind == T.OP_BN:
                N = x[0].shape[0] * x[0].shape[2] * x[0].shape[3]
                mu, r, gam, bet = self.bn_channel(oi, op, K, N)
                xj = [x[0]] + ([torch.zeros_like(x[0])] * 2 if first else [x[1], x[2]])
                xh = jmul(jsub(xj[:K + 1], mu, K), r, K)
                mask = self.masks[oi] if relu else 1.0
                gm = [g[k] * mask for k in range(K + 1)]
                st = self.stats[oi].setdefault("bw", {"G": [None] * 3, "X": [None] * 3})
                st["G"][K] = gm[K].sum((0, 2, 3))
                st["X"][K] = jmul(gm, xh, K)[K].sum((0, 2, 3))
                C = st["G"][K].numel()
                out[op.b_off:op.b_off + C] += st["G"][K]
                out[op.w_off:op.w_off + C] += st["X"][K]
                if not first:
                    Gj = [(st["G"][k] / N).view(1, -1, 1, 1) for k in range(K + 1)]
                    Xj = [(st["X"][k] / N).view(1, -1, 1, 1) for k in range(K + 1)]
                    m1, m2 = jmul(gam, Gj, K), jmul(gam, Xj, K)
                    u = jsub(jsub(jmul(gam, gm, K), m1, K), jmul(xh, m2, K), K)
                    emit(jmul(r, u, K)[K])
            elif op.kind == T.OP_RELU:
                if not first:
                    emit(g[K] * self.masks[oi])
            elif op.kind == T.OP_MAXPOOL:
                if not first:
                    idx = self.argmax[oi]
                    val = torch.zeros_like(xbar).flatten(2).scatter_add_(2, idx.flatten(2), g[K].flatten(2)).view_as(xbar)
                    emit(val)
            elif op.kind == T.OP_AVGPOOL:
                if not first:
                    k = op.geom[0]
                    up = torch.zeros_like(xbar)
                    oh, ow = g[K].shape[2], g[K].shape[3]
                    up[:, :, :oh * k, :ow * k] = g[K].repeat_interleave(k, 2).repeat_interleave(k, 3) / (k * k)
                    emit(up)
            elif op.kind == T.OP_COPY:
                if not first:
                    emit(g[K])
            elif op.kind == T.OP_ADD:
                val = g[K] * self.masks[oi] if relu else g[K]
                if not first:
                    emit(val)
                self._emit2(self.view(self.bw, K, op.inp2), op, val)

    def _emit2(self, xbar2, op, val):
        if self.tape.tensors[op.inp2].buf == self.tape.tensors[0].buf:
            return
        if op.flags & T.F_BWD_ACC2:
            xbar2.add_(val)
        else:
            xbar2.copy_(val)

    # ---- reference-compatible third order through BatchNorm -------------------------------------
    def bn_third_order_defect(self):
        """What the reference's vGHv LACKS relative to the exact grad_w(v^T H v) on BatchNorm models.

        torch differentiates ``native_batch_norm_backward`` with ``batchnorm_double_backward``
        (torch/csrc/autograd/FunctionsManual.cpp), which reads the batch mean and inverse std from
        the *saved, non-differentiable* outputs of the forward.  The second sweep (Hv) is exact
        because the formula accounts for mu(x), r(x) analytically, but when the third sweep of
        ``HVPOperator.vGHv`` (opt.py:143) differentiates those expressions again, mu and r are
        constants: every path  phi = v^T H v  ->  (mu, r inside the double-backward node)  ->  x
        is dropped.  With Psi(mu, r) = <xdot, gI> + <gammadot, gG> + <R ybar, ggO> (the double
        backward's three outputs contracted with their third-sweep adjoints), the dropped adjoint
        injected at the BN input is  m = dPsi/dmu * dmu/dx + dPsi/dr * dr/dx, and it travels on to
        the earlier layers by an ordinary first-order backward sweep.  Must be called after
        run(0), run(1, v).  reference vGHv = run(2) - this.
        """
        tp = self.tape
        out = torch.zeros(self.P, dtype=DT)
        cw = [torch.zeros_like(b) for b in self.bw[0]]

        def cview(t):
            vt = tp.tensors[t]
            c, h, w = vt.shape
            return cw[vt.buf][:, vt.offset:vt.offset + vt.numel].view(self.B, c, h, w)

        sm = lambda t: t.sum((0, 2, 3), keepdim=True)   # noqa: E731
        for oi in range(len(tp.ops) - 1, -1, -1):
            op = tp.ops[oi]
            first = bool(op.flags & T.F_FIRST)
            relu = bool(op.flags & T.F_RELU)
            acc = bool(op.flags & T.F_BWD_ACC)
            g = cview(op.out)
            x0 = self.view(self.fw, 0, op.inp)
            xbar = cview(op.inp)

            def emit(val):
                if acc:
                    xbar.add_(val)
                else:
                    xbar.copy_(val)

            if op.kind == T.OP_CONV:
                cin = tp.tensors[op.inp].shape[0]
                cout = tp.tensors[op.out].shape[0]
                kh, kw, sh, sw, ph, pw = op.geom
                if relu:
                    g.mul_(self.masks[oi])
                W = self.wslice(self.w, op.w_off, (cout, cin, kh, kw))
                wg = torch.nn.grad.conv2d_weight(x0, (cout, cin, kh, kw), g, (sh, sw), (ph, pw))
                out[op.w_off:op.w_off + wg.numel()] += wg.reshape(-1)
                if op.b_off >= 0:
                    out[op.b_off:op.b_off + cout] += g.sum((0, 2, 3))
                if not first:
                    emit(torch.nn.grad.conv2d_input(x0.shape, W, g, (sh, sw), (ph, pw)))
            elif op.kind == T.OP_BN:
                N = x0.shape[0] * x0.shape[2] * x0.shape[3]
                mu, r, gam, bet = self.bn_channel(oi, op, 1, N)
                r0, ga, gd = r[0], gam[0], gam[1]
                mask = self.masks[oi] if relu else 1.0
                yb = self.view(self.bw, 0, op.out) * mask       # gO
                h = self.view(self.bw, 1, op.out) * mask        # R ybar
                xd = torch.zeros_like(x0) if first else self.view(self.fw, 1, op.inp)
                c = x0 - mu[0]
                A, D, Bc, Ec, S, X2 = sm(yb), sm(xd), sm(yb * c), sm(xd * c), sm(yb * xd), sm(xd * xd)
                Hs, Hc, Hx = sm(h), sm(h * c), sm(h * xd)
                Tt = A * D / N - S
                r2, r3 = r0 ** 2, r0 ** 3
                dmu = (ga * (r3 / N) * (-2 * Tt * D - (3 * r2 / N) * (A * Ec ** 2 + 2 * Bc * Ec * D) - A * (D * D / N - X2))
                       + 2 * gd * (r3 / N) * (A * Ec + Bc * D) + ga * (r3 / N) * (D * Hc + Ec * Hs) - gd * r0 * Hs)
                dr = (ga * (3 * r2 / N) * (2 * Ec * Tt + Bc * (D * D / N - X2)) + ga * (5 * r0 ** 4 / N) * (3 * Bc * Ec ** 2 / N)
                      + 2 * gd * (-Tt - 3 * r2 * Bc * Ec / N) + (ga / N) * (N * Hx - D * Hs) - ga * (3 * r2 / N) * Ec * Hc
                      + gd * Hc)
                gm = g * mask
                xh = c * r0
                G, X = sm(gm), sm(gm * xh)
                C = G.numel()
                out[op.b_off:op.b_off + C] += G.view(-1)
                out[op.w_off:op.w_off + C] += X.view(-1)
                if not first:
                    emit(r0 * ga * (gm - G / N - xh * X / N) + dmu / N - dr * r3 * c / N)
            elif op.kind == T.OP_RELU:
                if not first:
                    emit(g * self.masks[oi])
            elif op.kind == T.OP_MAXPOOL:
                if not first:
                    idx = self.argmax[oi]
                    emit(torch.zeros_like(xbar).flatten(2).scatter_add_(2, idx.flatten(2), g.flatten(2)).view_as(xbar))
            elif op.kind == T.OP_AVGPOOL:
                if not first:
                    k = op.geom[0]
                    up = torch.zeros_like(xbar)
                    oh, ow = g.shape[2], g.shape[3]
                    up[:, :, :oh * k, :ow * k] = g.repeat_interleave(k, 2).repeat_interleave(k, 3) / (k * k)
                    emit(up)
            elif op.kind == T.OP_COPY:
                if not first:
                    emit(g)
            elif op.kind == T.OP_ADD:
                val = g * self.masks[oi] if relu else g
                if not first:
                    emit(val)
                self._emit2(cview(op.inp2), op, val)
        return out

    def vghv(self, v, mode="reference"):
        """grad_w(v^T H v): 'exact' = the true third derivative; 'reference' = what nested torch.autograd
        returns (identical unless the model has train-mode BatchNorm)."""
        self.run(1, v)
        exact = self.run(2).clone()
        if mode == "exact" or not any(op.kind == T.OP_BN for op in self.tape.ops):
            return exact
        return exact - self.bn_third_order_defect()
