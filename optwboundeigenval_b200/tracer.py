"""nn.Module + loss  ->  static tape for libb200spectral.

The reference hands an arbitrary ``nn.Module`` and a criterion to ``HVPOperator``
(opt.py:55-62) and lets autograd discover the graph on every minibatch
(opt.py:181-192).  The B200 path needs the topology once: ``torch.fx`` extracts
the module graph on the host (no tensor data flows), and this file lowers it to
the op/tensor tables of ``include/b200_spectral.h``:

* parameters get offsets in ``model.parameters()`` order (opt.py:102,191),
  shared parameters once (forest_data.py:85-86 applies ``fc2`` twice);
* ReLUs are fused into their producer (Conv/Linear/BatchNorm);
* ``torch.cat`` along channels becomes aliasing: operands are views into the
  concatenated buffer (densenet.py:21,42; torchvision ``_DenseBlock``);
* a trailing softmax / sigmoid is folded into the loss head
  (forest_data.py:88, usps_data.py:335, dcnn.py:275);
* for the backward sweep each op is told whether it is the first writer of its
  input adjoint (overwrite) or a later one (accumulate).

Unsupported modules raise ``UnsupportedModel`` naming the node -- there is no
CPU or autograd fallback.
"""
from __future__ import annotations

import ctypes
import operator
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import torch
import torch.fx as fx
import torch.nn as nn
import torch.nn.functional as F

OP_CONV, OP_BN, OP_RELU, OP_MAXPOOL, OP_AVGPOOL, OP_COPY, OP_ADD = 1, 2, 3, 4, 5, 6, 7
F_RELU, F_FIRST, F_BWD_ACC, F_BWD_ACC2 = 1, 2, 4, 8
HEAD_CE, HEAD_SOFTMAX_CE, HEAD_WBCE, HEAD_SIGMOID_WBCE = 1, 2, 3, 4


class UnsupportedModel(RuntimeError):
    pass


class CTensor(ctypes.Structure):
    _fields_ = [("buf", ctypes.c_int32), ("C", ctypes.c_int32), ("H", ctypes.c_int32), ("W", ctypes.c_int32),
                ("offset", ctypes.c_int64), ("sample_stride", ctypes.c_int64)]


class COp(ctypes.Structure):
    _fields_ = [("kind", ctypes.c_int32), ("flags", ctypes.c_int32), ("in_", ctypes.c_int32),
                ("out", ctypes.c_int32), ("w_off", ctypes.c_int64), ("b_off", ctypes.c_int64),
                ("kh", ctypes.c_int32), ("kw", ctypes.c_int32), ("sh", ctypes.c_int32), ("sw", ctypes.c_int32),
                ("ph", ctypes.c_int32), ("pw", ctypes.c_int32), ("slot", ctypes.c_int32),
                ("eps", ctypes.c_float), ("momentum", ctypes.c_float)]


@dataclass
class VT:
    """virtual tensor: per-sample shape, later placed into a buffer"""
    shape: Tuple[int, int, int]
    producer: Optional[int] = None       # index into vops
    users: int = 0
    alias_of: Optional[int] = None       # a pure reshape of another tensor (same storage)
    buf: int = -1
    offset: int = 0

    @property
    def numel(self):
        c, h, w = self.shape
        return c * h * w


@dataclass
class VOp:
    kind: int
    inp: int
    out: int
    flags: int = 0
    w_off: int = -1
    b_off: int = -1
    geom: Tuple[int, int, int, int, int, int] = (1, 1, 1, 1, 0, 0)
    slot: int = -1
    eps: float = 0.0
    momentum: float = 0.0
    name: str = ""
    module: Optional[nn.Module] = None
    inp2: int = -1                       # OP_ADD: the second operand


@dataclass
class Tape:
    tensors: List[VT]
    ops: List[VOp]
    buf_elems: List[int]
    logits: int
    head: int
    n_params: int
    input_shape: Tuple[int, ...]
    bn_modules: List[nn.Module] = field(default_factory=list)
    param_offsets: Dict[int, int] = field(default_factory=dict)
    conv_layers: List[int] = field(default_factory=list)     # op indices of Conv/Linear, first use per module
    # KLDivLoss on a one-hot target (opt.py:182-185,566-569): cross entropy of the logits times this reduction factor
    kl_reduction: Optional[str] = None

    def head_scale(self, batch: int, classes: int) -> float:
        """factor on the SUM over samples of the per-sample cross entropy"""
        if self.kl_reduction is None or self.kl_reduction == "batchmean":
            return 1.0 / float(batch)
        if self.kl_reduction == "mean":                       # nn.KLDivLoss default: mean over all B*C elements
            return 1.0 / (float(batch) * float(classes))
        return 1.0                                            # "sum"

    def c_tensors(self):
        arr = (CTensor * len(self.tensors))()
        for i, t in enumerate(self.tensors):
            c, h, w = t.shape
            arr[i] = CTensor(t.buf, c, h, w, t.offset, self.buf_elems[t.buf])
        return arr

    def c_ops(self):
        arr = (COp * len(self.ops))()
        for i, o in enumerate(self.ops):
            kh, kw, sh, sw, ph, pw = o.geom
            arr[i] = COp(o.kind, o.flags, o.inp, o.out, o.w_off, o.b_off, kh, kw, sh, sw, ph, pw,
                         o.inp2 if o.kind == OP_ADD else o.slot, o.eps, o.momentum)
        return arr

    def describe(self) -> str:
        names = {OP_CONV: "conv", OP_BN: "bn", OP_RELU: "relu", OP_MAXPOOL: "maxpool", OP_AVGPOOL: "avgpool",
                 OP_COPY: "copy", OP_ADD: "add"}
        lines = []
        for i, o in enumerate(self.ops):
            ti, to = self.tensors[o.inp], self.tensors[o.out]
            lines.append("%3d %-8s %-28s in=t%d%s@b%d+%d out=t%d%s@b%d+%d flags=%d" % (
                i, names[o.kind], o.name, o.inp, ti.shape, ti.buf, ti.offset, o.out, to.shape, to.buf, to.offset,
                o.flags))
        return "\n".join(lines)


def _pair(v):
    return (v, v) if isinstance(v, int) else tuple(v)


def head_kind(criterion, tail: Optional[str]) -> int:
    name = criterion.__class__.__name__
    if tail == "log_softmax" and name != "KLDivLoss":
        raise UnsupportedModel("log_softmax output is supported with KLDivLoss only")
    if name == "CrossEntropyLoss":
        if getattr(criterion, "reduction", "mean") != "mean" or getattr(criterion, "weight", None) is not None \
                or getattr(criterion, "label_smoothing", 0.0) != 0.0 or getattr(criterion, "ignore_index", -100) != -100:
            raise UnsupportedModel("CrossEntropyLoss is supported with default arguments only")
        if tail == "sigmoid":
            raise UnsupportedModel("sigmoid output followed by CrossEntropyLoss is not supported")
        return HEAD_SOFTMAX_CE if tail == "softmax" else HEAD_CE
    if name in ("W_BCEWithLogitsLoss", "WeightedBCEWithLogits"):
        if tail == "softmax":
            raise UnsupportedModel("softmax output followed by the weighted BCE loss is not supported")
        return HEAD_SIGMOID_WBCE if tail == "sigmoid" else HEAD_WBCE
    if name == "KLDivLoss":
        # opt.py:182-185: criterion(output, one_hot(target)).  With log-probabilities as the model output (what
        # KLDivLoss takes) and a one-hot target, sum_c t*(log t - out) = -log_softmax(z)[y]: the cross-entropy head on
        # the logits, scaled by the loss's reduction (Tape.head_scale).
        if tail != "log_softmax":
            raise UnsupportedModel("KLDivLoss needs a model that ends in log_softmax (it takes log-probabilities)")
        if getattr(criterion, "log_target", False) or getattr(criterion, "reduction", "mean") not in ("mean", "batchmean", "sum"):
            raise UnsupportedModel("KLDivLoss is supported with log_target=False and reduction mean / batchmean / sum")
        return HEAD_CE
    raise UnsupportedModel("loss %s has no B200 head (supported: CrossEntropyLoss, W_BCEWithLogitsLoss, KLDivLoss on "
                           "log-probabilities)" % name)


class _Builder:
    def __init__(self, model: nn.Module, input_shape):
        self.model = model
        self.vts: List[VT] = []
        self.vops: List[VOp] = []
        self.poff: Dict[int, int] = {}
        off = 0
        for p in model.parameters():
            self.poff[id(p)] = off
            off += p.numel()
        self.n_params = off
        self.bn_modules: List[nn.Module] = []
        self.cats: List[Tuple[int, List[int]]] = []      # (output vt, input vts)
        shp = tuple(int(s) for s in input_shape)
        if len(shp) == 1:
            shp = (shp[0], 1, 1)
        elif len(shp) == 2:
            shp = (1,) + shp
        if len(shp) != 3:
            raise UnsupportedModel("input must be [B,F], [B,H,W] or [B,C,H,W]; got per-sample shape %r" % (input_shape,))
        self.new_vt(shp)      # tensor 0 = network input

    def new_vt(self, shape, producer=None, alias_of=None) -> int:
        self.vts.append(VT(tuple(int(s) for s in shape), producer, 0, alias_of))
        return len(self.vts) - 1

    def root(self, t: int) -> int:
        while self.vts[t].alias_of is not None:
            t = self.vts[t].alias_of
        return t

    def add_op(self, kind, inp, out_shape, **kw) -> int:
        out = self.new_vt(out_shape, producer=len(self.vops))
        self.vops.append(VOp(kind, inp, out, **kw))
        self.vts[self.root(inp)].users += 1
        return out

    # ---- layer lowering ------------------------------------------------------------------
    def conv(self, t, m: nn.Conv2d, name):
        c, h, w = self.vts[t].shape
        if m.groups != 1 or _pair(m.dilation) != (1, 1) or m.padding_mode != "zeros":
            raise UnsupportedModel("%s: grouped / dilated / non-zero-padded convolutions are not supported" % name)
        if c != m.in_channels:
            raise UnsupportedModel("%s: expected %d input channels, got %d" % (name, m.in_channels, c))
        kh, kw = _pair(m.kernel_size)
        sh, sw = _pair(m.stride)
        if isinstance(m.padding, str):                       # torch's spellings of the two symmetric cases
            if m.padding == "valid":
                ph = pw = 0
            elif m.padding == "same" and kh % 2 == 1 and kw % 2 == 1 and (sh, sw) == (1, 1):
                ph, pw = kh // 2, kw // 2
            else:
                raise UnsupportedModel("%s: padding=%r needs asymmetric padding" % (name, m.padding))
        else:
            ph, pw = _pair(m.padding)
        oh = (h + 2 * ph - kh) // sh + 1
        ow = (w + 2 * pw - kw) // sw + 1
        return self.add_op(OP_CONV, t, (m.out_channels, oh, ow), w_off=self.poff[id(m.weight)],
                           b_off=self.poff[id(m.bias)] if m.bias is not None else -1,
                           geom=(kh, kw, sh, sw, ph, pw), name=name, module=m)

    def linear(self, t, m, name):
        c, h, w = self.vts[t].shape
        n_out, n_in = m._parameters["weight"].shape   # nn.Linear and dnet._linear (dnet.py:105-143, LinearFunction) alike
        if c * h * w != n_in:
            raise UnsupportedModel("%s: expected %d features, got %r" % (name, n_in, (c, h, w)))
        if (h, w) != (1, 1):
            t = self.reshape(t, (c * h * w, 1, 1))
        return self.add_op(OP_CONV, t, (n_out, 1, 1), w_off=self.poff[id(m.weight)],
                           b_off=self.poff[id(m.bias)] if m.bias is not None else -1, name=name, module=m)

    def bn(self, t, m, name):
        c, h, w = self.vts[t].shape
        if not m.affine or not m.track_running_stats or m.momentum is None:
            raise UnsupportedModel("%s: BatchNorm needs affine=True, track_running_stats=True, momentum set" % name)
        if c != m.num_features:
            raise UnsupportedModel("%s: expected %d channels, got %d" % (name, m.num_features, c))
        slot = len(self.bn_modules)
        self.bn_modules.append(m)
        return self.add_op(OP_BN, t, (c, h, w), w_off=self.poff[id(m.weight)], b_off=self.poff[id(m.bias)],
                           slot=slot, eps=float(m.eps), momentum=float(m.momentum), name=name, module=m)

    def relu(self, t, name, inplace=False):
        return self.add_op(OP_RELU, t, self.vts[t].shape, name=name, geom=(1 if inplace else 0, 0, 0, 0, 0, 0))

    def maxpool(self, t, k, s, p, name, dilation=1, ceil_mode=False, return_indices=False):
        if _pair(dilation) != (1, 1) or ceil_mode or return_indices:
            raise UnsupportedModel("%s: dilated / ceil_mode / return_indices max pooling is not supported" % name)
        kh, kw = _pair(k)
        sh, sw = _pair(s if s is not None else k)
        ph, pw = _pair(p)
        c, h, w = self.vts[t].shape
        oh = (h + 2 * ph - kh) // sh + 1
        ow = (w + 2 * pw - kw) // sw + 1
        return self.add_op(OP_MAXPOOL, t, (c, oh, ow), geom=(kh, kw, sh, sw, ph, pw), name=name)

    def avgpool(self, t, k, s, p, name, ceil_mode=False):
        kh, kw = _pair(k)
        sh, sw = _pair(s if s is not None else k)
        ph, pw = _pair(p)
        if kh != kw or (sh, sw) != (kh, kw) or (ph, pw) != (0, 0) or ceil_mode:
            raise UnsupportedModel("%s: average pooling needs a square kernel with stride == kernel, no padding" % name)
        c, h, w = self.vts[t].shape
        return self.add_op(OP_AVGPOOL, t, (c, h // kh, w // kw), geom=(kh, kw, kh, kw, 0, 0), name=name)

    def add(self, t1, t2, name):
        """residual connection (torchvision Bottleneck: ``out += identity``, dcnn.py:222-225 MyResNet50)"""
        if self.vts[t1].shape != self.vts[t2].shape:
            raise UnsupportedModel("%s: element-wise add of tensors with different shapes %r / %r (no broadcasting)" % (
                name, self.vts[t1].shape, self.vts[t2].shape))
        out = self.add_op(OP_ADD, t1, self.vts[t1].shape, name=name, inp2=t2)
        self.vts[self.root(t2)].users += 1
        return out

    def reshape(self, t, shape):
        if self.vts[t].numel != shape[0] * shape[1] * shape[2]:
            raise UnsupportedModel("reshape changes the per-sample element count: %r -> %r" % (self.vts[t].shape, shape))
        if tuple(shape) == self.vts[t].shape:
            return t
        return self.new_vt(shape, alias_of=t)

    def cat(self, ts: List[int], name):
        shapes = [self.vts[t].shape for t in ts]
        if len({s[1:] for s in shapes}) != 1:
            raise UnsupportedModel("%s: torch.cat operands differ in spatial size" % name)
        if len(ts) == 1:
            return ts[0]
        out = self.new_vt((sum(s[0] for s in shapes),) + shapes[0][1:])
        self.cats.append((out, list(ts)))
        return out


def _is_fn(target, *cands):
    return any(target is c for c in cands)


def _custom_function_kind(m: nn.Module) -> Optional[str]:
    """dnet.py wraps two hand-written autograd Functions in modules: ``_relu`` (MyReLU, dnet.py:30-60) and ``_linear``
    (LinearFunction, dnet.py:64-143).  Their backward passes are written with differentiable tensor ops, so nested
    autograd differentiates them like the built-in layers: ``_linear`` is exactly ``nn.Linear``; ``MyReLU`` is ReLU
    except that its gradient passes where the input is exactly 0.0 (``grad_input[input < 0] = 0``), a set that is
    empty for floating-point pre-activations of BatchNorm outputs -- it is lowered to the same mask kernels."""
    name = m.__class__.__name__
    params = m.__dict__.get("_parameters", {})        # not getattr: fx turns parameter access into proxies while tracing
    if name == "_relu" and "f" in m.__dict__ and not params:
        return "relu"
    if name == "_linear" and params.get("weight") is not None and params["weight"].dim() == 2:
        return "linear"
    return None


class _Tracer(fx.Tracer):
    def is_leaf_module(self, m, qualname):
        return _custom_function_kind(m) is not None or super().is_leaf_module(m, qualname)


def trace(model: nn.Module, criterion, input_shape) -> Tape:
    """Lower ``model`` (+ its loss) to a tape. ``input_shape`` is the per-sample input shape."""
    was_training = model.training
    try:
        graph = _Tracer().trace(model)
        gm = fx.GraphModule(model, graph)
    except Exception as exc:   # noqa: BLE001
        raise UnsupportedModel("torch.fx could not trace %s: %s" % (model.__class__.__name__, exc)) from exc
    finally:
        model.train(was_training)
    b = _Builder(model, input_shape)
    env: Dict[fx.Node, object] = {}
    tail = None
    result = None
    BATCH = "batch"
    for node in gm.graph.nodes:
        nm = node.name
        if node.op == "placeholder":
            if env:
                raise UnsupportedModel("models with more than one input are not supported")
            env[node] = 0
            continue
        if node.op == "output":
            result = env[node.args[0]] if isinstance(node.args[0], fx.Node) else None
            if not isinstance(result, int):
                raise UnsupportedModel("the model must return a single tensor")
            continue
        if tail is not None:
            raise UnsupportedModel("%s: softmax / log_softmax / sigmoid is supported only as the last op of the model" % nm)

        def arg(i, key=None, default=None):
            if len(node.args) > i:
                return node.args[i]
            return node.kwargs.get(key, default)

        def tin(a):
            v = env[a]
            if not isinstance(v, int):
                raise UnsupportedModel("%s: expected a tensor operand" % nm)
            return v

        if node.op == "call_module":
            m = gm.get_submodule(node.target)
            t = tin(node.args[0])
            if isinstance(m, nn.Conv2d):
                env[node] = b.conv(t, m, node.target)
            elif isinstance(m, nn.Linear):
                env[node] = b.linear(t, m, node.target)
            elif isinstance(m, (nn.BatchNorm2d, nn.BatchNorm1d)):
                env[node] = b.bn(t, m, node.target)
            elif isinstance(m, nn.ReLU):
                env[node] = b.relu(t, node.target, m.inplace)
            elif _custom_function_kind(m) == "relu":
                env[node] = b.relu(t, node.target, False)
            elif _custom_function_kind(m) == "linear":
                env[node] = b.linear(t, m, node.target)
            elif isinstance(m, nn.MaxPool2d):
                env[node] = b.maxpool(t, m.kernel_size, m.stride, m.padding, node.target, m.dilation, m.ceil_mode,
                                      m.return_indices)
            elif isinstance(m, nn.AvgPool2d):
                if not m.count_include_pad and _pair(m.padding) != (0, 0):
                    raise UnsupportedModel("%s: count_include_pad=False with padding" % nm)
                env[node] = b.avgpool(t, m.kernel_size, m.stride, m.padding, node.target, m.ceil_mode)
            elif isinstance(m, nn.AdaptiveAvgPool2d):
                if _pair(m.output_size) != (1, 1):
                    raise UnsupportedModel("%s: adaptive average pooling is supported to (1,1) only" % nm)
                env[node] = b.avgpool(t, b.vts[t].shape[1:], None, 0, node.target) if b.vts[t].shape[1] == b.vts[t].shape[2] \
                    else _unsupported(nm, "adaptive pooling of a non-square map")
            elif isinstance(m, (nn.Dropout, nn.Dropout2d)):
                if m.p > 0:
                    raise UnsupportedModel("%s: dropout with p > 0 is not supported in the spectral passes" % nm)
                env[node] = t
            elif isinstance(m, nn.Identity):
                env[node] = t
            elif isinstance(m, nn.Flatten):
                env[node] = b.reshape(t, (b.vts[t].numel, 1, 1))
            elif isinstance(m, nn.Softmax):
                if m.dim not in (1, -1):
                    raise UnsupportedModel("%s: softmax over dim %r" % (nm, m.dim))
                env[node] = t
                tail = "softmax"
            elif isinstance(m, nn.Sigmoid):
                env[node] = t
                tail = "sigmoid"
            elif isinstance(m, nn.LogSoftmax):
                if m.dim not in (1, -1):
                    raise UnsupportedModel("%s: log_softmax over dim %r" % (nm, m.dim))
                env[node] = t
                tail = "log_softmax"
            else:
                raise UnsupportedModel("module %s (%s) has no B200 kernel" % (node.target, m.__class__.__name__))
        elif node.op == "call_function":
            tg = node.target
            if _is_fn(tg, torch.relu, F.relu, torch.relu_, F.relu_):
                env[node] = b.relu(tin(node.args[0]), nm, bool(arg(1, "inplace", False)) or tg in (torch.relu_, F.relu_))
            elif _is_fn(tg, operator.add, operator.iadd, torch.add) and len(node.args) == 2 \
                    and all(isinstance(a, fx.Node) and isinstance(env.get(a), int) for a in node.args) and not node.kwargs:
                env[node] = b.add(tin(node.args[0]), tin(node.args[1]), nm)
            elif _is_fn(tg, torch.cat, torch.concat):
                dim = arg(1, "dim", 0)
                if dim != 1:
                    raise UnsupportedModel("%s: torch.cat is supported along dim 1 only" % nm)
                env[node] = b.cat([tin(a) for a in node.args[0]], nm)
            elif _is_fn(tg, F.avg_pool2d):
                env[node] = b.avgpool(tin(node.args[0]), arg(1, "kernel_size"), arg(2, "stride", None),
                                      arg(3, "padding", 0), nm, arg(4, "ceil_mode", False))
            elif _is_fn(tg, F.max_pool2d):
                env[node] = b.maxpool(tin(node.args[0]), arg(1, "kernel_size"), arg(2, "stride", None),
                                      arg(3, "padding", 0), nm, arg(4, "dilation", 1), arg(5, "ceil_mode", False),
                                      arg(6, "return_indices", False))
            elif _is_fn(tg, F.adaptive_avg_pool2d):
                t = tin(node.args[0])
                if _pair(arg(1, "output_size")) != (1, 1) or b.vts[t].shape[1] != b.vts[t].shape[2]:
                    raise UnsupportedModel("%s: adaptive average pooling is supported to (1,1) on square maps" % nm)
                env[node] = b.avgpool(t, b.vts[t].shape[1], None, 0, nm)
            elif _is_fn(tg, torch.flatten):
                t = tin(node.args[0])
                if arg(1, "start_dim", 0) != 1:
                    raise UnsupportedModel("%s: flatten must keep the batch dimension" % nm)
                env[node] = b.reshape(t, (b.vts[t].numel, 1, 1))
            elif _is_fn(tg, F.softmax, torch.softmax):
                if arg(1, "dim", None) not in (1, -1):
                    raise UnsupportedModel("%s: softmax over dim %r" % (nm, arg(1, "dim", None)))
                env[node] = tin(node.args[0])
                tail = "softmax"
            elif _is_fn(tg, torch.sigmoid, F.sigmoid):
                env[node] = tin(node.args[0])
                tail = "sigmoid"
            elif _is_fn(tg, F.log_softmax, torch.log_softmax):
                if arg(1, "dim", None) not in (1, -1):
                    raise UnsupportedModel("%s: log_softmax over dim %r" % (nm, arg(1, "dim", None)))
                env[node] = tin(node.args[0])
                tail = "log_softmax"
            elif _is_fn(tg, F.dropout):
                if arg(1, "p", 0.5) > 0 and arg(2, "training", True):
                    raise UnsupportedModel("%s: dropout with p > 0" % nm)
                env[node] = tin(node.args[0])
            elif _is_fn(tg, operator.getitem) and env.get(node.args[0]) == "shape" and node.args[1] == 0:
                env[node] = BATCH
            else:
                raise UnsupportedModel("function %s (node %s) has no B200 kernel" % (getattr(tg, "__name__", tg), nm))
        elif node.op == "call_method":
            t0 = node.args[0]
            if node.target in ("view", "reshape"):
                t = tin(t0)
                dims = node.args[1:]
                if len(dims) == 1 and isinstance(dims[0], (tuple, list)):
                    dims = tuple(dims[0])
                env[node] = _lower_view(b, t, dims, env, nm, BATCH)
            elif node.target == "flatten":
                t = tin(t0)
                if arg(1, "start_dim", 0) != 1:
                    raise UnsupportedModel("%s: flatten must keep the batch dimension" % nm)
                env[node] = b.reshape(t, (b.vts[t].numel, 1, 1))
            elif node.target == "size":
                if len(node.args) > 1 and node.args[1] == 0:
                    env[node] = BATCH
                elif len(node.args) == 1:
                    env[node] = "shape"
                else:
                    raise UnsupportedModel("%s: only x.size(0) is supported" % nm)
            elif node.target in ("relu", "relu_"):
                env[node] = b.relu(tin(t0), nm, node.target == "relu_")
            elif node.target == "contiguous":
                env[node] = tin(t0)
            elif node.target in ("add", "add_") and len(node.args) == 2 and isinstance(node.args[1], fx.Node) \
                    and isinstance(env.get(node.args[1]), int) and not node.kwargs:
                env[node] = b.add(tin(t0), tin(node.args[1]), nm)
            else:
                raise UnsupportedModel("tensor method .%s() (node %s) has no B200 kernel" % (node.target, nm))
        elif node.op == "get_attr":
            raise UnsupportedModel("free tensor attribute %s is not supported" % node.target)
        else:
            raise UnsupportedModel("fx node kind %s" % node.op)

    if result is None:
        raise UnsupportedModel("model returns nothing")
    head = head_kind(criterion, tail)
    tape = _finish(b, result, head, input_shape)
    if tail == "log_softmax":
        tape.kl_reduction = getattr(criterion, "reduction", "mean")
    return tape


def _unsupported(nm, why):
    raise UnsupportedModel("%s: %s" % (nm, why))


def _lower_view(b: _Builder, t, dims, env, nm, BATCH):
    """x.view(-1, C, H, W) / x.view(-1, F) / x.view(x.size(0), -1): the batch dimension must survive."""
    vals = []
    for d in dims:
        if isinstance(d, fx.Node):
            d = env[d]
        vals.append(d)
    if not vals:
        _unsupported(nm, "empty view")
    first = vals[0]
    rest = vals[1:]
    numel = b.vts[t].numel
    if not (first == -1 or first == BATCH):
        _unsupported(nm, "view must keep the batch dimension first (got %r)" % (vals,))
    if any(r == BATCH for r in rest):
        _unsupported(nm, "batch size used in a non-leading view dimension")
    known = 1
    unknown = 0
    for r in rest:
        if r == -1:
            unknown += 1
        else:
            known *= int(r)
    if first == -1 and unknown:
        _unsupported(nm, "two inferred dimensions")
    if unknown > 1:
        _unsupported(nm, "two inferred dimensions")
    rest = [numel // known if r == -1 else int(r) for r in rest]
    prod = 1
    for r in rest:
        prod *= r
    if prod != numel:
        _unsupported(nm, "view folds samples together (%d elements per sample -> %r)" % (numel, rest))
    if len(rest) == 1:
        shape = (rest[0], 1, 1)
    elif len(rest) == 3:
        shape = tuple(rest)
    elif len(rest) == 2:
        shape = (1, rest[0], rest[1])
    else:
        _unsupported(nm, "view to %d dimensions" % (len(rest) + 1))
    return b.reshape(t, shape)


def _finish(b: _Builder, result: int, head: int, input_shape) -> Tape:
    vts, vops = b.vts, b.vops

    # ---- fuse ReLU into its producer ------------------------------------------------------
    consumers: Dict[int, int] = {}
    for op in vops:
        r = b.root(op.inp)
        consumers[r] = consumers.get(r, 0) + 1
        if op.inp2 >= 0:
            r = b.root(op.inp2)
            consumers[r] = consumers.get(r, 0) + 1
    for out, ins in b.cats:
        for t in ins:
            r = b.root(t)
            consumers[r] = consumers.get(r, 0) + 1
    res_root = b.root(result)
    consumers[res_root] = consumers.get(res_root, 0) + 1
    replaced: Dict[int, int] = {}         # relu output vt -> producer output vt
    keep = [True] * len(vops)
    for i, op in enumerate(vops):
        if op.kind != OP_RELU:
            continue
        src = op.inp
        if vts[src].alias_of is not None:
            continue
        prod = vts[src].producer
        if prod is None or consumers.get(src, 0) != 1:
            if op.geom[0] and consumers.get(b.root(src), 0) > 1:
                raise UnsupportedModel("%s: in-place ReLU on a tensor that has other consumers" % op.name)
            continue
        p = vops[prod]
        if p.kind in (OP_CONV, OP_BN, OP_ADD) and not (p.flags & F_RELU):
            p.flags |= F_RELU
            keep[i] = False
            replaced[op.out] = src

    def resolve(t):
        while t in replaced:
            t = replaced[t]
        return t

    for op in vops:
        op.inp = resolve(op.inp)
        if op.inp2 >= 0:
            op.inp2 = resolve(op.inp2)
    for vt in vts:
        if vt.alias_of is not None:
            vt.alias_of = resolve(vt.alias_of)
    b.cats = [(out, [resolve(t) for t in ins]) for out, ins in b.cats]
    result = resolve(result)
    ops = [op for i, op in enumerate(vops) if keep[i]]

    # ---- concatenation by aliasing ---------------------------------------------------------
    # placement[t] = (cat_root_vt, channel offset); processed from the last concatenation backwards
    # so that the widest one allocates and the narrower ones become prefixes/views of it.
    placement: Dict[int, Tuple[int, int]] = {}
    copies: List[Tuple[int, int, int]] = []       # (src vt, cat vt, channel offset) physical fallbacks
    for out, ins in reversed(b.cats):
        ins_r = ins
        if out not in placement:
            # is the operand list already laid out contiguously in some buffer? then the cat is a view
            pl = [placement.get(t) for t in ins_r]
            if all(p is not None for p in pl) and len({p[0] for p in pl}) == 1:
                base = pl[0][1]
                ok, run = True, base
                for t, p in zip(ins_r, pl):
                    if p[1] != run:
                        ok = False
                        break
                    run += vts[t].shape[0]
                if ok:
                    placement[out] = (pl[0][0], base)
                    continue
            placement[out] = (out, 0)              # fresh buffer owned by this concatenation
        root, base = placement[out]
        run = base
        for t in ins_r:
            want = (root, run)
            if vts[t].alias_of is not None:
                copies.append((t, out, run - base))
            elif t not in placement:
                placement[t] = want
            elif placement[t] != want:
                copies.append((t, out, run - base))
            run += vts[t].shape[0]

    # ---- buffers -----------------------------------------------------------------------------
    buf_elems: List[int] = []
    buf_of_root: Dict[int, int] = {}

    def place(t: int):
        vt = vts[t]
        if vt.buf >= 0:
            return
        if vt.alias_of is not None:
            place(vt.alias_of)
            src = vts[vt.alias_of]
            vt.buf, vt.offset = src.buf, src.offset
            return
        if t in placement:
            root, choff = placement[t]
            if root not in buf_of_root:
                buf_of_root[root] = len(buf_elems)
                buf_elems.append(vts[root].numel)
            hw = vt.shape[1] * vt.shape[2]
            vt.buf, vt.offset = buf_of_root[root], choff * hw
            return
        vt.buf, vt.offset = len(buf_elems), 0
        buf_elems.append(vt.numel)

    place(0)
    for op in ops:
        place(op.inp)
        if op.inp2 >= 0:
            place(op.inp2)
        place(op.out)
    place(result)
    for t in range(len(vts)):
        if vts[t].buf < 0 and (t in placement):
            place(t)

    # physical copies for concatenations that could not alias (inserted right after the producer)
    if copies:
        new_ops = []
        pending = list(copies)
        produced = {0}
        for op in ops:
            new_ops.append(op)
            produced.add(op.out)
            for c in list(pending):
                src, cat_vt, choff = c
                if b.root(src) in produced or src in produced:
                    hw = vts[src].shape[1] * vts[src].shape[2]
                    dst = len(vts)
                    vts.append(VT(vts[src].shape, buf=vts[cat_vt].buf, offset=vts[cat_vt].offset + choff * hw))
                    new_ops.append(VOp(OP_COPY, src, dst, name="cat_copy"))
                    pending.remove(c)
        if pending:
            raise UnsupportedModel("could not schedule concatenation copies")
        ops = new_ops

    # ---- flags: first layer, backward overwrite/accumulate ------------------------------------
    in_buf = vts[0].buf
    for op in ops:
        if vts[op.inp].buf == in_buf:
            op.flags |= F_FIRST
    written: Dict[int, List[Tuple[int, int]]] = {}
    lg = vts[result]
    written.setdefault(lg.buf, []).append((lg.offset, lg.offset + lg.numel))     # the head writes the logits adjoint
    for op in reversed(ops):
        for operand, flag in ((op.inp, F_BWD_ACC), (op.inp2, F_BWD_ACC2)):
            if operand < 0 or vts[operand].buf == in_buf:      # network data has no adjoint
                continue
            vt = vts[operand]
            lo, hi = vt.offset, vt.offset + vt.numel
            ivs = written.setdefault(vt.buf, [])
            covered = _covered(ivs, lo, hi)
            if covered == "none":
                ivs.append((lo, hi))
            elif covered == "all":
                op.flags |= flag
            else:
                raise UnsupportedModel("%s: input adjoint region is partially written by later ops; "
                                       "this fan-out pattern is not supported" % op.name)
    # every op output adjoint must have been written by someone (its consumers) before the op runs backward:
    # guaranteed by topological order as long as each tensor has at least one consumer.

    # tensors that no longer exist (fused ReLU outputs, unused nodes) alias their replacement so the
    # table stays valid
    for t, vt in enumerate(vts):
        if vt.buf < 0:
            r = resolve(t)
            if r != t and vts[r].buf >= 0:
                vt.buf, vt.offset = vts[r].buf, vts[r].offset
            else:
                vt.buf, vt.offset, vt.shape = 0, 0, (1, 1, 1)

    conv_layers, seen = [], set()
    for i, op in enumerate(ops):
        if op.kind == OP_CONV and id(op.module) not in seen:
            seen.add(id(op.module))
            conv_layers.append(i)
    return Tape(vts, ops, buf_elems, result, head, b.n_params, tuple(input_shape), b.bn_modules, b.poff, conv_layers)


def _covered(ivs, lo, hi) -> str:
    """'none' if [lo,hi) is disjoint from all intervals, 'all' if fully inside their union, else 'partial'."""
    pts = sorted(ivs)
    overlap = [(max(a, lo), min(c, hi)) for a, c in pts if a < hi and c > lo]
    if not overlap:
        return "none"
    cur = lo
    for a, c in sorted(overlap):
        if a > cur:
            return "partial"
        cur = max(cur, c)
    return "all" if cur >= hi else "partial"
