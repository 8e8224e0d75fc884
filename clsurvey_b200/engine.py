"""Engine: compiles a reference-style nn.Module (features / avgpool / classifier) into a static layer plan and runs
forward, loss head and backward through the C ABI (include/clb.h) on flat fp32 buffers.

PyTorch is used here for device memory, streams and (optionally) CUDA-graph capture only -- no torch op computes
anything on the hot path.  Replaces `model(inputs)`, `criterion(outputs, labels)` and `loss.backward()` of the
reference's train_model loops (src/methods/EWC/train_EWC.py:178-189 and twins) and of its importance passes
(src/methods/EWC/main_EWC.py:138-157, src/methods/MAS/train_MAS.py:508-567).

Layout in HBM:
  theta, grad [, omega, theta_star, momentum, w]  : flat fp32 buffers, one slot per parameter tensor in
      model.parameters() order, slot offsets rounded up to 4 elements (16 B) so every tensor is float4-aligned.
      `p.data` / `p.grad` of the nn.Module are re-pointed at views of these buffers, so state_dict(), pickling and the
      reference's `model.reg_params` dict keep working.
  activations                                    : NCHW fp32, one buffer per layer output (kept for backward),
      uint8 arg-max per max-pool, two ping-pong gradient buffers.
"""
import ctypes
import os
import weakref

import torch
import torch.nn as nn

from . import _capi
from ._capi import call

LOSS_MEAN_CE, LOSS_SUM_NLL, LOSS_SUM_SQ = 0, 1, 2
_ENGINES = weakref.WeakValueDictionary()   # id(parameter) -> Engine (lets the reference-style optimisers find it)


def engine_of(params):
    for p in params:
        e = _ENGINES.get(id(p))
        if e is not None:
            return e
    raise _capi.ClbError("parameters are not bound to a clsurvey_b200 Engine (call Engine(model, ...) first)")


def get_engine(model, input_shape=None, max_batch=200, use_avgpool=True):
    """Engine bound to `model` (created on first use; re-bound after a head swap / unpickling)."""
    eng = getattr(model, "_clb_engine", None)
    if eng is None:
        if input_shape is None:
            raise _capi.ClbError("get_engine: model has no engine yet, input_shape is required")
        return Engine(model, input_shape, max_batch, use_avgpool=use_avgpool)
    if (input_shape is not None and tuple(input_shape) != eng.input_shape) or max_batch > eng.max_batch:
        return Engine(model, input_shape or eng.input_shape, max(max_batch, eng.max_batch), use_avgpool=use_avgpool)
    eng.bind(model)
    return eng


def _ptr(t):
    return 0 if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


class Engine:
    def __init__(self, model, input_shape=(3, 64, 64), max_batch=200, device="cuda", use_avgpool=True):
        _capi.lib()                                   # fail loudly if the CUDA library is missing
        if not torch.cuda.is_available():
            raise _capi.ClbError("clsurvey_b200.Engine needs a CUDA device; there is no CPU fallback")
        self.device = torch.device(device)
        self.model = model
        self.input_shape = tuple(input_shape)
        self.max_batch = int(max_batch)
        self.use_avgpool = use_avgpool
        self.theta = self.grad = None
        self.omega = self.theta_star = self.momentum = self.w = None
        self.momentum_valid = False
        self._grad_alt = None
        self.conv_events = None       # bench.py: list of (start, end) CUDA events around every conv launch
        self.bind(model)

    def _timed(self, fn, *a):
        """Run one C-ABI call; when bench.py asked for it, bracket it with CUDA events on the launching stream."""
        if self.conv_events is None:
            return fn(*a)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn(*a)
        e1.record()
        self.conv_events.append((e0, e1))

    # an Engine never travels with a pickled / deep-copied model (torch.save(model), copy.deepcopy(model))
    def __reduce__(self):
        return (type(None), ())

    def __deepcopy__(self, memo):
        return None

    # ------------------------------------------------------------------ parameters <-> flat buffers
    def bind(self, model=None):
        """(Re)adopt the module's parameters into the flat buffers.  Call again after a head swap
        (utils.replace_last_classifier_layer / main_EWC.py:49-53): values of the fresh head are copied in,
        everything else (incl. omega/theta_star slots of unchanged tensors) keeps its place."""
        model = self.model if model is None else model
        self.model = model
        params = list(model.parameters())
        offs, off = [], 0
        for p in params:
            offs.append(off)
            off += (p.numel() + 3) // 4 * 4
        total = off
        same_layout = (self.theta is not None and total == self.theta.numel()
                       and [tuple(p.shape) for p in params] == self.shapes)
        old_theta = self.theta
        if not same_layout:
            self.theta = torch.zeros(total, dtype=torch.float32, device=self.device)
            self.grad = torch.zeros(total, dtype=torch.float32, device=self.device)
            self._grad_alt = None
            self.omega = self.theta_star = self.momentum = self.w = None
            self.momentum_valid = False
        # parameters that LEAVE the model (a head swapped out by utils.replace_last_classifier_layer, or kept by the caller
        # for evaluation: get_prev_heads, utils.py:235-262) must stop aliasing their old slot of the flat buffers: the slot
        # is about to be overwritten by their successor, and a later re-bind compares data_ptrs to decide about copying
        new_ids = {id(p) for p in params}
        with torch.no_grad():
            for p_old in getattr(self, "params", None) or []:
                if id(p_old) not in new_ids:
                    p_old.data = p_old.data.clone()
                    p_old.grad = None
        for k in [k for k, e in list(_ENGINES.items()) if e is self]:
            del _ENGINES[k]
        object.__setattr__(model, "_clb_engine", self)
        self.params, self.offsets, self.numels = params, offs, [p.numel() for p in params]
        self.shapes = [tuple(p.shape) for p in params]
        self.total, self.n_real = total, sum(self.numels)
        with torch.no_grad():
            for p, o in zip(params, offs):
                view = self.theta[o:o + p.numel()].view(p.shape)
                if p.data.data_ptr() != view.data_ptr():
                    view.copy_(p.data.to(self.device, torch.float32))
                    p.data = view
                p.grad = self.grad[o:o + p.numel()].view(p.shape)
                _ENGINES[id(p)] = self
        del old_theta
        # the plan (ops, activation buffers, workspaces) and captured graphs only depend on the layer structure and on the
        # flat-buffer addresses: a re-bind that changes neither (get_output_def calls get_engine for every batch; a head
        # swapped for one of the same shape) keeps them
        sig = (tuple(self.shapes), tuple(type(m).__name__ for m in model.features.children()),
               tuple(type(m).__name__ for m in model.classifier.children()), type(getattr(model, "avgpool", None)).__name__)
        if same_layout and getattr(self, "_sig", None) == sig and getattr(self, "ops", None):
            for op in self.ops:                       # Dropout ops look at their module's p
                if op["kind"] == "dropout":
                    op["module"] = list(model.classifier.children())[op["cls_idx"]]
            return
        self._sig = sig
        self._graphs = {}
        self._dp_cut_cache = None                     # offsets may have moved (head swap)
        self._dp_buckets_cache = None
        self._compile()

    def view(self, flat, i):
        o = self.offsets[i]
        return flat[o:o + self.numels[i]].view(self.shapes[i])

    def ensure(self, name):
        if getattr(self, name) is None:
            setattr(self, name, torch.zeros(self.total, dtype=torch.float32, device=self.device))
        return getattr(self, name)

    # ------------------------------------------------------------------ plan
    def _compile(self):
        pidx = {id(p): i for i, p in enumerate(self.params)}
        C, H, W = self.input_shape
        ops = []

        def add(**kw):
            ops.append(kw)
            return kw

        feats = list(self.model.features.children())
        i = 0
        while i < len(feats):
            m = feats[i]
            if isinstance(m, nn.Conv2d):
                assert m.groups == 1 and m.dilation == (1, 1) and m.stride[0] == m.stride[1] \
                    and m.padding[0] == m.padding[1], "unsupported conv"
                relu = i + 1 < len(feats) and isinstance(feats[i + 1], nn.ReLU)
                R, S = m.kernel_size
                P = (H + 2 * m.padding[0] - R) // m.stride[0] + 1
                Q = (W + 2 * m.padding[0] - S) // m.stride[0] + 1
                add(kind="conv", C=C, H=H, W=W, K=m.out_channels, R=R, S=S, stride=m.stride[0], pad=m.padding[0],
                    relu=relu, w=pidx[id(m.weight)], b=pidx[id(m.bias)] if m.bias is not None else None,
                    out_shape=(m.out_channels, P, Q))
                C, H, W = m.out_channels, P, Q
                i += 2 if relu else 1
            elif isinstance(m, nn.MaxPool2d):
                k = m.kernel_size if isinstance(m.kernel_size, int) else m.kernel_size[0]
                s = m.stride if isinstance(m.stride, int) else m.stride[0]
                assert (m.padding in (0, (0, 0))) and not m.ceil_mode, "unsupported pool"
                PH, PW = (H - k) // s + 1, (W - k) // s + 1
                add(kind="maxpool", C=C, H=H, W=W, k=k, stride=s, out_shape=(C, PH, PW))
                H, W = PH, PW
                i += 1
            elif isinstance(m, nn.ReLU):
                add(kind="relu", out_shape=(C, H, W))
                i += 1
            else:
                raise NotImplementedError("features layer %r is outside the hot path" % (m,))
        avg = getattr(self.model, "avgpool", None)
        if self.use_avgpool and isinstance(avg, nn.AdaptiveAvgPool2d):
            OH, OW = avg.output_size if isinstance(avg.output_size, tuple) else (avg.output_size,) * 2
            add(kind="avgpool", C=C, H=H, W=W, OH=OH, OW=OW, out_shape=(C, OH, OW))
            H, W = OH, OW
        feat = C * H * W
        cls = list(self.model.classifier.children())
        i = 0
        while i < len(cls):
            m = cls[i]
            if isinstance(m, nn.Linear):
                assert m.in_features == feat, "classifier input %d != features output %d" % (m.in_features, feat)
                relu = i + 1 < len(cls) and isinstance(cls[i + 1], nn.ReLU)
                add(kind="linear", inf=feat, outf=m.out_features, relu=relu, w=pidx[id(m.weight)],
                    b=pidx[id(m.bias)] if m.bias is not None else None, out_shape=(m.out_features,), cls_idx=i)
                feat = m.out_features
                i += 2 if relu else 1
            elif isinstance(m, nn.Dropout):
                add(kind="dropout", p=m.p, feat=feat, out_shape=(feat,), cls_idx=i, module=m)
                i += 1
            elif isinstance(m, nn.ReLU):
                add(kind="relu", out_shape=(feat,))
                i += 1
            else:
                raise NotImplementedError("classifier layer %r is outside the hot path" % (m,))
        self.ops = ops
        self.n_outputs = feat
        B = self.max_batch
        self._plan_planes()
        max_act = 1
        for op in ops:
            n = 1
            for d in op["out_shape"]:
                n *= d
            op["out_numel"] = n
            lay_out = op.get("lay_out", "nchw")
            if op.get("fused_first"):
                pass                                              # its fp32 output (210 MB for VGG-11 / batch 200) never exists
            elif lay_out == "nchw":
                op["out"] = torch.empty(B * n, dtype=torch.float32, device=self.device)
                max_act = max(max_act, n)
            elif not op.get("transient"):                 # bf16 hi / lo planes, NHWC (csrc/clb_planes_conv.cu)
                op["out_pl"] = torch.empty(2, B * n, dtype=torch.int16, device=self.device)
                if op.get("also_f32"):
                    op["out"] = torch.empty(B * n, dtype=torch.float32, device=self.device)
            if op["kind"] == "maxpool":
                op["argmax"] = torch.empty(B * n, dtype=torch.uint8, device=self.device)
        in_numel = self.input_shape[0] * self.input_shape[1] * self.input_shape[2]
        max_act = max(max_act, in_numel)
        self.dbuf = [torch.empty(B * max_act, dtype=torch.float32, device=self.device) for _ in range(2)]
        self.dlogits = torch.empty(B * self.n_outputs, dtype=torch.float32, device=self.device)
        ws_bytes, wt_elems = 16, 4
        pl_max = max([op["out_numel"] for op in ops if op.get("lay_out") in ("planes", "planes_flat")] +
                     [op["inf"] for op in ops if op["kind"] == "linear" and op.get("planes")] +
                     [op["C"] * op["H"] * op["W"] for op in ops if op["kind"] == "conv" and op.get("planes")] + [0])
        self.pl_scratch = self.dpl = None
        if pl_max:
            self.pl_scratch = torch.empty(2, B * pl_max, dtype=torch.int16, device=self.device)      # pre-pool conv outputs
            self.dpl = [torch.empty(2, B * pl_max, dtype=torch.int16, device=self.device) for _ in range(2)]
        for op in ops:
            if op["kind"] == "conv" and op.get("planes"):
                K, C = op["K"], op["C"]
                op["wf"] = torch.empty(2, K * 9 * C, dtype=torch.int16, device=self.device)   # [K][tap][C] hi / lo
                op["wt"] = torch.empty(2, K * 9 * C, dtype=torch.int16, device=self.device)   # [C][8 - tap][K] hi / lo
                ws_bytes = max(ws_bytes, _capi.lib().clb_planes_conv_wgrad_ws(B, op["H"], op["W"], C, K))
                continue
            if op["kind"] == "conv":
                ws_bytes = max(ws_bytes, _capi.lib().clb_conv2d_wgrad_ws(B, op["C"], op["H"], op["W"], op["K"], op["R"],
                                                                          op["S"], op["stride"], op["pad"]))
                # re-ordered weights for the tensor-core path: hi + lo plane, each max(K*C*R*S, 32*K, 32*C) floats
                wt_elems = max(wt_elems, 2 * (max(op["K"] * op["C"] * op["R"] * op["S"], op["K"] * 32, op["C"] * 32) + 4))
        if any(op.get("fused_first") for op in ops):
            ws_bytes = max(ws_bytes, _capi.lib().clb_planes_conv1_ws())
        for op in ops:
            if op["kind"] == "linear" and op.get("planes"):
                op["wf"] = torch.empty(2, op["outf"] * op["inf"], dtype=torch.int16, device=self.device)    # [out][in] hi / lo
                op["wt"] = torch.empty(2, op["outf"] * op["inf"], dtype=torch.int16, device=self.device)    # [in][out] hi / lo
                ws_bytes = max(ws_bytes, _capi.lib().clb_planes_linear_wgrad_ws(B, op["inf"], op["outf"]))
            elif op["kind"] == "linear":
                ws_bytes = max(ws_bytes, _capi.lib().clb_linear_ws(B, op["inf"], op["outf"]))
        # importance passes (EWC Fisher / MAS omega) fuse `omega (+)= f(dW)` into the split-K reduction of the planes convs
        # (clb_planes_conv_wgrad, imp_mode); the flat-buffer ranges NOT covered by those weights get the streaming kernel
        covered = sorted((self.offsets[op["w"]], self.offsets[op["w"]] + (self.numels[op["w"]] + 3) // 4 * 4)
                         for op in ops if op["kind"] in ("conv", "linear") and op.get("planes"))
        self._imp_rest, pos = [], 0
        for a, b in covered:
            if a > pos:
                self._imp_rest.append((pos, a - pos))
            pos = b
        if pos < self.total:
            self._imp_rest.append((pos, self.total - pos))
        self._wbatch = None
        pconvs = [op for op in ops if op["kind"] in ("conv", "linear") and op.get("planes")]
        if pconvs:
            n = len(pconvs)
            P, I = ctypes.c_void_p * n, ctypes.c_int * n
            lin = lambda op: op["kind"] == "linear"
            self._wbatch = (n, P(*[_ptr(self.view(self.theta, op["w"])) for op in pconvs]),
                            P(*[_ptr(op["wf"][0]) for op in pconvs]), P(*[_ptr(op["wf"][1]) for op in pconvs]),
                            P(*[_ptr(op["wt"][0]) for op in pconvs]), P(*[_ptr(op["wt"][1]) for op in pconvs]),
                            I(*[op["outf"] if lin(op) else op["K"] for op in pconvs]),
                            I(*[op["inf"] if lin(op) else op["C"] for op in pconvs]), I(*[1 if lin(op) else 9 for op in pconvs]))
        self.ws = torch.empty((ws_bytes + 3) // 4, dtype=torch.float32, device=self.device)
        self.wt_ws = torch.empty(wt_elems, dtype=torch.float32, device=self.device)
        self.loss_dev = torch.zeros(1, dtype=torch.float32, device=self.device)
        self.correct_dev = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.n_launch = 0

    def _plan_planes(self):
        """Mark the convs that run on the planes kernels (3x3/1/1 + ReLU, C % 64 == K % 64 == 0, power-of-two maps) and the
        layout ('nchw' fp32 | 'planes' bf16 hi/lo NHWC) of every features op's output.  A planes conv reads planes: from a
        planes conv directly or through a 2x2/2 max-pool (which converts from an fp32 NCHW producer if need be); it hands its
        output to a planes conv or to a 2x2/2 max-pool (which converts back to fp32 NCHW for any other consumer)."""
        ops = self.ops
        nfeat = next((i for i, op in enumerate(ops) if op["kind"] in ("avgpool", "linear", "dropout")), len(ops))
        lib = _capi.lib()
        self._plan_mode = lib.clb_get_matmul_mode()
        if os.environ.get("CLB_PLANES", "1") == "0" or self._plan_mode != 3:
            return

        def pool22(op):
            return op["kind"] == "maxpool" and op["k"] == 2 and op["stride"] == 2 and op["H"] % 2 == 0 and op["W"] % 2 == 0

        for i in range(nfeat):
            op = ops[i]
            op["planes"] = bool(op["kind"] == "conv" and op["relu"] and op["b"] is not None and lib.clb_planes_conv_supported(
                op["C"], op["H"], op["W"], op["K"], op["R"], op["S"], op["stride"], op["pad"]))
        changed = True
        while changed:
            changed = False
            for i in range(nfeat):
                op = ops[i]
                if not op.get("planes"):
                    continue
                prev = ops[i - 1] if i > 0 else None
                pprev = ops[i - 2] if i > 1 else None
                in_ok = prev is not None and (prev.get("planes") or (
                    pool22(prev) and pprev is not None and pprev["kind"] == "conv" and pprev["relu"] and
                    (pprev.get("planes") or (pprev["K"] % 64 == 0 and prev["W"] <= 128))))
                nxt = ops[i + 1] if i + 1 < nfeat else None
                out_ok = nxt is not None and (nxt.get("planes") or pool22(nxt))
                if not (in_ok and out_ok):
                    op["planes"] = False
                    changed = True
        for i in range(nfeat):
            op = ops[i]
            nxt = ops[i + 1] if i + 1 < nfeat else None
            if op["kind"] == "conv" and op.get("planes"):
                op["lay_out"] = "planes"
                op["transient"] = nxt is not None and nxt["kind"] == "maxpool"     # only the pool reads it
            elif op["kind"] == "maxpool":
                op["lay_in"] = "planes" if (i > 0 and ops[i - 1].get("planes")) else "nchw"
                op["lay_out"] = "planes" if (nxt is not None and nxt.get("planes")) else "nchw"
                prev = ops[i - 1] if i > 0 else None
                # first layer (C = 3 -> 64) + ReLU + this pool as ONE kernel each way (csrc/clb_planes_first.cu)
                if (i == 1 and op["lay_in"] == "nchw" and op["lay_out"] == "planes" and prev["kind"] == "conv" and prev["relu"]
                        and prev["b"] is not None and os.environ.get("CLB_PLANES_FIRST", "1") != "0"
                        and lib.clb_planes_conv1_supported(prev["C"], prev["H"], prev["W"], prev["K"], prev["R"], prev["S"],
                                                           prev["stride"], prev["pad"])):
                    prev["fused_first"] = op["fused_prev"] = True
        # classifier: nn.Linear + ReLU layers on the planes kernels (a 1x1 "conv" over a 1x1 map, rows = samples) when the
        # features end in a planes pool, no Dropout sits in the classifier (VGG) and in % 64 == out % 64 == 0.  The pool
        # then emits planes in the flatten order; the last planes Linear also leaves an fp32 copy for the head.
        cls_ops = ops[nfeat:]
        last_feat = ops[nfeat - 1] if nfeat > 0 else None
        if (os.environ.get("CLB_PLANES_LINEAR", "1") != "0" and last_feat is not None and last_feat["kind"] == "maxpool"
                and last_feat.get("lay_in") == "planes" and cls_ops and all(o["kind"] == "linear" for o in cls_ops)):
            chain = True
            for j, o in enumerate(cls_ops):
                o["planes"] = bool(chain and o["relu"] and o["b"] is not None and j + 1 < len(cls_ops)
                                   and lib.clb_planes_linear_supported(o["inf"], o["outf"]))
                chain = o["planes"]
            if cls_ops[0]["planes"]:
                last_feat["lay_out"] = "planes_flat"
            for j, o in enumerate(cls_ops):
                if o["planes"]:
                    o["lay_out"] = "planes"
                    o["also_f32"] = not cls_ops[j + 1].get("planes")      # the legacy layer behind reads fp32

    # ------------------------------------------------------------------ forward
    def forward(self, x, train=False, masks=None):
        """x: [n, C, H, W] fp32 CUDA contiguous.  Returns logits [n, n_outputs] (a view of an internal buffer).
        masks: {classifier child index: pre-scaled mask tensor [n, feat] or [feat]} for active Dropout layers."""
        n = x.shape[0]
        assert n <= self.max_batch and x.is_cuda and x.dtype == torch.float32
        x = x.contiguous()
        assert tuple(x.shape[1:]) == self.input_shape, (x.shape, self.input_shape)
        s = _stream()
        if _capi.lib().clb_get_matmul_mode() != self._plan_mode:     # the planes / legacy split follows the matmul mode
            self._graphs = {}
            self._compile()
        cur = x
        self._n = n
        self._train = train
        self._masks = {}
        wjoin = False
        if self._wbatch is not None:      # weights of all planes convs -> bf16 hi/lo planes (fwd + dgrad layouts), one launch
            wjoin = bool(self.ops and self.ops[0].get("fused_first"))
            self._side_call(wjoin, "clb_planes_weights_batch", *self._wbatch, s)
            self.n_launch += 1
        for op in self.ops:
            op["inp"] = cur
            k = op["kind"]
            if k == "conv" and op.get("fused_first"):
                pool = self.ops[1]
                out = pool["out_pl"]
                self._timed(call, "clb_planes_conv1_pool_fwd", _ptr(cur), _ptr(self.view(self.theta, op["w"])),
                            _ptr(self.view(self.theta, op["b"])), _ptr(out[0]), _ptr(out[1]), _ptr(pool["argmax"]), n, op["C"],
                            op["H"], op["W"], op["K"], s)
                if wjoin:                 # the fp32 first layer ran next to the weight conversion on the side stream
                    call("clb_planes_join", s)
                    wjoin = False
                cur = out
            elif k == "maxpool" and op.get("fused_prev"):
                continue                                              # done by the conv kernel in front
            elif k == "conv" and op.get("planes"):
                # cur = (hi, lo) planes [n][H][W][C]
                out = self.pl_scratch if op["transient"] else op["out_pl"]
                self._timed(call, "clb_planes_conv_fwd", _ptr(cur[0]), _ptr(cur[1]), _ptr(op["wf"][0]), _ptr(op["wf"][1]),
                            _ptr(self.view(self.theta, op["b"])), _ptr(out[0]), _ptr(out[1]), n, op["H"], op["W"], op["C"],
                            op["K"], 1, s)
                cur = out
                self.n_launch += 1
            elif k == "maxpool" and (op.get("lay_in") == "planes" or op.get("lay_out") == "planes"):
                if op["lay_in"] == "nchw":                            # fp32 NCHW (legacy conv + ReLU) -> planes
                    out = op["out_pl"]
                    call("clb_planes_pool_fwd_nchw", _ptr(cur), _ptr(out[0]), _ptr(out[1]), _ptr(op["argmax"]), n, op["C"],
                         op["H"], op["W"], s)
                    cur = out
                elif op["lay_out"] == "planes":
                    out = op["out_pl"]
                    call("clb_planes_pool_fwd", _ptr(cur[0]), _ptr(cur[1]), _ptr(out[0]), _ptr(out[1]), 0, _ptr(op["argmax"]),
                         n, op["H"], op["W"], op["C"], s)
                    cur = out
                elif op["lay_out"] == "planes_flat":                  # planes in the classifier's flatten order
                    out = op["out_pl"]
                    call("clb_planes_pool_fwd_flat", _ptr(cur[0]), _ptr(cur[1]), _ptr(out[0]), _ptr(out[1]), _ptr(op["argmax"]),
                         n, op["H"], op["W"], op["C"], s)
                    cur = out
                else:                                                 # planes -> fp32 NCHW (classifier / legacy consumer)
                    call("clb_planes_pool_fwd", _ptr(cur[0]), _ptr(cur[1]), 0, 0, _ptr(op["out"]), _ptr(op["argmax"]), n,
                         op["H"], op["W"], op["C"], s)
                    cur = op["out"]
            elif k == "conv":
                self._timed(call, "clb_conv2d_fwd", _ptr(cur), _ptr(self.view(self.theta, op["w"])),
                     _ptr(self.view(self.theta, op["b"])) if op["b"] is not None else 0, _ptr(op["out"]),
                     _ptr(self.wt_ws), n, op["C"], op["H"], op["W"], op["K"], op["R"], op["S"], op["stride"], op["pad"],
                     int(op["relu"]), s)
                cur = op["out"]
            elif k == "maxpool":
                call("clb_maxpool_fwd", _ptr(cur), _ptr(op["out"]), _ptr(op["argmax"]), n, op["C"], op["H"], op["W"],
                     op["k"], op["stride"], s)
                cur = op["out"]
            elif k == "avgpool":
                call("clb_adaptive_avgpool_fwd", _ptr(cur), _ptr(op["out"]), n, op["C"], op["H"], op["W"], op["OH"],
                     op["OW"], s)
                cur = op["out"]
            elif k == "linear" and op.get("planes"):
                out = op["out_pl"]
                self._timed(call, "clb_planes_linear_fwd", _ptr(cur[0]), _ptr(cur[1]), _ptr(op["wf"][0]), _ptr(op["wf"][1]),
                            _ptr(self.view(self.theta, op["b"])), _ptr(out[0]), _ptr(out[1]), n, op["inf"], op["outf"], 1, s)
                cur = out
                if op["also_f32"]:                                    # the head (not a multiple of 64 wide) runs on fp32
                    call("clb_planes_to_f32", _ptr(out[0]), _ptr(out[1]), _ptr(op["out"]), n * op["outf"], s)
                    cur = op["out"]
            elif k == "linear":
                call("clb_linear_fwd", _ptr(cur), _ptr(self.view(self.theta, op["w"])),
                     _ptr(self.view(self.theta, op["b"])) if op["b"] is not None else 0, _ptr(op["out"]),
                     _ptr(self.ws), self.ws.numel() * 4, n, op["inf"], op["outf"], int(op["relu"]), s)
                cur = op["out"]
            elif k == "dropout":
                if not train:
                    continue
                mask = None if masks is None else masks.get(op["cls_idx"])
                if mask is None:
                    mask = self.draw_dropout_mask(op, n)
                mask = mask.to(self.device, torch.float32).contiguous()
                self._masks[op["cls_idx"]] = mask
                rows = 1 if mask.dim() == 1 else mask.shape[0]
                call("clb_mask_mul", _ptr(cur), _ptr(mask), _ptr(op["out"]), n, op["feat"], rows, s)
                cur = op["out"]
            elif k == "relu":
                raise NotImplementedError("stand-alone ReLU (not following Conv2d/Linear)")
            self.n_launch += 1
        self.logits = cur[:n * self.n_outputs].view(n, self.n_outputs)
        return self.logits

    def draw_dropout_mask(self, op, n):
        """Element mask from the HOST torch generator: F.dropout(x, p) == x * bernoulli(1-p)/(1-p) under the same seed
        (SURVEY.md hard part 4) -- the engine receives masks, it does not regenerate them."""
        keep = 1.0 - op["p"]
        if getattr(self, "dropout_rng", "host") == "device":      # throughput mode: no per-step H2D of the masks
            return torch.empty(n, op["feat"], device=self.device).bernoulli_(keep).div_(keep)
        return torch.empty(n, op["feat"]).bernoulli_(keep).div_(keep)

    # ------------------------------------------------------------------ loss head
    def loss_head(self, labels, mode=LOSS_MEAN_CE, denom=None, col_off=0, ncols=None, want_grad=True):
        """Fused softmax / loss / #correct / dlogits on the last forward's logits.  Device scalars only (no sync)."""
        n = self._n
        ncols = self.n_outputs - col_off if ncols is None else ncols
        denom = float(n if denom is None else denom)
        call("clb_memset_zero", _ptr(self.loss_dev), 4, _stream())
        call("clb_memset_zero", _ptr(self.correct_dev), 4, _stream())
        if labels is not None:
            labels = labels.to(self.device, torch.int64).contiguous()
        self._labels = labels
        call("clb_softmax_loss", _ptr(self.logits), self.n_outputs, col_off, ncols, _ptr(labels), n, mode, denom,
             _ptr(self.loss_dev), _ptr(self.correct_dev), _ptr(self.dlogits) if want_grad else 0, _stream())
        self.n_launch += 1

    # ------------------------------------------------------------------ backward
    def _dp_cut(self):
        """Data parallelism: (op index, flat offset) at which the gradient buffer is split into a small head (the first
        layers, <= 10 % of the parameters) and the tail.  backward() produces the tail first, so its all-reduce can run
        on a side stream underneath the backward pass of the head layers.  (0, 0) = no split."""
        if getattr(self, "_dp_cut_cache", None) is None:
            cut = (0, 0)
            for i, op in enumerate(self.ops):
                if op["kind"] in ("conv", "linear"):
                    o = self.offsets[op["w"]]
                    if 0 < o <= 0.1 * self.total and o % 4 == 0:
                        cut = (i, o)
            self._dp_cut_cache = cut
        return self._dp_cut_cache

    def _dp_buckets(self):
        """The tail behind _dp_cut() in up to three buckets, [(op index, flat offset, end offset)] in the order backward()
        completes them (back to front): a bucket is closed at a layer boundary once it holds >= 20 % of the parameters.
        Each bucket's all-reduce starts as soon as its first layer's gradient is final (VGG-11: [conv8, fc1, fc2, head],
        [conv7], [conv5, conv6] -- 14.7 / 9.4 / 14.2 MB -- then the head conv1-4, 3.8 MB, after the last layer)."""
        if getattr(self, "_dp_buckets_cache", None) is None:
            cut_op, cut_off = self._dp_cut()
            buckets = []
            if cut_off:
                end, acc = self.total, 0
                layers = [(i, self.offsets[op["w"]]) for i, op in enumerate(self.ops) if op["kind"] in ("conv", "linear")]
                for i, o in reversed(layers):
                    if o < cut_off:
                        break
                    acc = end - o
                    if (o == cut_off) or (acc >= 0.2 * self.total and o % 4 == 0 and len(buckets) < 2):
                        buckets.append((i, o, end))
                        end = o
            self._dp_buckets_cache = buckets
        return self._dp_buckets_cache

    def backward(self, accumulate=False, dp_overlap=False, importance=None):
        """dlogits -> parameter gradients (flat self.grad).  importance = (1, data_len) [EWC: omega += g*g/data_len,
        main_EWC.py:151-156] or (2, prev_size, curr_size) [MAS: omega = (omega*prev + |g|)/curr, train_MAS.py:163-177]
        applies that update to self.omega as part of this backward pass: inside the split-K reduction of the planes convs
        (which holds every final dW element in a register anyway), one streaming launch per remaining range.  accumulate=True adds to the existing gradient
        (GEM memory mini-batches, gem.py:239-256, never zero the grads in between).  dp_overlap=True (data-parallel
        training steps only): the caller promises to call dist.allreduce_grads(self) next; the tail of the flat
        gradient is then all-reduced on a side stream as soon as its last layer is done."""
        n, s = self._n, _stream()
        self._dp_pending = None
        imp_mode, imp_a, imp_b = 0, 0.0, 0.0
        if importance is not None:
            assert not accumulate and self.omega is not None
            imp_mode, imp_a = int(importance[0]), float(importance[1])
            imp_b = float(importance[2]) if len(importance) > 2 else 0.0
        buckets = []
        if dp_overlap and not accumulate:
            from . import dist as _dist
            if _dist.is_distributed() and self.grad.is_cuda:
                buckets = list(self._dp_buckets())
        gdst = self.grad
        if accumulate:
            if self._grad_alt is None:
                self._grad_alt = torch.zeros_like(self.grad)
            gdst = self._grad_alt
        d, other, pl_other = self.dlogits, 0, 0
        first_param_op = next(i for i, op in enumerate(self.ops) if op["kind"] in ("conv", "linear"))
        relu_done = set()
        for i in range(len(self.ops) - 1, -1, -1):
            op = self.ops[i]
            k = op["kind"]
            while buckets and i == buckets[0][0] - 1:                 # every gradient at offset >= buckets[0][1] is final
                from . import dist as _dist
                _, b_off, b_end = buckets.pop(0)
                self._dp_pending = _dist.start_tail_allreduce(self, b_off, b_end)
            if k == "linear" and op.get("planes"):
                x_pl = op["inp"]                                      # planes of this layer's input
                if not isinstance(d, (list, tuple)) and d.dim() == 1:  # fp32 gradient from the head: ReLU backward + planes
                    nxt = self.dpl[pl_other]
                    call("clb_planes_from_f32", _ptr(d), _ptr(op["out_pl"][0]), _ptr(nxt[0]), _ptr(nxt[1]), n * op["outf"], s)
                    d, pl_other = nxt, pl_other ^ 1
                self._timed(self._side_call, True, "clb_planes_linear_wgrad", _ptr(x_pl[0]), _ptr(x_pl[1]), _ptr(d[0]), _ptr(d[1]),
                            _ptr(self.view(gdst, op["w"])), _ptr(self.view(gdst, op["b"])), _ptr(self.ws), self.ws.numel() * 4,
                            n, op["inf"], op["outf"], imp_mode, _ptr(self.view(self.omega, op["w"])) if imp_mode else 0,
                            imp_a, imp_b, s)
                prev = self.ops[i - 1]
                nxt = self.dpl[pl_other]
                mask = _ptr(x_pl[0]) if (prev["kind"] == "linear" and prev.get("planes")) else 0     # ReLU of the Linear in front
                self._timed(call, "clb_planes_linear_dgrad", _ptr(d[0]), _ptr(d[1]), _ptr(op["wt"][0]), _ptr(op["wt"][1]), mask,
                            _ptr(nxt[0]), _ptr(nxt[1]), n, op["inf"], op["outf"], s)
                call("clb_planes_join", s)                           # split-K reduce + bias grad of this layer (side stream)
                d, pl_other = nxt, pl_other ^ 1
                self.n_launch += 5
            elif k == "linear":
                if op["relu"]:
                    call("clb_relu_bwd", _ptr(d), _ptr(op["out"]), _ptr(d), n * op["outf"], s)
                call("clb_linear_wgrad", _ptr(op["inp"]), _ptr(d), _ptr(self.view(gdst, op["w"])),
                     _ptr(self.view(gdst, op["b"])) if op["b"] is not None else 0, _ptr(self.ws), self.ws.numel() * 4,
                     n, op["inf"], op["outf"], s)
                if i != first_param_op:
                    nxt = self.dbuf[other]
                    call("clb_linear_dgrad", _ptr(d), _ptr(self.view(self.theta, op["w"])), _ptr(nxt), _ptr(self.ws),
                         self.ws.numel() * 4, n, op["inf"], op["outf"], s)
                    d, other = nxt, other ^ 1
                self.n_launch += 3
            elif k == "dropout":
                if self._train and op["cls_idx"] in self._masks:
                    mask = self._masks[op["cls_idx"]]
                    rows = 1 if mask.dim() == 1 else mask.shape[0]
                    call("clb_mask_mul", _ptr(d), _ptr(mask), _ptr(d), n, op["feat"], rows, s)
            elif k == "avgpool":
                nxt = self.dbuf[other]
                call("clb_adaptive_avgpool_bwd", _ptr(d), _ptr(nxt), n, op["C"], op["H"], op["W"], op["OH"], op["OW"], s)
                d, other = nxt, other ^ 1
            elif k == "maxpool" and op.get("fused_prev"):
                continue                                              # d stays the planes gradient of the pooled output
            elif k == "conv" and op.get("fused_first"):
                pool = self.ops[1]
                self._timed(call, "clb_planes_conv1_pool_bwd", _ptr(op["inp"]), _ptr(d[0]), _ptr(d[1]), _ptr(pool["out_pl"][0]),
                            _ptr(pool["argmax"]), _ptr(self.view(gdst, op["w"])), _ptr(self.view(gdst, op["b"])), _ptr(self.ws),
                            self.ws.numel() * 4, n, op["C"], op["H"], op["W"], op["K"], s)
                self.n_launch += 2
            elif k == "maxpool" and (op.get("lay_in") == "planes" or op.get("lay_out") == "planes"):
                # max-pool backward fused with the ReLU backward of the conv in front (pooled > 0 <=> selected input > 0)
                if op["lay_in"] == "nchw":                            # d planes -> fp32 NCHW dY of a legacy conv
                    nxt = self.dbuf[other]
                    call("clb_planes_pool_bwd_nchw", _ptr(d[0]), _ptr(d[1]), _ptr(op["out_pl"][0]), _ptr(op["argmax"]),
                         _ptr(nxt), n, op["C"], op["H"], op["W"], s)
                    d, other = nxt, other ^ 1
                elif op["lay_out"] == "planes_flat":                  # gradient / pooled activation as planes in flatten order
                    nxt = self.dpl[pl_other]
                    call("clb_planes_pool_bwd_flat", _ptr(d[0]), _ptr(d[1]), _ptr(op["out_pl"][0]), _ptr(op["argmax"]),
                         _ptr(nxt[0]), _ptr(nxt[1]), n, op["H"], op["W"], op["C"], s)
                    d, pl_other = nxt, pl_other ^ 1
                else:
                    nxt = self.dpl[pl_other]
                    if op["lay_out"] == "nchw":                       # fp32 NCHW d (classifier side) -> planes
                        call("clb_planes_pool_bwd", 0, 0, _ptr(d), 0, _ptr(op["out"]), _ptr(op["argmax"]), _ptr(nxt[0]),
                             _ptr(nxt[1]), n, op["H"], op["W"], op["C"], s)
                    else:
                        call("clb_planes_pool_bwd", _ptr(d[0]), _ptr(d[1]), 0, _ptr(op["out_pl"][0]), 0, _ptr(op["argmax"]),
                             _ptr(nxt[0]), _ptr(nxt[1]), n, op["H"], op["W"], op["C"], s)
                    d, pl_other = nxt, pl_other ^ 1
                relu_done.add(i - 1)
                self.n_launch += 1
            elif k == "conv" and op.get("planes"):
                x_pl = op["inp"]                                      # planes of this conv's (post-ReLU) input
                self._timed(self._side_call, True, "clb_planes_conv_wgrad", _ptr(x_pl[0]), _ptr(x_pl[1]), _ptr(d[0]), _ptr(d[1]),
                            _ptr(self.view(gdst, op["w"])), _ptr(self.view(gdst, op["b"])), _ptr(self.ws), self.ws.numel() * 4,
                            n, op["H"], op["W"], op["C"], op["K"], imp_mode,
                            _ptr(self.view(self.omega, op["w"])) if imp_mode else 0, imp_a, imp_b, s)
                prev = self.ops[i - 1]
                nxt = self.dpl[pl_other]
                mask = 0
                if prev.get("planes"):                                # ReLU backward of the conv in front, fused
                    mask = _ptr(x_pl[0])
                    relu_done.add(i - 1)
                self._timed(call, "clb_planes_conv_dgrad", _ptr(d[0]), _ptr(d[1]), _ptr(op["wt"][0]), _ptr(op["wt"][1]), mask,
                            _ptr(nxt[0]), _ptr(nxt[1]), n, op["H"], op["W"], op["C"], op["K"], s)
                call("clb_planes_join", s)                           # split-K reduce + bias grad of this layer (side stream)
                d, pl_other = nxt, pl_other ^ 1
                self.n_launch += 5
            elif k == "maxpool":
                prev = self.ops[i - 1] if i > 0 else None
                fuse = prev is not None and prev["kind"] == "conv" and prev["relu"]
                nxt = self.dbuf[other]
                call("clb_maxpool_bwd", _ptr(d), _ptr(op["argmax"]), _ptr(prev["out"]) if fuse else 0, _ptr(nxt), n,
                     op["C"], op["H"], op["W"], op["k"], op["stride"], s)
                if fuse:
                    relu_done.add(i - 1)
                d, other = nxt, other ^ 1
                self.n_launch += 1
            elif k == "conv":
                if op["relu"] and i not in relu_done:
                    call("clb_relu_bwd", _ptr(d), _ptr(op["out"]), _ptr(d), n * op["out_numel"], s)
                self._timed(call, "clb_conv2d_wgrad", _ptr(op["inp"]), _ptr(d), _ptr(self.view(gdst, op["w"])),
                     _ptr(self.view(gdst, op["b"])) if op["b"] is not None else 0, _ptr(self.ws), self.ws.numel() * 4,
                     n, op["C"], op["H"], op["W"], op["K"], op["R"], op["S"], op["stride"], op["pad"], s)
                if i != first_param_op:
                    nxt = self.dbuf[other]
                    self._timed(call, "clb_conv2d_dgrad", _ptr(d), _ptr(self.view(self.theta, op["w"])), _ptr(nxt),
                         _ptr(self.wt_ws), n, op["C"], op["H"], op["W"], op["K"], op["R"], op["S"], op["stride"],
                         op["pad"], s)
                    d, other = nxt, other ^ 1
                self.n_launch += 4
        if accumulate:
            call("clb_axpby", _ptr(self.grad), _ptr(self.grad), _ptr(gdst), 1.0, self.total, s)
        if imp_mode:
            for off, cnt in self._imp_rest:
                if imp_mode == 1:
                    call("clb_fisher_accum", _ptr(self.omega[off:]), _ptr(self.grad[off:]), imp_a, cnt, s)
                else:
                    call("clb_mas_accum", _ptr(self.omega[off:]), _ptr(self.grad[off:]), imp_a, imp_b, cnt, s)
                self.n_launch += 1

    def _side_call(self, defer, name, *args):
        """A planes call whose memory-bound tail runs on the library's side stream; defer=True leaves the join to the
        caller's clb_planes_join (include/clb.h), which must follow before anything reads the call's results.  Not while
        bench.py times the calls one by one: each call then carries its own tail."""
        if not defer or self.conv_events is not None:
            return call(name, *args)
        call("clb_planes_defer_join", 1)
        try:
            return call(name, *args)
        finally:
            call("clb_planes_defer_join", 0)

    def zero_grad(self):
        call("clb_memset_zero", _ptr(self.grad), self.grad.numel() * 4, _stream())

    def backward_skip(self, dp_overlap=False):
        """A rank whose shard of the mini-batch is empty contributes a zero gradient -- through the same sequence of
        collectives as the ranks that ran backward(dp_overlap=...)."""
        self.zero_grad()
        self._dp_pending = None
        if dp_overlap:
            from . import dist as _dist
            if _dist.is_distributed() and self.grad.is_cuda:
                for _, b_off, b_end in self._dp_buckets():
                    self._dp_pending = _dist.start_tail_allreduce(self, b_off, b_end)

    # ------------------------------------------------------------------ composite steps
    def fwd_loss_bwd(self, x, y, mode=LOSS_MEAN_CE, denom=None, train=True, masks=None, col_off=0, ncols=None,
                     accumulate=False, dp_overlap=False, importance=None):
        self.forward(x, train=train, masks=masks)
        self.loss_head(y, mode, denom, col_off, ncols, want_grad=True)
        self.backward(accumulate=accumulate, dp_overlap=dp_overlap, importance=importance)

    def fwd_loss(self, x, y, mode=LOSS_MEAN_CE, col_off=0, ncols=None):
        self.forward(x, train=False)
        self.loss_head(y, mode, None, col_off, ncols, want_grad=False)

    # ------------------------------------------------------------------ CUDA-graph replay of a whole step
    def graphed(self, key, n, body):
        """Capture `body(x_static, y_static)` (a sequence of C-ABI launches on the current stream) once per `key` and
        return a callable(x, y) that copies the batch into the static buffers and replays the graph.  Capturing does
        not execute anything, so the caller must have run `body` eagerly at least once before (lazy kernel attributes,
        momentum buffers).  Turns the ~75 launches + ctypes calls of a VGG-11 step into one graph launch."""
        if not hasattr(self, "_graphs"):
            self._graphs = {}
        ent = self._graphs.get(key)
        if ent is None:
            xs = torch.empty((n,) + self.input_shape, dtype=torch.float32, device=self.device)
            ys = torch.zeros(n, dtype=torch.int64, device=self.device)
            g = torch.cuda.CUDAGraph()
            torch.cuda.synchronize()
            # thread_local: the background checkpoint writer (methods/trainers.py) may issue D2H copies from its own thread
            # while this thread captures
            with torch.cuda.graph(g, capture_error_mode="thread_local"):
                body(xs, ys)
            ent = (g, xs, ys)
            self._graphs[key] = ent
        g, xs, ys = ent

        def run(x, y):
            xs.copy_(x, non_blocking=True)
            ys.copy_(y, non_blocking=True)
            g.replay()
        return run

    def drop_graphs(self):
        self._graphs = {}

    def read_loss_correct(self):
        """One device->host read of (loss, #correct) -- the reference does this every batch (train_EWC.py:196-197)."""
        return float(self.loss_dev.item()), int(self.correct_dev.item())
