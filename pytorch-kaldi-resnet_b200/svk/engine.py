"""Host-side plan executor for the speaker-embedding network.

`SpeakerNetEngine` owns (a) the flat fp32 parameter / gradient buffers the module's nn.Parameters are views of,
(b) the packed conv weights, (c) per-shape activation workspaces, and strings the libsvk kernels into the
forward / backward passes of the reference network:

    reference                                   here
    ---------                                   ----
    ResNet.forward          model.py:246-269    _trunk_train / _trunk_eval
    BasicBlock.forward      model.py:48-64      _block_fwd_train / _block_bwd / eval loop in _trunk_eval
    StatsPooling.forward    model.py:441-455    svk_statspool_fwd/bwd
    fc1                     model.py:384        svk_sgemm (+bias)
    AAMLayer.forward        model.py:483-501    svk_l2norm_rows_* + svk_sgemm + svk_aam_margin_*
    loss.backward()         train_resnet.py:327 _backward_train (one autograd.Function around the whole network)

Activations are NHWC, bf16 (product) or fp32 (validation mode); torch is used for memory, streams and autograd
plumbing only — every arithmetic step is a libsvk kernel.
"""
import math
import os

import torch

from . import lib
from .lib import call, make_conv_desc

_EPS = 1e-5
_MOM = 0.1


def _ptr(t):
    return 0 if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


class _ConvRec(object):
    """One conv layer: parameter holder + packed-weight views + geometry."""
    __slots__ = ("mod", "cin", "cout", "R", "stride", "w_fwd", "w_dgrad", "dw_packed", "name")

    def __init__(self, mod, name):
        self.mod = mod
        self.name = name
        self.cout, self.cin, self.R, _ = mod.weight.shape
        self.stride = mod.stride[0]
        self.w_fwd = self.w_dgrad = self.dw_packed = None


class _BNRec(object):
    __slots__ = ("mod", "C", "idx", "name")

    def __init__(self, mod, idx, name):
        self.mod = mod
        self.C = mod.weight.numel()
        self.idx = idx
        self.name = name


class _BlockRec(object):
    __slots__ = ("conv1", "bn1", "conv2", "bn2", "convd", "bnd", "name")


class SpeakerNetEngine(object):
    def __init__(self, model, precision="bf16", impl=None):
        if precision not in ("bf16", "fp32"):
            raise ValueError("precision must be 'bf16' or 'fp32'")
        self.model = model
        self.precision = precision
        self.act_dtype = torch.bfloat16 if precision == "bf16" else torch.float32
        self.dcode = lib.BF16 if precision == "bf16" else lib.F32
        if impl is None:
            impl = "tcgen05" if precision == "bf16" else "simt"
        if impl not in ("tcgen05", "simt"):
            raise ValueError("impl must be 'tcgen05' or 'simt'")
        if impl == "tcgen05" and precision != "bf16":
            raise ValueError("the tcgen05 convolution path is bf16 only")
        self.impl = lib.IMPL_TCGEN05 if impl == "tcgen05" else lib.IMPL_SIMT
        self.device = None
        self._ws = {}
        self._flat = None
        self._eval_ready = False
        self._packed_version = -1
        self._param_version = 0
        self._train_ws = {}             # (B, F, T) -> list of training workspaces; one per forward whose backward is pending
        self._grads_fresh = False       # a backward wrote the flat gradient buffer and no optimizer step / zero_grad followed
        self.grad_ready_cb = None       # called as cb(bucket_index) during backward (data-parallel hook)
        self.fuse_bn_bwd = os.environ.get("SVK_DISABLE_BN_FUSE", "0") != "1"   # A/B switch for the dgrad-epilogue fusion
        # Weight gradients run on a side stream: they only feed the optimizer / the gradient all-reduce, so their (tensor-
        # bound) kernels can share the SMs with the HBM-bound BatchNorm passes of the main chain instead of queueing
        # between them.  SVK_WGRAD_STREAM=0 keeps everything on one stream.
        self.wgrad_side = os.environ.get("SVK_WGRAD_STREAM", "1") != "0"
        self.side_shortcut = os.environ.get("SVK_SIDE_SHORTCUT", "1") != "0"    # forward: 1x1/s2 shortcut convs on the side stream
        self._side_stream = None
        self._side_pending = {}
        # test hook: delay every side-stream weight gradient by this many GPU cycles, so that a missing stream dependency
        # (the main stream overwriting a buffer the side stream still reads) produces wrong gradients deterministically
        self._side_delay = int(os.environ.get("SVK_WGRAD_STREAM_DELAY_CYCLES", "0"))
        self.fused_head = os.environ.get("SVK_DISABLE_FUSED_HEAD", "0") != "1"     # A/B switch: csrc/aam_fused.cu vs ~17 small launches
        self.last_head = None           # {"loss", "rank"} of the most recent fused AAM head forward
        self.debug_masked = set()       # debug taps that hold the gradient already multiplied by the ReLU mask
        self.debug = None               # tests set this to a dict to capture per-layer gradients (clones)
        self._index_modules()

    # ------------------------------------------------------------------------------------------ structure
    def _index_modules(self):
        m = self.model
        res = m.res
        self.convs, self.bns, self.blocks = [], [], []
        self.stem_conv = res.conv1
        self.stem_bn = self._bn(res.bn1, "res.bn1")
        for li in range(1, 5):
            layer = getattr(res, "layer%d" % li)
            for bi, blk in enumerate(layer):
                b = _BlockRec()
                b.name = "res.layer%d.%d" % (li, bi)
                b.conv1 = self._conv(blk.conv1, b.name + ".conv1")
                b.bn1 = self._bn(blk.bn1, b.name + ".bn1")
                b.conv2 = self._conv(blk.conv2, b.name + ".conv2")
                b.bn2 = self._bn(blk.bn2, b.name + ".bn2")
                if blk.downsample is not None:
                    b.convd = self._conv(blk.downsample[0], b.name + ".downsample.0")
                    b.bnd = self._bn(blk.downsample[1], b.name + ".downsample.1")
                else:
                    b.convd = b.bnd = None
                self.blocks.append(b)
        self.head_bn = self._bn(m.bn1, "bn1") if hasattr(m, "bn1") else None
        self.c0 = self.stem_conv.weight.shape[0]
        self.c_last = self.blocks[-1].conv2.cout
        # gradient buckets (reverse execution order): 0 = head + fc1, 1 = layer4, 2 = everything before
        self.bucket_of_block = [1 if b.name.startswith("res.layer4") else 2 for b in self.blocks]

    def _conv(self, mod, name):
        r = _ConvRec(mod, name)
        self.convs.append(r)
        return r

    def _bn(self, mod, name):
        r = _BNRec(mod, len(self.bns), name)
        self.bns.append(r)
        return r

    # ------------------------------------------------------------------------------------------ parameters
    def params_in_order(self):
        return [p for p in self.model.parameters()]

    def ensure_device(self):
        """Flatten parameters into one fp32 buffer (and gradients into another) on the module's CUDA device."""
        params = self.params_in_order()
        dev = params[0].device
        if dev.type != "cuda":
            raise lib.SvkError("NeuralSpeakerModel runs on CUDA only (no CPU fallback): call .cuda() first")
        if self._flat is not None and self._flat_ok(params):
            return
        lib.load()
        self.device = dev
        with torch.no_grad():
            offs, total = [], 0
            for p in params:
                if p.dtype != torch.float32:
                    raise lib.SvkError("parameters must be fp32 master weights")
                offs.append(total)
                total += (p.numel() + 3) // 4 * 4          # keep every view 16-byte aligned
            flat = torch.zeros(total, dtype=torch.float32, device=dev)
            grad = torch.zeros(total, dtype=torch.float32, device=dev)
            for p, o in zip(params, offs):
                v = flat[o:o + p.numel()].view(p.shape)
                v.copy_(p.data)
                p.data = v
                p._svk_flat = (self, o)
                p.grad = None
            self._flat, self._grad, self._offs, self._total = flat, grad, offs, total
            self._grad_views = [grad[o:o + p.numel()].view(p.shape) for p, o in zip(params, offs)]
            self._params = params
            self._gview = {id(p): g for p, g in zip(params, self._grad_views)}
            # bucket ranges in the flat buffer: [stem..layer3], [layer4], [fc1, head]
            names = [n for n, _ in self.model.named_parameters()]
            first_l4 = next(i for i, n in enumerate(names) if n.startswith("res.layer4"))
            first_tail = next(i for i, n in enumerate(names) if not n.startswith("res."))
            self.bucket_ranges = {2: (0, offs[first_l4]), 1: (offs[first_l4], offs[first_tail]),
                                  0: (offs[first_tail], total)}
            # packed conv weights + packed conv weight-gradient scratch
            nconv = sum(c.cout * c.cin * c.R * c.R for c in self.convs)
            self._wf = torch.empty(nconv, dtype=self.act_dtype, device=dev)
            self._wd = torch.empty(nconv, dtype=self.act_dtype, device=dev)
            o = 0
            poff = {id(p): off for p, off in zip(params, offs)}
            rows = []
            for c in self.convs:
                n = c.cout * c.cin * c.R * c.R
                c.w_fwd, c.w_dgrad = self._wf[o:o + n], self._wd[o:o + n]
                rows.append([poff[id(c.mod.weight)], o, c.cout, c.cin, c.R * c.R, o])
                o += n
            self._pack_table = torch.tensor(rows, dtype=torch.int64, device=dev)
            self._pack_total = o
            nb = len(self.bns)
            self._nbt = torch.zeros(nb, dtype=torch.long, device=dev)
            for bn in self.bns:
                self._nbt[bn.idx] = bn.mod.num_batches_tracked.to(dev)
                bn.mod._buffers["num_batches_tracked"] = self._nbt[bn.idx]
            self._cmax = max(b.C for b in self.bns)
            self._coef = torch.zeros(nb, 4, self._cmax, dtype=torch.float32, device=dev)  # scale, shift, mean, rstd
            self._stats = torch.zeros(nb, 2 * self._cmax, dtype=torch.float64, device=dev)
            self._bsums = torch.zeros(nb, 3 * self._cmax, dtype=torch.float64, device=dev)
        self._ws = {}
        self._train_ws = {}
        self._eval_ready = False
        self._packed_version = -1

    def _flat_ok(self, params):
        p0, p1 = params[0], params[-1]
        return (getattr(p0, "_svk_flat", (None,))[0] is self and p0.data_ptr() == self._flat.data_ptr() and
                p1.data_ptr() == self._flat.data_ptr() + 4 * self._offs[-1] and len(params) == len(self._offs))

    def invalidate(self):
        """Parameters or buffers changed behind our back (load_state_dict, optimizer step, .train())."""
        self._param_version += 1
        self._eval_ready = False

    @property
    def flat_params(self):
        return self._flat

    @property
    def flat_grads(self):
        return self._grad

    def _pack_weights(self):
        if self._packed_version == self._param_version:
            return
        call.svk_pack_conv_weights_batched(self._flat.data_ptr(), self._wf.data_ptr(), self._wd.data_ptr(),
                                           self._pack_table.data_ptr(), len(self.convs), self._pack_total, self.dcode,
                                           _stream())
        self._packed_version = self._param_version

    # ------------------------------------------------------------------------------------------ workspaces
    def _workspace(self, key):
        ws = self._ws.get(key)
        if ws is None:
            if len(self._ws) > 8:       # variable-length extraction: do not hoard one arena per length
                self._ws.clear()
            ws = {}
            self._ws[key] = ws
        return ws

    def _lease_train_ws(self, key):
        """A training workspace no pending backward still reads.  Saved-for-backward state (activations, BatchNorm
        coefficients) lives in the workspace, so a second training-mode forward before backward() gets its OWN workspace
        instead of overwriting the first one's (re-entrancy: two model(x, y) calls, or predict() in training mode)."""
        pool = self._train_ws.setdefault(key, [])
        for ws in pool:
            if not ws["_busy"]:
                return ws
        if len(pool) >= 4:
            raise lib.SvkError("%d training forwards of shape %s are waiting for their backward: call backward() (or run "
                               "the extra forwards under torch.no_grad())" % (len(pool), (key,)))
        if not pool and len(self._train_ws) > 4:      # a new batch shape: drop idle workspaces of other shapes
            for k in [k for k, v in self._train_ws.items() if k != key and not any(w["_busy"] for w in v)]:
                del self._train_ws[k]
        ws = {"_busy": False,
              "coef": torch.zeros(len(self.bns), 4, self._cmax, dtype=torch.float32, device=self.device)}
        pool.append(ws)
        return ws

    def _buf(self, ws, name, shape, dtype=None):
        t = ws.get(name)
        if t is None:
            t = torch.empty(shape, dtype=dtype or self.act_dtype, device=self.device)
            ws[name] = t
        return t

    def _arena(self, name, shape, dtype=None):
        """Shape-agnostic scratch (variable-length extraction): a flat buffer per name, grown geometrically."""
        dtype = dtype or self.act_dtype
        n = 1
        for s in shape:
            n *= s
        t = self._arena_bufs.get(name) if hasattr(self, "_arena_bufs") else None
        if t is None or t.numel() < n or t.dtype != dtype or t.device != self.device:
            if not hasattr(self, "_arena_bufs"):
                self._arena_bufs = {}
            t = torch.empty(int(n * 1.25) + 64, dtype=dtype, device=self.device)
            self._arena_bufs[name] = t
        return t[:n].view(shape)

    def _coefs(self, bn):
        c = self._coef[bn.idx]
        return c[0], c[1], c[2], c[3]

    # ------------------------------------------------------------------------------------------ kernels, thin wrappers
    def _desc(self, N, H, W, conv):
        return make_conv_desc(N, H, W, conv.cin, conv.cout, conv.R, conv.stride, self.dcode, self.impl)

    def _conv_fwd(self, d, conv, x, y, stats=None, scale=None, shift=None, res=None, relu=0, valid=None):
        call.svk_conv2d_fwd(d, x.data_ptr(), conv.w_fwd.data_ptr(), y.data_ptr(), _ptr(stats), _ptr(scale), _ptr(shift),
                            _ptr(res), relu, _ptr(valid), _stream())

    def _bn_train(self, bn, c, M, stats_done=False):
        """Batch statistics of c ([M, C]) -> scale/shift (+ running-stat update)."""
        st = _stream()
        stats = self._stats[bn.idx]
        if not stats_done:
            call.svk_channel_stats(c.data_ptr(), M, bn.C, c_dtype_code(c), stats.data_ptr(), st)
        sc, sh, mu, rs = self._coefs(bn)
        m = bn.mod
        call.svk_bn_finalize(stats.data_ptr(), M, bn.C, m.weight.data_ptr(), m.bias.data_ptr(),
                             m.running_mean.data_ptr(), m.running_var.data_ptr(), _MOM, _EPS, sc.data_ptr(),
                             sh.data_ptr(), mu.data_ptr(), rs.data_ptr(), st)
        return sc, sh

    def _bn_train_act(self, bn, x, out, M, relu, res=None, bn_b=None):
        """Training BN (statistics already accumulated in self._stats) + activation (+ residual [through bn_b])."""
        m = bn.mod
        coef = self._coef[bn.idx]
        if bn_b is not None:
            mb = bn_b.mod
            bargs = (self._stats[bn_b.idx].data_ptr(), mb.weight.data_ptr(), mb.bias.data_ptr(), mb.running_mean.data_ptr(),
                     mb.running_var.data_ptr(), self._coef[bn_b.idx].data_ptr())
        else:
            bargs = (0, 0, 0, 0, 0, 0)
        call.svk_bn_train_act_fwd(x.data_ptr(), self._stats[bn.idx].data_ptr(), m.weight.data_ptr(), m.bias.data_ptr(),
                                  m.running_mean.data_ptr(), m.running_var.data_ptr(), coef.data_ptr(), _ptr(res),
                                  bargs[0], bargs[1], bargs[2], bargs[3], bargs[4], bargs[5], self._cmax, _MOM, _EPS, relu,
                                  out.data_ptr(), M, bn.C, c_dtype_code(x), _stream())

    def _bn_act(self, x, sc, sh, out, M, C, relu, res=None, rsc=None, rsh=None):
        call.svk_bn_act_fwd(x.data_ptr(), sc.data_ptr(), sh.data_ptr(), _ptr(res), _ptr(rsc), _ptr(rsh), relu,
                            out.data_ptr(), M, C, c_dtype_code(x), _stream())

    def _gemm(self, A, lda, a_k, Bm, ldb, b_k, C, ldc, M, N, K, bias=None):
        """C[M,N] = A[M,K] * B[N,K]^T (+bias).  a_k / b_k: operand stored with K contiguous (else transposed view).
        Product mode: tcgen05 kind::tf32 GEMM; fp32 validation mode: CUDA-core fp32 GEMM."""
        st = _stream()
        if self.precision == "bf16" and lda % 4 == 0 and ldb % 4 == 0:
            need = lib.load().svk_gemm_tf32_workspace_bytes(M, N, K)
            ws = self._arena("gemm_ws", ((need + 3) // 4 + 4,), torch.float32) if need else None
            call.svk_gemm_tf32(A.data_ptr(), lda, 1 if a_k else 0, Bm.data_ptr(), ldb, 1 if b_k else 0, C.data_ptr(), ldc,
                               M, N, K, _ptr(bias), _ptr(ws), need, st)
        else:
            a_sm, a_sk = (lda, 1) if a_k else (1, lda)
            b_sk, b_sn = (1, ldb) if b_k else (ldb, 1)
            call.svk_sgemm(A.data_ptr(), a_sm, a_sk, Bm.data_ptr(), b_sk, b_sn, C.data_ptr(), ldc, M, N, K, 1.0, 0.0,
                           _ptr(bias), st)

    # ------------------------------------------------------------------------------------------ training forward
    def forward_train(self, x, y, with_head=True, save=True):
        """x: (B, F, T) fp32 CUDA, y: (B,) int64 -> (logits (B, spk_num) fp32, saved state for backward_train).
        with_head=False returns the embeddings only (predict() in training mode: batch statistics, nothing saved)."""
        self.ensure_device()
        self._param_version += 1        # parameters change every optimiser step: always re-pack, never reuse eval caches
        self._eval_ready = False
        self._pack_weights()
        model = self.model
        B, F, T = x.shape
        x = x.contiguous().float()
        ws = self._lease_train_ws((B, F, T))
        self._coef = ws["coef"]         # scale, shift, mean, rstd of every BatchNorm of THIS forward (read again by its backward)
        st = _stream()
        self._stats.zero_()
        self._nbt.add_(1)
        sv = {"x": x, "B": B, "F": F, "T": T, "ws": ws, "blocks": []}
        # ---- stem
        C0 = self.c0
        M = B * F * T
        c0 = self._buf(ws, "c0", (B, F, T, C0))
        a0 = self._buf(ws, "a0", (B, F, T, C0))
        call.svk_stem_conv_fwd(x.data_ptr(), self.stem_conv.weight.data_ptr(), c0.data_ptr(), B, F, T, C0, self.dcode,
                               0, 0, 0, 0, self._stats[self.stem_bn.idx].data_ptr(), st)
        self._bn_train_act(self.stem_bn, c0, a0, M, 1)
        cur, H, W = a0, F, T
        # ---- residual blocks
        for bi, b in enumerate(self.blocks):
            d1 = self._desc(B, H, W, b.conv1)
            Ho, Wo = d1.Ho, d1.Wo
            Mo = B * Ho * Wo
            Co = b.conv1.cout
            c1 = self._buf(ws, "c1_%d" % bi, (B, Ho, Wo, Co))
            a1 = self._buf(ws, "a1_%d" % bi, (B, Ho, Wo, Co))
            c2 = self._buf(ws, "c2_%d" % bi, (B, Ho, Wo, Co))
            out = self._buf(ws, "o_%d" % bi, (B, Ho, Wo, Co))
            cd = dd = None
            cd_done = None
            if b.convd is not None:
                # the 1x1/s2 shortcut convolution depends only on the block input: on the side stream it runs beside the
                # BatchNorm passes of the main branch; the main stream waits for it before the block's last BatchNorm
                dd = self._desc(B, H, W, b.convd)
                cd = self._buf(ws, "cd_%d" % bi, (B, Ho, Wo, Co))
                if self.wgrad_side and self.side_shortcut:
                    main = torch.cuda.current_stream()
                    if self._side_stream is None or self._side_stream.device != self.device:
                        self._side_stream = torch.cuda.Stream(device=self.device)
                    ready = torch.cuda.Event()
                    ready.record(main)               # `cur` (and the zeroed statistics table) are complete
                    self._side_stream.wait_event(ready)
                    with torch.cuda.stream(self._side_stream):
                        if self._side_delay:
                            torch.cuda._sleep(self._side_delay)
                        self._conv_fwd(dd, b.convd, cur, cd, stats=self._stats[b.bnd.idx])
                        cd_done = torch.cuda.Event()
                        cd_done.record(self._side_stream)
            self._conv_fwd(d1, b.conv1, cur, c1, stats=self._stats[b.bn1.idx])
            self._bn_train_act(b.bn1, c1, a1, Mo, 1)
            d2 = self._desc(B, Ho, Wo, b.conv2)
            self._conv_fwd(d2, b.conv2, a1, c2, stats=self._stats[b.bn2.idx])
            if b.convd is not None:
                if cd_done is not None:
                    torch.cuda.current_stream().wait_event(cd_done)
                else:
                    self._conv_fwd(dd, b.convd, cur, cd, stats=self._stats[b.bnd.idx])
                self._bn_train_act(b.bn2, c2, out, Mo, 1, res=cd, bn_b=b.bnd)
            else:
                self._bn_train_act(b.bn2, c2, out, Mo, 1, res=cur)
            sv["blocks"].append((cur, H, W, c1, a1, c2, cd, out, d1, d2, dd, Ho, Wo))
            cur, H, W = out, Ho, Wo
        # ---- pooling + embedding FC
        mode = 1 if model.pool.pooling == "mean+std" else 0
        Cl = self.c_last
        pdim = Cl * H * (2 if mode else 1)
        pooled = self._buf(ws, "pooled", (B, pdim), torch.float32)
        call.svk_statspool_fwd(cur.data_ptr(), pooled.data_ptr(), B, H, W, Cl, mode, 0, self.dcode, st)
        fc = model.fc1
        E = fc.weight.shape[0]
        emb = self._buf(ws, "emb", (B, E), torch.float32)
        self._gemm(pooled, pdim, True, fc.weight, pdim, True, emb, E, B, E, pdim, fc.bias)
        if not with_head:
            return emb.clone()
        sv.update(last=cur, Hl=H, Wl=W, mode=mode, pooled=pooled, emb=emb, pdim=pdim, E=E)
        logits = self._head_fwd(emb, y, ws, sv, train=True)
        if save:
            ws["_busy"] = True          # released by backward_train, or when autograd drops the graph (_Lease.__del__)
            sv["lease"] = _Lease(ws)
        return logits, sv

    # ------------------------------------------------------------------------------------------ heads
    def _head_fwd(self, emb, y, ws, sv, train):
        model = self.model
        st = _stream()
        B, E = emb.shape
        loss = model.loss
        h = emb
        if loss in ("softmax", "AAM-v1"):
            bn = self.head_bn
            h = self._buf(ws, "head_h", (B, E), torch.float32)
            if train:
                sc, sh = self._bn_train(bn, emb, B)
            else:
                sc, sh = self._eval_coefs(bn)
            call.svk_bn_act_fwd(emb.data_ptr(), sc.data_ptr(), sh.data_ptr(), 0, 0, 0, 1, h.data_ptr(), B, E, lib.F32, st)
        last = model.last
        C = last.weight.shape[0]
        logits = torch.empty(B, C, dtype=torch.float32, device=self.device)
        if loss == "softmax":
            self._gemm(h, E, True, last.weight, E, True, logits, C, B, C, E, last.bias)
        else:
            if y is None:
                raise lib.SvkError("AAM head needs labels: call model(x, y)")
            y = y.contiguous()
            if E == 256 and self.fused_head:
                # ONE fused pass (csrc/aam_fused.cu): row normalisation, cosine GEMM, margin, scale, log-sum-exp, loss, rank
                need = lib.load().svk_aam_ce_workspace_bytes(B, E, C)
                hws = self._arena("aam_ws", ((need + 3) // 4,), torch.float32)
                cos_t = self._buf(ws, "aam_cos_t", (B,), torch.float32)
                lse = self._buf(ws, "aam_lse", (B,), torch.float32)
                rank = torch.empty(B, dtype=torch.int32, device=self.device)
                loss_mean = torch.empty((), dtype=torch.float32, device=self.device)
                call.svk_aam_ce_fwd(h.data_ptr(), last.weight.data_ptr(), y.data_ptr(), logits.data_ptr(), cos_t.data_ptr(),
                                    lse.data_ptr(), 0, rank.data_ptr(), loss_mean.data_ptr(), B, E, C, last.cos_m, last.sin_m,
                                    last.th, last.mm, float(last.s), 1 if self.precision == "fp32" else 0, hws.data_ptr(),
                                    hws.numel() * 4, st)
                self.last_head = {"loss": loss_mean, "rank": rank}
                if sv is not None:
                    # detach(): the returned tensor gets the autograd node, which owns this state — no reference cycle
                    sv.update(y=y, cos_t=cos_t, lse=lse, logits=logits.detach(), fused=True)
            else:
                xh, xinv, wh, winv = self._aam_normalise(h, last.weight, ws, B, E, C)
                cos_t = self._buf(ws, "aam_cos_t", (B,), torch.float32)
                self._gemm(xh, E, True, wh, E, True, logits, C, B, C, E)
                call.svk_aam_margin_fwd(logits.data_ptr(), y.data_ptr(), cos_t.data_ptr(), B, C, last.cos_m, last.sin_m,
                                        last.th, last.mm, float(last.s), st)
                self.last_head = None
                if sv is not None:
                    sv.update(y=y, xh=xh, xinv=xinv, wh=wh, winv=winv, cos_t=cos_t, fused=False)
        if sv is not None:
            sv.update(h=h, C=C)
        return logits

    def _aam_normalise(self, h, weight, ws, B, E, C):
        st = _stream()
        xh = self._buf(ws, "aam_xh", (B, E), torch.float32)
        xinv = self._buf(ws, "aam_xinv", (B,), torch.float32)
        wh = self._buf(ws, "aam_wh", (C, E), torch.float32)
        winv = self._buf(ws, "aam_winv", (C,), torch.float32)
        call.svk_l2norm_rows_fwd(h.data_ptr(), xh.data_ptr(), xinv.data_ptr(), B, E, 1e-12, st)
        call.svk_l2norm_rows_fwd(weight.data_ptr(), wh.data_ptr(), winv.data_ptr(), C, E, 1e-12, st)
        return xh, xinv, wh, winv

    def _head_bwd(self, dlogits, sv, dloss=None):
        """d emb (B, E) from the upstream gradient; writes the head's parameter gradients.  Either dlogits (B, C) fp32
        (consumed in place), or — fused loss path — dloss, the scalar gradient of the mean cross-entropy."""
        model = self.model
        st = _stream()
        ws = sv["ws"]
        B, E, C = sv["B"], sv["E"], sv["C"]
        loss = model.loss
        last = model.last
        gw = self._gview[id(last.weight)]
        h = sv["h"]
        dh = self._buf(ws, "d_h", (B, E), torch.float32)
        if dloss is not None:
            # fused backward (csrc/aam_fused.cu): d_logits recomputed from the logits + lse, both normalisation Jacobians inside
            need = lib.load().svk_aam_ce_workspace_bytes(B, E, C)
            hws = self._arena("aam_ws", ((need + 3) // 4,), torch.float32)
            call.svk_aam_ce_bwd(h.data_ptr(), last.weight.data_ptr(), sv["y"].data_ptr(), sv["logits"].data_ptr(),
                                sv["lse"].data_ptr(), sv["cos_t"].data_ptr(), dloss.data_ptr(), dh.data_ptr(), gw.data_ptr(), B, E, C,
                                last.cos_m, last.sin_m, last.th, float(last.s), 1 if self.precision == "fp32" else 0,
                                hws.data_ptr(), hws.numel() * 4, st)
        elif loss == "softmax":
            Cp = dlogits.stride(0)                  # padded row pitch of the upstream-gradient copy
            self._gemm(dlogits, Cp, True, last.weight, E, False, dh, E, B, E, C)           # dh = dlogits * W
            self._gemm(dlogits, Cp, False, h, E, False, gw, E, C, E, B)                    # dW = dlogits^T * h
            call.svk_colsum(dlogits.data_ptr(), self._gview[id(last.bias)].data_ptr(), B, C, Cp, st)
        else:
            Cp = dlogits.stride(0)
            if sv.get("fused"):                     # the fused forward kept no normalised copies: rebuild them
                xh, xinv, wh, winv = self._aam_normalise(h, last.weight, ws, B, E, C)
            else:
                xh, xinv, wh, winv = sv["xh"], sv["xinv"], sv["wh"], sv["winv"]
            call.svk_aam_margin_bwd(dlogits.data_ptr(), sv["y"].data_ptr(), sv["cos_t"].data_ptr(), B, C, Cp, last.cos_m,
                                    last.sin_m, last.th, float(last.s), st)
            dxh = self._buf(ws, "d_xh", (B, E), torch.float32)
            dwh = self._buf(ws, "d_wh", (C, E), torch.float32)
            self._gemm(dlogits, Cp, True, wh, E, False, dxh, E, B, E, C)                   # dx_hat = dcos * W_hat
            self._gemm(dlogits, Cp, False, xh, E, False, dwh, E, C, E, B)                  # dW_hat = dcos^T * x_hat
            call.svk_l2norm_rows_bwd(dxh.data_ptr(), xh.data_ptr(), xinv.data_ptr(), dh.data_ptr(), B, E, st)
            call.svk_l2norm_rows_bwd(dwh.data_ptr(), wh.data_ptr(), winv.data_ptr(), gw.data_ptr(), C, E, st)
        if loss in ("softmax", "AAM-v1"):
            bn = self.head_bn
            sc, sh, mu, rs = self._coefs(bn)
            sums = self._bsums[bn.idx]
            demb = self._buf(ws, "d_emb", (B, E), torch.float32)
            emb = sv["emb"]
            call.svk_bn_bwd_reduce(dh.data_ptr(), h.data_ptr(), emb.data_ptr(), mu.data_ptr(), rs.data_ptr(), 0, 0, 0,
                                   sums.data_ptr(), B, E, lib.F32, st)
            call.svk_bn_bwd_apply(dh.data_ptr(), h.data_ptr(), emb.data_ptr(), mu.data_ptr(), rs.data_ptr(),
                                  bn.mod.weight.data_ptr(), demb.data_ptr(), 0, 0, 0, 0, 0, sums.data_ptr(),
                                  self._gview[id(bn.mod.weight)].data_ptr(), self._gview[id(bn.mod.bias)].data_ptr(),
                                  0, 0, B, E, lib.F32, st)
            return demb
        return dh

    # ------------------------------------------------------------------------------------------ training backward
    def backward_train(self, dlogits, sv, dloss=None):
        """Backward of the forward that produced `sv`: from dlogits (B, C), or (fused loss path) from dloss, the scalar
        gradient of the mean cross-entropy the fused head computed.  Parameter gradients are WRITTEN into the flat gradient buffer
        (p.grad views); when the previous backward's gradients were not consumed by SGD.step() / zero_grad() they are
        added on top, which is torch's accumulation semantics (the reference zeroes them every step, train_resnet.py:326)."""
        if sv is None or sv.get("done"):
            raise lib.SvkError("backward called twice for the same forward (retain_graph is not supported), or without one")
        sv["done"] = True
        keep = None
        if self._grads_fresh:
            self._join_side()
            keep = self._grad.clone()
        self._coef = sv["ws"]["coef"]
        model = self.model
        st = _stream()
        ws = sv["ws"]
        B = sv["B"]
        self._bsums.zero_()
        if dloss is not None:
            if not sv.get("fused"):
                raise lib.SvkError("the fused loss backward needs the fused AAM head forward")
            demb = self._head_bwd(None, sv, dloss=dloss.contiguous().float())
        else:
            # private copy of the upstream gradient (the head backward works in place) with a row pitch that is a multiple
            # of 4 floats, so the tensor-core GEMMs can map it with TMA
            Bq, Cq = dlogits.shape
            Cp = (Cq + 3) // 4 * 4
            pad = self._buf(ws, "d_logits_pad", (Bq, Cp), torch.float32)
            pad[:, :Cq].copy_(dlogits)
            if Cp != Cq:
                pad[:, Cq:].zero_()
            demb = self._head_bwd(pad[:, :Cq], sv)
        # ---- fc1
        fc = model.fc1
        pdim, E = sv["pdim"], sv["E"]
        pooled = sv["pooled"]
        dpool = self._buf(ws, "d_pool", (B, pdim), torch.float32)
        self._gemm(demb, E, True, fc.weight, pdim, False, dpool, pdim, B, pdim, E)                     # dpool = demb * W
        self._gemm(demb, E, False, pooled, pdim, False, self._gview[id(fc.weight)], pdim, E, pdim, B)  # dW = demb^T * pooled
        call.svk_colsum(demb.data_ptr(), self._gview[id(fc.bias)].data_ptr(), B, E, E, st)
        self._bucket_done(0)
        # ---- pooling
        last, Hl, Wl, Cl = sv["last"], sv["Hl"], sv["Wl"], self.c_last
        gmax = max(t[0].numel() for t in sv["blocks"])
        gmax = max(gmax, max(t[7].numel() for t in sv["blocks"]))
        gbuf = [self._buf(ws, "g%d" % i, (gmax,)) for i in range(5)]
        dO = gbuf[0][:last.numel()].view(last.shape)
        call.svk_statspool_bwd(last.data_ptr(), dpool.data_ptr(), dO.data_ptr(), B, Hl, Wl, Cl, sv["mode"],
                               1 if self.fuse_bn_bwd else 0, self.dcode, st)
        cur_bucket = 1
        # BatchNorm-backward fusion (svk_conv2d_dgrad_bn): every stride-1 data-gradient epilogue applies the ReLU mask
        # of the layer below and accumulates that layer's (sum g, sum g*xhat); the stand-alone reduce pass survives only
        # where a gradient is assembled by several launches (downsample blocks) or comes from the pooling layer.
        fuse = self.fuse_bn_bwd
        masked = fuse           # dO already carries the ReLU mask of `out` (the pooling backward applied it)
        reduced = False         # ... and bn2's sums are already in self._bsums
        if self.debug is not None:
            self.debug_masked = set()

        def bn_fuse(mask, c=None, bn=None):
            if c is None:
                return lib.BnBwdFuse(mask.data_ptr(), None, None, None, None)
            _, _, mu, rs = self._coefs(bn)
            return lib.BnBwdFuse(mask.data_ptr(), c.data_ptr(), mu.data_ptr(), rs.data_ptr(), self._bsums[bn.idx].data_ptr())

        # ---- blocks in reverse
        for bi in range(len(self.blocks) - 1, -1, -1):
            b = self.blocks[bi]
            if self.bucket_of_block[bi] != cur_bucket:
                self._bucket_done(cur_bucket)
                cur_bucket = self.bucket_of_block[bi]
            xin, H, W, c1, a1, c2, cd, out, d1, d2, dd, Ho, Wo = sv["blocks"][bi]
            Mo = B * Ho * Wo
            Co = b.conv1.cout
            n_out = out.numel()
            dc2 = gbuf[1][:n_out].view(out.shape)
            da1 = gbuf[2][:n_out].view(out.shape)
            dc1 = gbuf[3][:n_out].view(out.shape)
            dcd = gbuf[4][:n_out].view(out.shape) if cd is not None else None
            # free ping-pong slot for dx: dO lives in gbuf[0]; dx is written to gbuf[2] (da1 is dead by then)
            _, _, mu2, rs2 = self._coefs(b.bn2)
            sums2 = self._bsums[b.bn2.idx]
            g2 = b.bn2.mod
            omask = 0 if masked else out.data_ptr()
            self._wait_side("dc2")           # the previous block's side-stream weight gradients still read these buffers
            self._wait_side("dcd")
            if cd is not None:
                _, _, mud, rsd = self._coefs(b.bnd)
                gd = b.bnd.mod
                call.svk_bn_bwd_reduce(dO.data_ptr(), omask, c2.data_ptr(), mu2.data_ptr(), rs2.data_ptr(),
                                       cd.data_ptr(), mud.data_ptr(), rsd.data_ptr(), sums2.data_ptr(), Mo, Co, self.dcode, st)
                call.svk_bn_bwd_apply(dO.data_ptr(), omask, c2.data_ptr(), mu2.data_ptr(), rs2.data_ptr(),
                                      g2.weight.data_ptr(), dc2.data_ptr(), cd.data_ptr(), mud.data_ptr(), rsd.data_ptr(),
                                      gd.weight.data_ptr(), dcd.data_ptr(), sums2.data_ptr(),
                                      self._gview[id(g2.weight)].data_ptr(), self._gview[id(g2.bias)].data_ptr(),
                                      self._gview[id(gd.weight)].data_ptr(), self._gview[id(gd.bias)].data_ptr(),
                                      Mo, Co, self.dcode, st)
            else:
                if not reduced:
                    call.svk_bn_bwd_reduce(dO.data_ptr(), omask, c2.data_ptr(), mu2.data_ptr(), rs2.data_ptr(),
                                           0, 0, 0, sums2.data_ptr(), Mo, Co, self.dcode, st)
                call.svk_bn_bwd_apply(dO.data_ptr(), omask, c2.data_ptr(), mu2.data_ptr(), rs2.data_ptr(),
                                      g2.weight.data_ptr(), dc2.data_ptr(), 0, 0, 0, 0, 0, sums2.data_ptr(),
                                      self._gview[id(g2.weight)].data_ptr(), self._gview[id(g2.bias)].data_ptr(), 0, 0,
                                      Mo, Co, self.dcode, st)
            # conv2: weight gradient + data gradient
            if self.debug is not None:
                self.debug[b.name] = dO.clone()
                if masked:
                    self.debug_masked.add(b.name)
                self.debug[b.name + ".conv2"] = dc2.clone()
                if dcd is not None:
                    self.debug[b.name + ".downsample.0"] = dcd.clone()
            self._wgrad(d2, b.conv2, a1, dc2, "dc2")
            # bn1 (+ReLU mask from a1): mask and sums come out of conv2's data-gradient epilogue
            _, _, mu1, rs1 = self._coefs(b.bn1)
            sums1 = self._bsums[b.bn1.idx]
            g1 = b.bn1.mod
            if fuse:
                call.svk_conv2d_dgrad_bn(d2, dc2.data_ptr(), b.conv2.w_dgrad.data_ptr(), da1.data_ptr(), 0, 0, 0,
                                         bn_fuse(a1, c1, b.bn1), st)
            else:
                call.svk_conv2d_dgrad(d2, dc2.data_ptr(), b.conv2.w_dgrad.data_ptr(), da1.data_ptr(), 0, 0, 0, st)
                call.svk_bn_bwd_reduce(da1.data_ptr(), a1.data_ptr(), c1.data_ptr(), mu1.data_ptr(), rs1.data_ptr(), 0, 0, 0,
                                       sums1.data_ptr(), Mo, Co, self.dcode, st)
            self._wait_side("dc1")
            call.svk_bn_bwd_apply(da1.data_ptr(), 0 if fuse else a1.data_ptr(), c1.data_ptr(), mu1.data_ptr(), rs1.data_ptr(),
                                  g1.weight.data_ptr(), dc1.data_ptr(), 0, 0, 0, 0, 0, sums1.data_ptr(),
                                  self._gview[id(g1.weight)].data_ptr(), self._gview[id(g1.bias)].data_ptr(), 0, 0,
                                  Mo, Co, self.dcode, st)
            if self.debug is not None:
                self.debug[b.name + ".conv1"] = dc1.clone()
            self._wgrad(d1, b.conv1, xin, dc1, "dc1")
            dx = gbuf[2][:xin.numel()].view(xin.shape)
            # which BatchNorm consumes dx: the stem's, or bn2 of the block below (whose sums can only be fused when that
            # block has no downsample BN sharing the gradient)
            if bi == 0:
                below = (ws["c0"], self.stem_bn)
            elif self.blocks[bi - 1].convd is None:
                below = (sv["blocks"][bi - 1][5], self.blocks[bi - 1].bn2)
            else:
                below = None
            if fuse:
                bnf = bn_fuse(xin, *below) if below is not None else bn_fuse(xin)
            if cd is not None:
                self._wgrad(dd, b.convd, xin, dcd, "dcd")
                if fuse:
                    call.svk_downsample_dgrad_bn(d1, dc1.data_ptr(), b.conv1.w_dgrad.data_ptr(), dd, dcd.data_ptr(),
                                                 b.convd.w_dgrad.data_ptr(), dx.data_ptr(), bnf, st)
                else:
                    call.svk_conv2d_dgrad(d1, dc1.data_ptr(), b.conv1.w_dgrad.data_ptr(), dx.data_ptr(), 0, 0, 0, st)
                    # 1x1/s2 data gradient accumulates into the block-input gradient (res aliases dx)
                    call.svk_conv2d_dgrad(dd, dcd.data_ptr(), b.convd.w_dgrad.data_ptr(), dx.data_ptr(), dx.data_ptr(), 0, 0, st)
            elif fuse:
                # identity shortcut: + dO (already masked), then the mask of xin and the sums for the BatchNorm below
                call.svk_conv2d_dgrad_bn(d1, dc1.data_ptr(), b.conv1.w_dgrad.data_ptr(), dx.data_ptr(), dO.data_ptr(), 0, 0,
                                         bnf, st)
            else:
                # identity shortcut: + dO * (out > 0), fused into the dgrad epilogue
                call.svk_conv2d_dgrad(d1, dc1.data_ptr(), b.conv1.w_dgrad.data_ptr(), dx.data_ptr(), 0, dO.data_ptr(),
                                      out.data_ptr(), st)
            reduced = fuse and below is not None
            gbuf[0], gbuf[2] = gbuf[2], gbuf[0]
            dO = dx
        # ---- stem
        if cur_bucket != 2:
            self._bucket_done(cur_bucket)
            cur_bucket = 2
        F, T, C0 = sv["F"], sv["T"], self.c0
        M = B * F * T
        c0, a0 = ws["c0"], ws["a0"]
        bn = self.stem_bn
        _, _, mu, rs = self._coefs(bn)
        sums = self._bsums[bn.idx]
        dc0 = gbuf[1][:c0.numel()].view(c0.shape)
        self._wait_side("dc2")           # dc0 aliases the buffer the first block's conv2 weight gradient reads on the side stream
        amask = 0 if masked else a0.data_ptr()
        if not reduced:
            call.svk_bn_bwd_reduce(dO.data_ptr(), amask, c0.data_ptr(), mu.data_ptr(), rs.data_ptr(), 0, 0, 0,
                                   sums.data_ptr(), M, C0, self.dcode, st)
        call.svk_bn_bwd_apply(dO.data_ptr(), amask, c0.data_ptr(), mu.data_ptr(), rs.data_ptr(),
                              bn.mod.weight.data_ptr(), dc0.data_ptr(), 0, 0, 0, 0, 0, sums.data_ptr(),
                              self._gview[id(bn.mod.weight)].data_ptr(), self._gview[id(bn.mod.bias)].data_ptr(), 0, 0,
                              M, C0, self.dcode, st)
        if self.debug is not None:
            self.debug["res.conv1"] = dc0.clone()
        call.svk_stem_conv_wgrad(sv["x"].data_ptr(), dc0.data_ptr(), self._gview[id(self.stem_conv.weight)].data_ptr(),
                                 B, F, T, C0, self.dcode, st)
        self._bucket_done(2, last=True)
        # publish gradients: p.grad is a view of the flat gradient buffer (overwrite semantics)
        for p, g in zip(self._params, self._grad_views):
            if p.grad is not g:
                p.grad = g
        if keep is not None:
            self._grad.add_(keep)
        self._grads_fresh = True
        lease = sv.pop("lease", None)
        if lease is not None:
            lease.release()

    def grads_consumed(self):
        """SGD.step() / zero_grad(): the next backward overwrites the gradient buffer instead of accumulating into it."""
        self._grads_fresh = False

    def _wgrad(self, d, conv, x, dy, tag=None):
        """Weight gradient of one convolution.  With the side stream, `tag` names the gradient buffer `dy` lives in: the main
        stream calls _wait_side(tag) before it overwrites that buffer."""
        need = lib.load().svk_conv2d_wgrad_workspace_bytes(d)
        if need == 0:
            raise lib.SvkError("svk_conv2d_wgrad_workspace_bytes failed: " + lib.load().svk_last_error_string().decode())
        old = getattr(self, "_arena_bufs", {}).get("wgrad_ws")
        if old is not None and old.numel() < (need + 3) // 4 and self._side_stream is not None:
            torch.cuda.current_stream().wait_stream(self._side_stream)     # the workspace is about to be replaced: let its readers finish
        ws = self._arena("wgrad_ws", ((need + 3) // 4,), torch.float32)
        if not self.wgrad_side or tag is None:
            call.svk_conv2d_wgrad(d, x.data_ptr(), dy.data_ptr(), self._gview[id(conv.mod.weight)].data_ptr(), ws.data_ptr(),
                                  ws.numel() * 4, _stream())
            return
        main = torch.cuda.current_stream()
        if self._side_stream is None or self._side_stream.device != self.device:
            self._side_stream = torch.cuda.Stream(device=self.device)
        side = self._side_stream
        ready = torch.cuda.Event()
        ready.record(main)                       # dy (written by the kernel just launched on the main stream) is complete
        side.wait_event(ready)
        with torch.cuda.stream(side):
            if self._side_delay:
                torch.cuda._sleep(self._side_delay)
            call.svk_conv2d_wgrad(d, x.data_ptr(), dy.data_ptr(), self._gview[id(conv.mod.weight)].data_ptr(), ws.data_ptr(),
                                  ws.numel() * 4, _stream())
            done = torch.cuda.Event()
            done.record(side)
        self._side_pending[tag] = done

    def _wait_side(self, tag):
        """The main stream is about to overwrite gradient buffer `tag`: wait for the side-stream kernel that reads it."""
        ev = self._side_pending.pop(tag, None)
        if ev is not None:
            torch.cuda.current_stream().wait_event(ev)

    def _join_side(self):
        """Everything issued on the side stream so far becomes visible to the main stream (gradient buckets, optimizer)."""
        if self._side_stream is not None and self._side_pending:
            torch.cuda.current_stream().wait_stream(self._side_stream)
            self._side_pending.clear()

    def _bucket_done(self, idx, last=False):
        """The gradients of bucket `idx` have been issued.  Mid-backward the main stream does NOT wait for the side stream:
        the gradient all-reduce (svk/parallel.py) waits for `bucket_side_event` on its own stream instead; the last bucket
        joins the streams (the optimizer follows)."""
        self.bucket_side_event = None
        if last:
            self._join_side()
        elif self._side_stream is not None and self._side_pending and self.grad_ready_cb is not None:
            self.bucket_side_event = torch.cuda.Event()
            self.bucket_side_event.record(self._side_stream)
        if self.grad_ready_cb is not None:
            self.grad_ready_cb(idx)

    # ------------------------------------------------------------------------------------------ eval / extraction
    def _eval_coefs(self, bn):
        sc, sh, _, _ = self._coefs_eval[bn.idx]
        return sc, sh

    def _prepare_eval(self):
        if self._eval_ready:
            return
        self._pack_weights()
        st = _stream()
        if not hasattr(self, "_coef_eval_buf") or self._coef_eval_buf.device != self.device:
            self._coef_eval_buf = torch.zeros(len(self.bns), 4, self._cmax, dtype=torch.float32, device=self.device)
        self._coefs_eval = [self._coef_eval_buf[i] for i in range(len(self.bns))]
        for bn in self.bns:
            m = bn.mod
            sc, sh = self._coefs_eval[bn.idx][0], self._coefs_eval[bn.idx][1]
            call.svk_bn_eval_coeffs(m.weight.data_ptr(), m.bias.data_ptr(), m.running_mean.data_ptr(),
                                    m.running_var.data_ptr(), _EPS, bn.C, sc.data_ptr(), sh.data_ptr(), st)
        self._eval_ready = True

    def forward_eval(self, x, y=None, lengths=None, with_head=False):
        """Eval-mode network (BatchNorm folded into the conv epilogues).  x: (B, F, T) fp32, zero-padded past
        `lengths` (int32 CUDA tensor of valid frame counts) when utterances of different length are batched; every
        layer re-zeroes the padding so each row equals its batch-1 result (decode.py:198 semantics)."""
        self.ensure_device()
        self._prepare_eval()
        model = self.model
        B, F, T = x.shape
        x = x.contiguous().float()
        ws = self._workspace(("evalhead", B))
        st = _stream()
        valid = None
        if lengths is not None:
            valid = [lengths.to(torch.int32).contiguous()]
            for _ in range(3):
                valid.append(((valid[-1] + 1) // 2).to(torch.int32))
        C0 = self.c0
        a0 = self._arena("e_a0", (B, F, T, C0))
        sc, sh = self._eval_coefs(self.stem_bn)
        call.svk_stem_conv_fwd(x.data_ptr(), self.stem_conv.weight.data_ptr(), a0.data_ptr(), B, F, T, C0, self.dcode,
                               sc.data_ptr(), sh.data_ptr(), 1, _ptr(valid[0]) if valid else 0, 0, st)
        cur, H, W = a0, F, T
        stage = 0
        for bi, b in enumerate(self.blocks):
            d1 = self._desc(B, H, W, b.conv1)
            Ho, Wo = d1.Ho, d1.Wo
            Co = b.conv1.cout
            if b.conv1.stride == 2:
                stage += 1
            v = valid[stage] if valid else None
            a1 = self._arena("e_a1", (B, Ho, Wo, Co))
            out = self._arena("e_o%d" % (bi % 2), (B, Ho, Wo, Co))
            sc1, sh1 = self._eval_coefs(b.bn1)
            self._conv_fwd(d1, b.conv1, cur, a1, scale=sc1, shift=sh1, relu=1, valid=v)
            res = cur
            if b.convd is not None:
                dd = self._desc(B, H, W, b.convd)
                cd = self._arena("e_cd", (B, Ho, Wo, Co))
                scd, shd = self._eval_coefs(b.bnd)
                self._conv_fwd(dd, b.convd, cur, cd, scale=scd, shift=shd, relu=0, valid=v)
                res = cd
            d2 = self._desc(B, Ho, Wo, b.conv2)
            sc2, sh2 = self._eval_coefs(b.bn2)
            self._conv_fwd(d2, b.conv2, a1, out, scale=sc2, shift=sh2, res=res, relu=1, valid=v)
            cur, H, W = out, Ho, Wo
        mode = 1 if model.pool.pooling == "mean+std" else 0
        Cl = self.c_last
        pdim = Cl * H * (2 if mode else 1)
        pooled = self._arena("e_pool", (B, pdim), torch.float32)
        call.svk_statspool_fwd(cur.data_ptr(), pooled.data_ptr(), B, H, W, Cl, mode, _ptr(valid[3]) if valid else 0,
                               self.dcode, st)
        fc = model.fc1
        E = fc.weight.shape[0]
        emb = torch.empty(B, E, dtype=torch.float32, device=self.device)
        self._gemm(pooled, pdim, True, fc.weight, pdim, True, emb, E, B, E, pdim, fc.bias)
        if not with_head:
            return emb
        return self._head_fwd(emb, y, ws, None, train=False)


def c_dtype_code(t):
    if t.dtype == torch.bfloat16:
        return lib.BF16
    if t.dtype == torch.float32:
        return lib.F32
    raise lib.SvkError("unsupported activation dtype %s" % t.dtype)


class _Lease(object):
    """Marks a training workspace busy until its backward has run — or until autograd frees the graph without one."""
    __slots__ = ("ws",)

    def __init__(self, ws):
        self.ws = ws

    def release(self):
        if self.ws is not None:
            self.ws["_busy"] = False
            self.ws = None

    def __del__(self):
        self.release()


def run_train(engine, x, y):
    """Training forward as the registered custom op `svk::speaker_net_train` (svk/ops.py): autograd records ONE node whose
    backward is `svk::speaker_net_train_backward`."""
    from . import ops
    return ops.speaker_net_train(engine, x, y)
