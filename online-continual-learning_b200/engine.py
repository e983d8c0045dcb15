"""Host handle of the Reduced-ResNet18 / SupConResNet engine (C ABI: b200ocl_net_*).

Owns the flat device arenas (parameters, gradients, packed weights, BN running
statistics) as torch tensors and hands raw pointers to libb200ocl.so.  The layer
plan, arena layout and every kernel live on the C side (csrc/net_*.cu*, conv.cu);
this file is plumbing.  Mirrors what the reference gets from nn.Module +
torch.optim.SGD (models/resnet.py, utils/setup_elements.py:46-82).
"""
import ctypes
from ctypes import c_int, c_int64, c_size_t, c_void_p

import torch

from . import _native
from .ops import _stream, _workspace, _need_cuda

HEAD_CODES = {None: 0, 'classifier': 0, 'linear': 1, 'mlp': 2, 'None': 3}


class NetDesc(ctypes.Structure):
    _fields_ = [('in_h', c_int), ('in_w', c_int), ('nf', c_int), ('num_classes', c_int), ('head', c_int),
                ('feat_dim', c_int)]


class NetState(ctypes.Structure):
    _fields_ = [('params', c_void_p), ('grads', c_void_p), ('packed', c_void_p), ('bn_stats', c_void_p),
                ('bn_tracked', c_void_p)]


class NetInfo(ctypes.Structure):
    _fields_ = [('n_params', c_size_t), ('n_packed', c_size_t), ('n_bn_stats', c_size_t), ('n_bn', c_int),
                ('n_tensors', c_int), ('dim_in', c_int), ('out_dim', c_int)]


def _lib():
    return _native.lib()


def describe(in_hw, num_classes, head=None, feat_dim=128, nf=20):
    """Host-only: arena sizes and tensor table for a network description (no GPU needed)."""
    desc = NetDesc(int(in_hw), int(in_hw), int(nf), int(num_classes), HEAD_CODES[head], int(feat_dim))
    info = NetInfo()
    lib = _lib()
    _native.check(lib.b200ocl_net_query(ctypes.byref(desc), ctypes.byref(info)), 'b200ocl_net_query')
    table = []
    off, num, hg = c_size_t(), c_size_t(), c_int()
    for i in range(info.n_tensors):
        _native.check(lib.b200ocl_net_tensor(ctypes.byref(desc), i, ctypes.byref(off), ctypes.byref(num),
                                             ctypes.byref(hg)), 'b200ocl_net_tensor')
        table.append((off.value, num.value, bool(hg.value)))
    return desc, info, table


class ArenaState:
    """One set of arenas (the live model, or MIR's virtual copy)."""

    def __init__(self, info, device, with_grads=True):
        self.params = torch.zeros(info.n_params, dtype=torch.float32, device=device)
        self.grads = torch.zeros(info.n_params, dtype=torch.float32, device=device) if with_grads else None
        self.packed = torch.zeros(info.n_packed, dtype=torch.float32, device=device)
        self.bn_stats = torch.zeros(info.n_bn_stats, dtype=torch.float32, device=device)
        self.bn_tracked = torch.zeros(info.n_bn, dtype=torch.int64, device=device)
        self.c = NetState(self.params.data_ptr(), self.grads.data_ptr() if with_grads else None,
                          self.packed.data_ptr(), self.bn_stats.data_ptr(), self.bn_tracked.data_ptr())


# CUDA graphs for the network-level calls (40-120 kernel launches each): after two eager calls of a given
# (call, batch size, workspace) the launch sequence is captured once and replayed.  Replays read their inputs from
# static buffers (one device-to-device copy per call).  B200OCL_GRAPHS=0 or set_graphs(False) turns this off
# (the per-launch profiler needs eager launches).
import os as _os
_GRAPHS = _os.environ.get('B200OCL_GRAPHS', '1') != '0'


def set_graphs(on):
    global _GRAPHS
    _GRAPHS = bool(on)


_replayed = [0]      # kernel launches issued through graph replays (the library's own counter only sees eager ones)


def graph_launch_count():
    return _replayed[0]


class _Graphed:
    """One captured call: static inputs / outputs, the graphs, and the number of eager warm-up calls so far.
    The same launch sequence is captured into N_EXEC executable graphs used round-robin, so that a replay never
    has to wait for the previous launch of the same executable when the host runs several steps ahead."""
    __slots__ = ('inputs', 'outputs', 'graphs', 'calls', 'kernels', 'turn')
    N_EXEC = 3

    def __init__(self, inputs, outputs):
        self.inputs, self.outputs, self.graphs, self.calls, self.kernels, self.turn = inputs, outputs, [], 0, 0, 0

    def run(self, launch):
        if self.calls < 1:                       # eager once: lets every launcher configure its kernel
            launch()
            self.calls += 1
            return
        if not self.graphs:                      # second call: capture all executables, run the first
            for _ in range(self.N_EXEC):
                before = _native.launch_count()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    launch()
                self.kernels = int(_native.launch_count() - before)
                self.graphs.append(g)
        self.graphs[self.turn].replay()
        self.turn = (self.turn + 1) % self.N_EXEC
        _replayed[0] += self.kernels


class Engine:
    def __init__(self, in_hw, num_classes, head=None, feat_dim=128, device='cuda'):
        self.desc, self.info, self.table = describe(in_hw, num_classes, head, feat_dim)
        self.device = torch.device(device)
        if self.device.type != 'cuda':
            raise _native.NativeError('the b200ocl engine runs on CUDA only; there is no CPU fallback')
        self.head = head
        self.in_hw = in_hw
        self.dim_in, self.out_dim = self.info.dim_in, self.info.out_dim
        self.state = ArenaState(self.info, self.device)
        self._virtual = None
        self._eval_ws = {}
        self._train_ws = {}
        self._graphs = {}

    # ------------------------------------------------------------------ state
    def param_views(self, shapes=None):
        """Views of the parameter arena, one per tensor in parameters() order."""
        return [self.state.params[o:o + n] for (o, n, _) in self.table]

    def grad_views(self):
        return [self.state.grads[o:o + n] for (o, n, _) in self.table]

    def load(self, params, bn_state=None):
        """params: iterable of tensors in parameters() order (any device); bn_state: iterable of
        (running_mean, running_var[, num_batches_tracked]) per BatchNorm2d in module order."""
        params = list(params)
        if len(params) != len(self.table):
            raise ValueError('expected %d parameter tensors, got %d' % (len(self.table), len(params)))
        for (o, n, _), t in zip(self.table, params):
            if t.numel() != n:
                raise ValueError('parameter size mismatch: %d vs %d' % (t.numel(), n))
            self.state.params[o:o + n].copy_(t.detach().reshape(-1).to(torch.float32))
        if bn_state is not None:
            off = 0
            for i, entry in enumerate(bn_state):
                rm, rv = entry[0], entry[1]
                c = rm.numel()
                self.state.bn_stats[off:off + c].copy_(rm.detach().reshape(-1))
                self.state.bn_stats[off + c:off + 2 * c].copy_(rv.detach().reshape(-1))
                if len(entry) > 2:
                    self.state.bn_tracked[i] = int(entry[2])
                off += 2 * c
            if off != self.info.n_bn_stats:
                raise ValueError('BN statistics do not cover the network')
        self.pack()

    def bn_views(self):
        """[(running_mean, running_var)] views per BatchNorm2d in module order."""
        out, off = [], 0
        sizes = [n for (o, n, _) in self.table[1:3 * self.info.n_bn:3]]
        for c in sizes:
            out.append((self.state.bn_stats[off:off + c], self.state.bn_stats[off + c:off + 2 * c]))
            off += 2 * c
        return out

    def pack(self, state=None):
        st = state or self.state
        _native.check(_lib().b200ocl_net_pack(ctypes.byref(self.desc), ctypes.byref(st.c), _stream()),
                      'b200ocl_net_pack')

    def virtual_state(self):
        if self._virtual is None:
            self._virtual = ArenaState(self.info, self.device, with_grads=False)
        return self._virtual

    # ------------------------------------------------------------------ passes
    def _x(self, x):
        _need_cuda(x)
        x = x.detach().to(torch.float32).contiguous()
        if x.dim() != 4 or x.shape[1] != 3 or x.shape[2] != self.in_hw or x.shape[3] != self.in_hw:
            raise ValueError('expected images [N,3,%d,%d], got %s' % (self.in_hw, self.in_hw, tuple(x.shape)))
        return x

    def features_eval(self, x, state=None):
        """model.eval(); model.features(x) under no_grad -> [N, dim_in]."""
        x = self._x(x)
        n = x.shape[0]
        if n == 0:
            return torch.empty((0, self.dim_in), dtype=torch.float32, device=x.device)
        lib = _lib()
        ws = self._eval_ws.get(n)
        if ws is None:
            ws = _workspace(lib.b200ocl_net_eval_workspace_bytes(ctypes.byref(self.desc), n), x.device)
            if len(self._eval_ws) < 8:
                self._eval_ws[n] = ws
        st = state or self.state

        def launch(xin, feat):
            rc = lib.b200ocl_net_features_eval(ctypes.byref(self.desc), ctypes.byref(st.c), xin.data_ptr(), n,
                                               feat.data_ptr(), ws.data_ptr(), ws.numel(), _stream())
            _native.check(rc, 'b200ocl_net_features_eval')

        if _GRAPHS and state is None and self._eval_ws.get(n) is ws:
            key = ('eval', n)
            e = self._graphs.get(key)
            if e is None:
                e = self._graphs[key] = _Graphed([torch.empty_like(x)],
                                                 [torch.empty((n, self.dim_in), dtype=torch.float32, device=x.device)])
            e.inputs[0].copy_(x)
            e.run(lambda: launch(e.inputs[0], e.outputs[0]))
            return e.outputs[0].clone()          # the static buffer is overwritten by the next call
        feat = torch.empty((n, self.dim_in), dtype=torch.float32, device=x.device)
        launch(x, feat)
        return feat

    def new_train_workspace(self, n):
        return _workspace(_lib().b200ocl_net_train_workspace_bytes(ctypes.byref(self.desc), n), self.device)

    def train_workspace(self, n, slot=0):
        key = (n, slot)
        ws = self._train_ws.get(key)
        if ws is None:
            ws = self.new_train_workspace(n)
            if len(self._train_ws) < 8:
                self._train_ws[key] = ws
        return ws

    def forward_train(self, x, ws=None, state=None, slot=0, eval_stats=False, defer_stats=False):
        """model.train(); model.forward(x).  Returns (out [N,out_dim], workspace kept for backward).
        eval_stats=True: model.eval() forward that can be differentiated (running statistics, nothing updated).
        defer_stats=True: the BN running statistics are left alone; apply_running_stats(ws, N) moves them later (so that
        the train-mode passes of one step can run concurrently on different streams and still update the statistics in
        the reference's order)."""
        x = self._x(x)
        n = x.shape[0]
        graphed = _GRAPHS and ws is None and state is None and (n, slot) in self._train_ws and not eval_stats
        if ws is None:
            ws = self.train_workspace(n, slot)
        st = state or self.state
        name = ('b200ocl_net_forward_evalgrad' if eval_stats else
                'b200ocl_net_forward_train_deferred' if defer_stats else 'b200ocl_net_forward_train')

        def launch(xin, out):
            rc = getattr(_lib(), name)(ctypes.byref(self.desc), ctypes.byref(st.c), xin.data_ptr(), n, out.data_ptr(),
                                       ws.data_ptr(), ws.numel(), _stream())
            _native.check(rc, name)

        if graphed:
            key = ('fwd', n, slot, bool(defer_stats))
            e = self._graphs.get(key)
            if e is None:
                other = self._graphs.get(('fwd', n, slot, not defer_stats))     # one static input per (n, slot)
                e = self._graphs[key] = _Graphed([other.inputs[0] if other is not None else torch.empty_like(x)],
                                                 [torch.empty((n, self.out_dim), dtype=torch.float32, device=x.device)])
            e.inputs[0].copy_(x)
            e.run(lambda: launch(e.inputs[0], e.outputs[0]))
            return e.outputs[0].clone(), ws      # the static buffer is overwritten by the next call
        out = torch.empty((n, self.out_dim), dtype=torch.float32, device=x.device)
        launch(x, out)
        return out, ws

    def apply_running_stats(self, ws, n):
        """The running-statistics update of a forward_train(..., defer_stats=True) over n images kept in ws."""
        rc = _lib().b200ocl_net_apply_running_stats(ctypes.byref(self.desc), ctypes.byref(self.state.c), int(n), ws.data_ptr(),
                                                    ws.numel(), _stream())
        _native.check(rc, 'b200ocl_net_apply_running_stats')

    def alt_grads(self):
        """A second gradient arena (same layout): a backward pass that runs concurrently with another one writes here,
        add_alt_grads() folds it into the arena the optimizer reads."""
        if getattr(self, '_alt', None) is None:
            g = torch.zeros_like(self.state.grads)
            c = NetState(self.state.params.data_ptr(), g.data_ptr(), self.state.packed.data_ptr(),
                         self.state.bn_stats.data_ptr(), self.state.bn_tracked.data_ptr())
            self._alt = (g, c)
        return self._alt[0]

    def add_alt_grads(self):
        """grads += alt (one rounding per element: the same sum an accumulating backward pass forms)."""
        g, alt = self.state.grads, self.alt_grads()
        rc = _lib().b200ocl_sgd_step(g.data_ptr(), alt.data_ptr(), g.data_ptr(), g.numel(), -1.0, 0.0, _stream())
        _native.check(rc, 'b200ocl_sgd_step')

    def backward(self, x, dout, ws, accumulate=False, eval_stats=False, alt=False):
        """loss.backward() for the forward of the same x kept in ws; fills (or adds to) the grad arena.
        eval_stats=True: the forward was forward_train(..., eval_stats=True).  alt=True: into the second arena."""
        _need_cuda(dout)
        x = self._x(x)
        dout = dout.detach().to(torch.float32).contiguous()
        n = dout.shape[0]
        if x.shape[0] != n or dout.shape[1] != self.out_dim:
            raise ValueError('dout must be [N, out_dim] for the same N as x')
        if alt:
            self.alt_grads()
        st_c = self._alt[1] if alt else self.state.c

        def launch(xin, din):
            rc = _lib().b200ocl_net_backward(ctypes.byref(self.desc), ctypes.byref(st_c), xin.data_ptr(),
                                             din.data_ptr(), n, ws.data_ptr(), ws.numel(),
                                             (1 if accumulate else 0) | (2 if eval_stats else 0), _stream())
            _native.check(rc, 'b200ocl_net_backward')

        slot = next((k[1] for k, w in self._train_ws.items() if w is ws and k[0] == n), None) if (_GRAPHS and not eval_stats) else None
        fwd = (self._graphs.get(('fwd', n, slot, False)) or self._graphs.get(('fwd', n, slot, True))) if slot is not None else None
        if fwd is not None:
            # the images are the static copy the graphed forward read (same data as x)
            key = ('bwd', n, slot, bool(accumulate), bool(alt))
            e = self._graphs.get(key)
            if e is None:
                e = self._graphs[key] = _Graphed([fwd.inputs[0], torch.empty_like(dout)], [])
            e.inputs[1].copy_(dout)
            e.run(lambda: launch(e.inputs[0], e.inputs[1]))
            return
        launch(x, dout)

    def sgd_step(self, lr, weight_decay=0.0, dst=None):
        rc = _lib().b200ocl_net_sgd_step(ctypes.byref(self.desc), ctypes.byref(self.state.c), float(lr),
                                         float(weight_decay), ctypes.byref(dst.c) if dst is not None else None,
                                         _stream())
        _native.check(rc, 'b200ocl_net_sgd_step')


def ce_loss(logits, labels, want_grad=True, want_per_sample=False, want_correct=False):
    """Mean cross-entropy; returns dict(loss[1], dlogits, per_sample, n_correct[1])."""
    _need_cuda(logits, labels)
    logits = logits.detach().to(torch.float32).contiguous()
    labels = labels.detach().to(torch.int64).contiguous()
    n, c = logits.shape
    dev = logits.device
    out = {'loss': torch.empty(1, dtype=torch.float32, device=dev)}
    out['dlogits'] = torch.empty_like(logits) if want_grad else None
    out['per_sample'] = torch.empty(n, dtype=torch.float32, device=dev) if want_per_sample else None
    out['n_correct'] = torch.empty(1, dtype=torch.int64, device=dev) if want_correct else None
    ptr = lambda t: 0 if t is None else t.data_ptr()
    rc = _lib().b200ocl_ce_loss(logits.data_ptr(), labels.data_ptr(), n, c, out['loss'].data_ptr(),
                                ptr(out['per_sample']), ptr(out['dlogits']), ptr(out['n_correct']), _stream())
    _native.check(rc, 'b200ocl_ce_loss')
    return out
