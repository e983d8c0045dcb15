"""Agents of the replay path with the reference surface `cls(model, opt, params)`,
`.train_learner(x_train, y_train)`, `.evaluate(test_loaders)`:
ExperienceReplay (agents/exp_replay.py:10-104: ER / MIR / ASER) and SupContrastReplay
(agents/scr.py:11-69), over ContinualLearner (agents/base.py:14-113).

The loop structure, the order of train-mode forwards (they move BN running statistics) and the
order of buffer operations follow the reference step exactly; what changes is who does the
arithmetic (the CUDA engine, not autograd) and that dead work is not executed: in the ASER
branch the reference computes and then discards two backward passes (exp_replay.py:55,77,81) --
their forwards are kept for the BN side effect, their backwards are skipped.
"""
import numpy as np
import torch

from . import memory, ops
from .augment import SCRTransform
from .engine import ce_loss
from .memory import Buffer, input_size_match
from .nets import adopt, engine_of, EngineModel


# The train-mode passes of one replay step only interact through the BatchNorm running statistics and the gradient arena.
# With this switch on (default; B200OCL_CONCURRENT=0 turns it off) independent passes are issued on two streams -- SCR's two
# views (forward and backward), the memory / combined forwards of the ASER branch -- with deferred statistics applied in the
# reference's order and the second backward pass into a second gradient arena.  Measured on one B200: two N=110 passes
# 5.00 -> 3.91 ms, two N=10 forwards 0.76 -> 0.45 ms (tools/overlap_probe.py): most launches of this network fill a
# fraction of the GPU.
import os as _os
_CONCURRENT = _os.environ.get('B200OCL_CONCURRENT', '1') != '0'


def set_concurrent(on):
    global _CONCURRENT
    _CONCURRENT = bool(on)


class AverageMeter(object):
    """utils/utils.py:25-42, but values may stay on the device until avg() is read."""

    def __init__(self):
        self.reset()

    def reset(self):
        self.sum = 0
        self.count = 0

    def update(self, val, n):
        self.sum = self.sum + val * n
        self.count += n

    def avg(self):
        if self.count == 0:
            return 0
        return float(self.sum) / self.count


class StreamFeeder(object):
    """One task's stream: uint8 NHWC images -> fp32 NCHW in [0,1] on the device (ToTensor,
    continuum/data_utils.py:38-54), shuffled, batches of `batch`, last partial batch dropped
    (exp_replay.py:21-23).  The whole task is converted once; batches are views."""

    def __init__(self, x_train, y_train, batch, device):
        self.batch = batch
        y = np.asarray(y_train).astype(np.int64)
        if memory.parity():
            # the reference's DataLoader(shuffle=True, drop_last=True) itself, over slot numbers: consumes the
            # default CPU generator exactly as exp_replay.py:21-23 / scr.py:29-31 do (base seed + sampler seed)
            from torch.utils.data import DataLoader, TensorDataset
            order = [b[0] for b in DataLoader(TensorDataset(torch.arange(len(y))), batch_size=batch, shuffle=True,
                                              num_workers=0, drop_last=True)]
            perm = torch.cat(order).numpy() if order else np.zeros(0, dtype=np.int64)
        else:
            perm = torch.randperm(len(y)).numpy()        # DataLoader(shuffle=True) draws from the torch CPU generator
        x = torch.from_numpy(np.ascontiguousarray(np.asarray(x_train)))
        if x.dtype == torch.uint8 and torch.device(device).type == 'cuda':
            # one upload of the raw bytes, then shuffle + HWC->CHW + /255 in one kernel (csrc/misc.cu)
            x = ops.stream_prepare(x.to(device), torch.from_numpy(perm).to(device))
        elif x.dtype == torch.uint8:
            x = x[torch.from_numpy(perm)].permute(0, 3, 1, 2).to(torch.float32).div_(255.0).contiguous()
        else:                                             # already float NCHW
            x = x[torch.from_numpy(perm)].to(device=device, dtype=torch.float32).contiguous()
        self.x = x
        self.y_host = y[perm]
        self.y = torch.from_numpy(self.y_host).to(device)

    def __len__(self):
        return len(self.y_host) // self.batch

    def __iter__(self):
        b = self.batch
        for i in range(len(self)):
            yield self.x[i * b:(i + 1) * b], self.y[i * b:(i + 1) * b], self.y_host[i * b:(i + 1) * b]


class ContinualLearner(torch.nn.Module):
    """Label bookkeeping and loss dispatch of agents/base.py:14-113 for the replay path."""

    def __init__(self, model, opt, params):
        super().__init__()
        self.params = params
        self.model = model
        self.opt = opt
        self.data = params.data
        self.cuda = params.cuda
        self.epoch = params.epoch
        self.batch = params.batch
        self.verbose = params.verbose
        self.old_labels = []
        self.new_labels = []
        self.task_seen = 0
        self.lbl_inv_map = {}
        self.class_task_map = {}
        trick = getattr(params, 'trick', None) or {}
        unsupported = [k for k in ('labels_trick', 'kd_trick', 'separated_softmax', 'kd_trick_star') if trick.get(k)]
        if unsupported:
            raise NotImplementedError('tricks %s are outside the replay-path scope (SURVEY section 8f)' % unsupported)
        if isinstance(model, EngineModel):
            self.engine = model.engine
        else:
            self.engine = adopt(model, input_size_match[params.data][1])
        self.device = self.engine.device
        self.grad_sync = None      # data-parallel stream shards: callable(engine) summing gradients over ranks
        self.grad_world = 1

    def _fork(self):
        """(main, side) streams with side ordered after everything issued on main so far."""
        main = torch.cuda.current_stream()
        side = self.__dict__.get('_side_stream')
        if side is None:
            side = self.__dict__['_side_stream'] = torch.cuda.Stream(device=self.device)
        side.wait_stream(main)
        return main, side

    def _throttle(self, depth=2):
        """Keep the host at most `depth` replay steps ahead of the GPU: the stream never runs dry, the launch
        queue stays short, and a CUDA-graph executable is never relaunched while it is still running."""
        if not torch.cuda.is_available():
            return
        ring = self.__dict__.setdefault('_step_events', [])
        ev = torch.cuda.Event()
        ev.record()
        ring.append(ev)
        if len(ring) > depth:
            ring.pop(0).synchronize()

    def _lr_wd(self):
        """Step size / weight decay from the optimizer the caller built (run.py:40).  Only plain
        SGD is implemented (setup_elements.py:73-75, the reference default)."""
        opt = self.opt
        if opt is None:
            return float(self.params.learning_rate), float(getattr(self.params, 'weight_decay', 0.0))
        if not isinstance(opt, torch.optim.SGD):
            raise NotImplementedError('the b200ocl engine implements torch.optim.SGD only')
        g = opt.param_groups[0]
        if g.get('momentum', 0) or g.get('nesterov', False) or g.get('dampening', 0):
            raise NotImplementedError('SGD momentum/nesterov are not used by the reference and not implemented')
        return float(g['lr']), float(g['weight_decay'])

    def _optimizer_step(self, lr, wd):
        """opt.step(); with data-parallel stream shards the summed gradient is averaged first
        (one all-reduce of the flat gradient arena, folded into the step size)."""
        if self.grad_sync is not None:
            self.grad_sync(self.engine)
            if wd != 0.0:
                raise NotImplementedError('weight decay with gradient averaging')
            lr = lr / self.grad_world
        self.engine.sgd_step(lr, wd)

    def before_train(self, x_train, y_train):
        new_labels = list(set(np.asarray(y_train).tolist()))
        self.new_labels += new_labels
        for i, lbl in enumerate(new_labels):
            self.lbl_inv_map[lbl] = len(self.old_labels) + i
        for i in new_labels:
            self.class_task_map[i] = self.task_seen

    def after_train(self):
        self.old_labels += self.new_labels
        self.new_labels_zombie = list(self.new_labels)
        self.new_labels.clear()
        self.task_seen += 1
        if (getattr(self.params, 'trick', None) or {}).get('review_trick') and hasattr(self, 'buffer'):
            self._review()

    def _review(self):
        """Review trick (agents/base.py:62-88, the published SCR setting config_CVPR/agent/scr/scr_5k.yml:10): one
        pass over the filled memory in shuffled batches of eps_mem_batch (drop_last), gradients divided by 10.
        g/10 followed by SGD(lr, wd) is p -= lr*(g/10 + wd*p) = SGD(lr/10, 10*wd) on g: folded into the step."""
        eng = self.engine
        lr, wd = self._lr_wd()
        n = self.buffer.current_index
        bs = self.params.eps_mem_batch
        if n == 0 or n < bs:
            return
        self.model.train()
        if memory.parity():
            from torch.utils.data import DataLoader, TensorDataset
            batches = [b[0] for b in DataLoader(TensorDataset(torch.arange(n)), batch_size=bs, shuffle=True, num_workers=0,
                                                drop_last=True)]
        else:
            perm = torch.randperm(n)
            batches = [perm[i * bs:(i + 1) * bs] for i in range(n // bs)]
        scr = self.params.agent == 'SCR'
        for idx in batches:
            idx_t = memory.to_device_i64(idx.numpy(), self.device)
            bx = ops.gather_rows(self.buffer.buffer_img, idx_t)
            by = ops.gather_rows(self.buffer.buffer_label, idx_t)
            out, ws = eng.forward_train(bx, slot=0)                                  # base.py:76 (for SCR: BN side effect only)
            if scr:
                f1, ws1 = eng.forward_train(bx, slot=0)                              # base.py:78-79
                aug = self.transform(bx)
                f2, ws2 = eng.forward_train(aug, slot=1)
                loss, dfeat = ops.supcon(torch.stack((f1, f2), dim=1), by, self.params.temp)
                eng.backward(bx, dfeat[:, 0].contiguous(), ws1)
                eng.backward(aug, dfeat[:, 1].contiguous(), ws2, accumulate=True)
            else:
                ce = ce_loss(out, by, want_grad=True)
                eng.backward(bx, ce['dlogits'], ws)
            self._optimizer_step(lr / 10.0, wd * 10.0)                               # base.py:83-87
            self._throttle()

    def train_learner(self, x_train, y_train):
        raise NotImplementedError

    def forward(self, x):
        return self.model.forward(x)

    # ------------------------------------------------------------------ evaluate (next)
    def _ncm(self):
        return (getattr(self.params, 'trick', None) or {}).get('ncm_trick') or self.params.agent in ['ICARL', 'SCR', 'SCP']

    @torch.no_grad()
    def evaluate(self, test_loaders):
        """Accuracy per task (agents/base.py:118-227): nearest-class-mean over buffer features for SCR /
        ncm_trick, arg-max of the classifier otherwise.  Encoder features come from the engine's batched
        eval pass (the reference runs model.features once per buffered image, base.py:125-134); class
        means, nearest-mean / arg-max and the hit count are the kernels of csrc/ncm.cu; one device -> host
        read per test loader.  error_analysis is not implemented."""
        if getattr(self.params, 'error_analysis', False):
            raise NotImplementedError('error_analysis is outside the replay-path scope')
        eng = self.engine
        eng.pack()                  # the caller may have written the Parameters since the last step
        self.model.eval()           # base.py:119 (the next train_learner switches back)
        acc_array = np.zeros(len(test_loaders))
        ncm = self._ncm()
        if ncm:
            n = self.buffer.current_index
            class_ids = torch.tensor(self.old_labels, dtype=torch.int64, device=self.device)
            feats = torch.cat([eng.features_eval(self.buffer.buffer_img[s:s + 500]) for s in range(0, n, 500)]) \
                if n else torch.zeros((0, eng.dim_in), device=self.device)
            if n and not bool(torch.isin(self.buffer.buffer_label[:n], class_ids).all()):
                raise KeyError('a buffered label was never seen in training (the reference raises here, base.py:126)')
            means, counts = ops.ncm_class_means(feats, self.buffer.buffer_label[:n], class_ids)
            empty = (counts == 0).nonzero().flatten().tolist()
            for k in empty:         # base.py:135-137: a random direction for a class without exemplars
                mu = torch.normal(0, 1, size=(1, eng.dim_in)).to(self.device).squeeze()
                means[k] = mu / mu.norm()
        else:
            if isinstance(self.model, EngineModel):
                W, b = self.model.linear__weight, self.model.linear__bias
            else:
                W, b = self.model.linear.weight, self.model.linear.bias
        for task, loader in enumerate(test_loaders):
            hits = torch.zeros(1, dtype=torch.int64, device=self.device)
            total = 0
            for batch_x, batch_y in loader:
                batch_x, batch_y = batch_x.to(self.device), batch_y.to(self.device)
                f = eng.features_eval(batch_x)
                if ncm:
                    ops.ncm_classify(f, means, class_ids, truth=batch_y, n_correct=hits)
                else:
                    ops.linear_argmax(f, W, b, truth=batch_y, n_correct=hits)
                total += batch_y.numel()
            acc_array[task] = int(hits) / max(total, 1)
        print(acc_array)
        return acc_array


class ExperienceReplay(ContinualLearner):
    def __init__(self, model, opt, params):
        super().__init__(model, opt, params)
        self.buffer = Buffer(model, params)
        self.mem_size = params.mem_size
        self.eps_mem_batch = params.eps_mem_batch
        self.mem_iters = params.mem_iters
        self._aser_branch = params.update == 'ASER' or params.retrieve == 'ASER'
        self._needs_batch_grad = params.retrieve == 'MIR' or not self._aser_branch

    def replay_step(self, batch_x, batch_y, batch_y_host, meters=None):
        """One iteration of exp_replay.py:34-92."""
        eng = self.engine
        lr, wd = self._lr_wd()
        aser = self._aser_branch
        for _ in range(self.mem_iters):
            logits, ws = eng.forward_train(batch_x, slot=0)                         # :40  (BN stats move)
            ce = ce_loss(logits, batch_y, want_grad=self._needs_batch_grad, want_correct=meters is not None)
            if meters is not None:
                meters['acc_batch'].update(ce['n_correct'] / batch_y.size(0), batch_y.size(0))
                meters['losses_batch'].update(ce['loss'], batch_y.size(0))
            if self._needs_batch_grad:
                eng.backward(batch_x, ce['dlogits'], ws)                            # :54-55
            mem_x, mem_y = self.buffer.retrieve(x=batch_x, y=batch_y)               # :58
            both = _CONCURRENT and aser and mem_x.size(0) > 0
            if both:
                # ASER branch: the memory forward (:62, kept for its BN side effect and the meters) and the forward of the
                # concatenated batch (:84) are independent -- two streams, statistics applied in call order
                main, side = self._fork()
                with torch.cuda.stream(side):
                    mem_logits, ws_m = eng.forward_train(mem_x, slot=1, defer_stats=True)
                    if meters is not None:
                        ce_m = ce_loss(mem_logits, mem_y, want_grad=False, want_correct=True)
            elif mem_x.size(0) > 0:
                mem_logits, ws_m = eng.forward_train(mem_x, slot=1)                 # :62  (BN stats move)
                ce_m = ce_loss(mem_logits, mem_y, want_grad=not aser, want_correct=meters is not None)
            if mem_x.size(0) > 0 and not both:
                if meters is not None:
                    meters['losses_mem'].update(ce_m['loss'], mem_y.size(0))
                    meters['acc_mem'].update(ce_m['n_correct'] / mem_y.size(0), mem_y.size(0))
                if not aser:
                    eng.backward(mem_x, ce_m['dlogits'], ws_m, accumulate=True)     # :77 (gradients accumulate)
            if aser:
                combined = torch.cat((mem_x, batch_x))                              # :82-83
                labels = torch.cat((mem_y, batch_y))
                logits_c, ws_c = eng.forward_train(combined, slot=3, defer_stats=both)   # :84
                if both:
                    main.wait_stream(side)
                    eng.apply_running_stats(ws_m, mem_x.size(0))
                    eng.apply_running_stats(ws_c, combined.size(0))
                    if meters is not None:
                        meters['losses_mem'].update(ce_m['loss'], mem_y.size(0))
                        meters['acc_mem'].update(ce_m['n_correct'] / mem_y.size(0), mem_y.size(0))
                ce_c = ce_loss(logits_c, labels, want_grad=True)
                eng.backward(combined, ce_c['dlogits'], ws_c)                       # :86
                self.last_loss = ce_c['loss']
            else:
                self.last_loss = ce['loss']
            self._optimizer_step(lr, wd)                                            # :87 / :89
        # (Measured and dropped: issuing this update next to the FOLLOWING iteration's first forward on a second stream.
        # The update's eval-feature pass runs persistent one-CTA-per-SM kernels, the forward cannot share the SMs with
        # them: 1.40 ms for the pair against 0.96 + 0.41 ms one after the other, and the host then waits for the update's
        # decision at the next retrieval.)
        self.buffer.update(batch_x, batch_y, y_host=batch_y_host)                   # :92
        self._throttle()

    def train_learner(self, x_train, y_train):
        self.before_train(x_train, y_train)
        self.engine.pack()          # the caller may have written the Parameters (load_state_dict, weight surgery)
        self.model = self.model.train()
        meters = {k: AverageMeter() for k in ('losses_batch', 'losses_mem', 'acc_batch', 'acc_mem')}
        for ep in range(self.epoch):
            stream = StreamFeeder(x_train, y_train, self.batch, self.device)   # DataLoader(shuffle=True): new order per epoch
            for i, (batch_x, batch_y, y_host) in enumerate(stream):
                self.replay_step(batch_x, batch_y, y_host, meters if self.verbose else None)
                if i % 100 == 1 and self.verbose:
                    print('==>>> it: {}, avg. loss: {:.6f}, running train acc: {:.3f}'
                          .format(i, meters['losses_batch'].avg(), meters['acc_batch'].avg()))
                    print('==>>> it: {}, mem avg. loss: {:.6f}, running mem acc: {:.3f}'
                          .format(i, meters['losses_mem'].avg(), meters['acc_mem'].avg()))
        self.after_train()


class SupContrastReplay(ContinualLearner):
    def __init__(self, model, opt, params):
        super().__init__(model, opt, params)
        self.buffer = Buffer(model, params)
        self.mem_size = params.mem_size
        self.eps_mem_batch = params.eps_mem_batch
        self.mem_iters = params.mem_iters
        hw = input_size_match[params.data]
        self.transform = SCRTransform(size=(hw[1], hw[2]))        # scr.py:18-24

    def replay_step(self, batch_x, batch_y, batch_y_host, meters=None):
        """One iteration of scr.py:40-63."""
        eng = self.engine
        lr, wd = self._lr_wd()
        for _ in range(self.mem_iters):
            mem_x, mem_y = self.buffer.retrieve(x=batch_x, y=batch_y)               # :47
            if mem_x.size(0) > 0:                                                   # :49 (no training on an empty buffer)
                combined = torch.cat((mem_x, batch_x))                              # :52-53
                labels = torch.cat((mem_y, batch_y))
                combined_aug = self.transform(combined)                             # :54
                if _CONCURRENT:
                    n = combined.size(0)
                    main, side = self._fork()
                    with torch.cuda.stream(side):                                   # :55 two train-mode forwards, side by side
                        f2, ws2 = eng.forward_train(combined_aug, slot=1, defer_stats=True)
                    f1, ws1 = eng.forward_train(combined, slot=0, defer_stats=True)
                    main.wait_stream(side)
                    eng.apply_running_stats(ws1, n)                                 # the running statistics move in call order
                    eng.apply_running_stats(ws2, n)
                    feats = torch.stack((f1, f2), dim=1)
                    loss, dfeat = ops.supcon(feats, labels, self.params.temp)       # :56  (base.py:109-111)
                    d1, d2 = dfeat[:, 0].contiguous(), dfeat[:, 1].contiguous()
                    main, side = self._fork()
                    with torch.cuda.stream(side):                                   # :58-59 one backward per view
                        eng.backward(combined_aug, d2, ws2, alt=True)
                    eng.backward(combined, d1, ws1)
                    main.wait_stream(side)
                    eng.add_alt_grads()
                else:
                    f1, ws1 = eng.forward_train(combined, slot=0)                   # :55 two train-mode forwards
                    f2, ws2 = eng.forward_train(combined_aug, slot=1)
                    feats = torch.stack((f1, f2), dim=1)
                    loss, dfeat = ops.supcon(feats, labels, self.params.temp)       # :56  (base.py:109-111)
                    eng.backward(combined, dfeat[:, 0].contiguous(), ws1)           # :58-59
                    eng.backward(combined_aug, dfeat[:, 1].contiguous(), ws2, accumulate=True)
                self._optimizer_step(lr, wd)                                        # :60
                self.last_loss = loss
                if meters is not None:
                    meters['losses'].update(loss, batch_y.size(0))
        self.buffer.update(batch_x, batch_y, y_host=batch_y_host)                   # :63
        self._throttle()

    def train_learner(self, x_train, y_train):
        self.before_train(x_train, y_train)
        self.engine.pack()          # the caller may have written the Parameters (load_state_dict, weight surgery)
        self.model = self.model.train()
        meters = {'losses': AverageMeter()}
        for ep in range(self.epoch):
            stream = StreamFeeder(x_train, y_train, self.batch, self.device)   # DataLoader(shuffle=True): new order per epoch
            for i, (batch_x, batch_y, y_host) in enumerate(stream):
                self.replay_step(batch_x, batch_y, y_host, meters if self.verbose else None)
                if i % 100 == 1 and self.verbose:
                    print('==>>> it: {}, avg. loss: {:.6f}, '.format(i, meters['losses'].avg()))
        self.after_train()


class AGEM(ContinualLearner):
    """Averaged GEM (agents/agem.py:11-90) on the engine: the gradient of the stream batch is projected against the
    gradient of a memory batch of earlier tasks when their inner product is negative -- two train-mode
    forward / backward passes over the flat gradient arena and one projection kernel (no per-parameter Python loops)."""

    def __init__(self, model, opt, params):
        super().__init__(model, opt, params)
        self.buffer = Buffer(model, params)
        self.mem_size = params.mem_size
        self.eps_mem_batch = params.eps_mem_batch
        self.mem_iters = params.mem_iters
        self._g_cur = torch.empty_like(self.engine.state.grads)

    def replay_step(self, batch_x, batch_y, batch_y_host, meters=None):
        """One iteration of agem.py:36-84."""
        eng = self.engine
        lr, wd = self._lr_wd()
        for _ in range(self.mem_iters):
            logits, ws = eng.forward_train(batch_x, slot=0)                          # :39
            ce = ce_loss(logits, batch_y, want_grad=True, want_correct=meters is not None)
            if meters is not None:
                meters['acc_batch'].update(ce['n_correct'] / batch_y.size(0), batch_y.size(0))
                meters['losses_batch'].update(ce['loss'], batch_y.size(0))
            eng.backward(batch_x, ce['dlogits'], ws)                                 # :53-54
            self.last_loss = ce['loss']
            if self.task_seen > 0:
                mem_x, mem_y = self.buffer.retrieve()                                # :58 (no kwargs)
                if mem_x.size(0) > 0:
                    self._g_cur.copy_(eng.state.grads)                               # :62 grad of the current batch
                    mem_logits, ws_m = eng.forward_train(mem_x, slot=1)              # :65
                    ce_m = ce_loss(mem_logits, mem_y, want_grad=True)
                    eng.backward(mem_x, ce_m['dlogits'], ws_m)                       # :67-68 -> grad_ref in the arena
                    ops.agem_project(self._g_cur, eng.state.grads, out=eng.state.grads)   # :73-80
            self._optimizer_step(lr, wd)                                             # :81
        self.buffer.update(batch_x, batch_y, y_host=batch_y_host)                    # :83
        self._throttle()

    def train_learner(self, x_train, y_train):
        self.before_train(x_train, y_train)
        self.engine.pack()
        self.model = self.model.train()
        meters = {k: AverageMeter() for k in ('losses_batch', 'acc_batch')}
        for ep in range(self.epoch):
            stream = StreamFeeder(x_train, y_train, self.batch, self.device)
            for i, (batch_x, batch_y, y_host) in enumerate(stream):
                self.replay_step(batch_x, batch_y, y_host, meters if self.verbose else None)
                if i % 100 == 1 and self.verbose:
                    print('==>>> it: {}, avg. loss: {:.6f}, running train acc: {:.3f}'
                          .format(i, meters['losses_batch'].avg(), meters['acc_batch'].avg()))
        self.after_train()
