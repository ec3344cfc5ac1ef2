"""Adam with the reference's semantics (train_model.py:49: torch.optim.Adam(lr=1e-3, weight_decay=1e-5), i.e. L2 added to
the gradient, bias-corrected moments) as ONE kernel launch for all parameters (uz_adam_step_batched) instead of torch's
multi-tensor chunks.  Same state layout as torch.optim.Adam(capturable=True) (``step`` device scalar, ``exp_avg``, ``exp_avg_sq`` per parameter), so
state_dicts interchange; the step counter is a device scalar, the launch is CUDA-graph capturable.

The caller's own torch.optim.Adam keeps working with the drop-in modules -- this class is what b200.train.make_adam
returns for the step drivers."""
import struct

import torch

from . import _lib


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self._tables = {}            # group index -> [eager, captured] descriptor / chunk tables (pinned host + device)
        self.packer = None           # kern.WeightPacker whose bf16 copies this optimizer keeps current (attach_packer)

    def attach_packer(self, packer):
        """From now on the conv weights known to ``packer`` are updated by the fused Adam + packing kernel
        (uz_adam_pack_step): their bf16 tensor-core copies are rewritten from the new values in the same pass, and the
        per-forward packing launch becomes a no-op (``packer.external``).  The copies are brought up to date once here;
        whoever changes parameters behind the optimizer's back afterwards must call ``packer.refresh(force=True)``."""
        self.packer = packer
        if packer is not None:
            packer.refresh(force=True)
            packer.external = True
        self._tables = {}            # re-created with tile tables on the next (eager) step
        self.__dict__['_state_gen'] = self.__dict__.get('_state_gen', 0) + 1

    def load_state_dict(self, state_dict):
        """torch's loader installs NEW tensors in ``self.state``; the descriptor tables (and a captured CUDA graph that
        replays the launch) hold raw pointers to the old ``exp_avg`` / ``exp_avg_sq`` / ``step`` tensors.  The loaded
        values are therefore copied INTO the existing state tensors where they exist (pointers stay valid, a captured
        step keeps working on the restored state); state of parameters that had none yet is new and the eager tables
        are rebuilt from the pointer key on the next step."""
        old = {p: dict(st) for p, st in self.state.items() if st}
        super().load_state_dict(state_dict)
        for p, st in list(self.state.items()):
            prev = old.get(p)
            if not prev or not st:
                continue
            for k in ('exp_avg', 'exp_avg_sq', 'step'):
                if k in st and k in prev and torch.is_tensor(prev[k]) and prev[k].is_cuda:
                    new = st[k] if torch.is_tensor(st[k]) else torch.as_tensor(float(st[k]))
                    prev[k].copy_(new.to(device=prev[k].device, dtype=prev[k].dtype).reshape(prev[k].shape))
                    st[k] = prev[k]
        for bufs in self._tables.values():
            bufs[0]['key'] = None              # eager tables: rebuilt on the next step (cheap); captured ones stay valid
        self.__dict__['_state_gen'] = self.__dict__.get('_state_gen', 0) + 1

    def reset_state(self):
        """zero the moments and step counters IN PLACE (pointers, hence captured graphs, stay valid)"""
        for st in self.state.values():
            for k in ('exp_avg', 'exp_avg_sq', 'step'):
                if k in st and torch.is_tensor(st[k]):
                    st[k].zero_()

    def _state(self, p):
        st = self.state[p]
        if not st:
            st['step'] = torch.zeros((), dtype=torch.float32, device=p.device)     # torch's capturable layout
            st['exp_avg'] = torch.zeros_like(p, memory_format=torch.preserve_format)
            st['exp_avg_sq'] = torch.zeros_like(p, memory_format=torch.preserve_format)
        elif not (torch.is_tensor(st['step']) and st['step'].is_cuda and st['step'].dtype == torch.float32):
            st['step'] = torch.as_tensor(float(st['step']), dtype=torch.float32, device=p.device)   # loaded state_dict
        return st

    def finish_capture(self):
        """upload the descriptor tables recorded while a CUDA graph was being captured (call after the capture ends)"""
        for bufs in self._tables.values():
            tab = bufs[1]
            if tab.get('dirty'):
                tab['descs'].copy_(tab['host_d'], non_blocking=True)
                tab['chunks'].copy_(tab['host_c'], non_blocking=True)
                if tab['ni']:
                    tab['packs'].copy_(tab['host_p'], non_blocking=True)
                    tab['items'].copy_(tab['host_i'], non_blocking=True)
                tab['dirty'] = False
        torch.cuda.current_stream().synchronize()

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for gi, group in enumerate(self.param_groups):
            live = [p for p in group['params'] if p.grad is not None]
            if live:
                self._step_list(gi, group, live, len(group['params']),
                                sum((p.numel() + self._chunk() - 1) // self._chunk() for p in group['params']))
        return loss

    @staticmethod
    def _chunk():
        return _lib.raw('uz_adam_chunk_elems')()

    @torch.no_grad()
    def step_params(self, params, key):
        """Adam step for a SUBSET of the parameters (those of ``params`` that have a gradient) with its own descriptor
        table ``key``: b200.dp runs it per gradient bucket as soon as the bucket is complete, next to the rest of backward
        (reference train_model.py:119 optimizer.step(), executed piecewise; every parameter is still updated exactly once
        per step with its complete gradient).  Hyper-parameters come from the parameter group of params[0]."""
        live = [p for p in params if p.grad is not None]
        if not live:
            return
        group = next(g for g in self.param_groups if any(q is live[0] for q in g['params']))
        chunk = self._chunk()
        self._step_list(('subset', key), group, live, len(params), sum((p.numel() + chunk - 1) // chunk for p in params))

    def _step_list(self, gi, group, live, nparam_max, nchunk_max):
        chunk = self._chunk()
        dev = live[0].device
        if not live[0].is_cuda:
            raise _lib.UnetZooLibError('FusedAdam needs CUDA parameters (no CPU fallback path)')
        # cheap key first: the gradient pointers (new view objects every step, same buffers) + the identity of the state
        # tensors (replaced only by load_state_dict / reset); the full rows are rebuilt only when it changes
        gen = self.__dict__.setdefault('_state_gen', 0)
        state = self.state
        # weight gradients still sitting in their split-K slabs (kern.wgrad_reducer on hold): the fused update + packing
        # kernel sums the splits itself, the OIHW gradient tensor is never written
        slabs = {}
        if self.packer is not None:
            from . import kern
            if kern.wgrad_reducer.hold and kern.wgrad_reducer.items:
                slabs = kern.wgrad_reducer.take([p.grad.data_ptr() for p in live if p.data_ptr() in self.packer.rows])
                kern.wgrad_reducer.flush()                       # whatever the packer does not cover
        slab_key = tuple((slabs[p.grad.data_ptr()][0].data_ptr(), slabs[p.grad.data_ptr()][1])
                         if p.grad.data_ptr() in slabs else (0, 0) for p in live) if slabs else ()
        quick = (gen, tuple(p.grad.data_ptr() for p in live),
                 tuple(id(state[p].get('exp_avg')) if p in state else 0 for p in live), slab_key)
        cache = self.__dict__.setdefault('_row_cache', {})
        hit = cache.get(gi)
        if hit is not None and hit[0] == quick:
            rows, key = hit[1], hit[2]
        else:
            rows, key = [], []
            for p in live:
                st = self._state(p)
                g = p.grad
                if g.dtype != torch.float32 or not g.is_contiguous() or not p.is_contiguous():
                    raise _lib.UnetZooLibError('FusedAdam expects dense fp32 parameters and gradients')
                rows.append((p.data_ptr(), g.data_ptr(), st['exp_avg'].data_ptr(), st['exp_avg_sq'].data_ptr(),
                             st['step'].data_ptr(), p.numel()))
                key.append(rows[-1])     # every raw pointer the table holds: a replaced state tensor invalidates it
            key = tuple(key) + (slab_key,)
            cache[gi] = (quick, rows, key)
        # two persistent table sets per group, allocated on first use (never inside a stream capture): one for eager
        # steps, one for a captured step whose memcpy nodes must keep reading the pointers they were captured with
        capturing = torch.cuda.is_current_stream_capturing()
        bufs = self._tables.get(gi)
        if bufs is None:
            nparam = nparam_max
            nitem_max = 0
            if self.packer is not None:
                for q in group['params']:
                    hit = self.packer.rows.get(q.data_ptr())
                    if hit is not None:
                        f = struct.unpack('<QQQiiiiii', hit[0])
                        nitem_max += _lib.raw('uz_adam_pack_items')(f[6], f[7], f[5])
            bufs = []
            for _ in range(2):
                bufs.append({'key': None, 'n': 0, 'nt': 0, 'ni': 0,
                             'host_d': torch.empty(nparam * 48, dtype=torch.uint8).pin_memory(),
                             'host_c': torch.empty(nchunk_max * 2, dtype=torch.int32).pin_memory(),
                             'host_p': torch.empty(nparam * 56, dtype=torch.uint8).pin_memory(),
                             'descs': torch.empty(nparam * 48, dtype=torch.uint8, device=dev),
                             'chunks': torch.empty(nchunk_max * 2, dtype=torch.int32, device=dev),
                             'packs': torch.empty(nparam * 56, dtype=torch.uint8, device=dev),
                             'host_i': torch.empty(max(2 * nitem_max, 2), dtype=torch.int32).pin_memory(),
                             'items': torch.empty(max(2 * nitem_max, 2), dtype=torch.int32, device=dev)})
            self._tables[gi] = bufs
        tab = bufs[1 if capturing else 0]
        if tab['key'] != key:
            if tab.get('uploaded') is not None:
                # the previous step's non-blocking upload reads the pinned buffers that are rewritten below
                tab['uploaded'].synchronize()
            raw = b''.join(struct.pack('<QQQQQq', *r) for r in rows)
            table, prow, items = [], [], []
            pk_rows = self.packer.rows if self.packer is not None else {}
            for ti, r in enumerate(rows):
                hit = pk_rows.get(r[0])
                if hit is not None:
                    # conv weight with packed copies: fused update + re-pack, one block per 32 x 32 x 9-tap tile
                    _, wf, wd, cout, cin, taps, coutp, cinp, _ = struct.unpack('<QQQiiiiii', hit[0])
                    for it in range(_lib.raw('uz_adam_pack_items')(coutp, cinp, taps)):
                        items += [len(prow), it]
                    sl = slabs.get(r[1])
                    if sl is not None and (sl[2], sl[3]) != (coutp, cinp):
                        raise _lib.UnetZooLibError('FusedAdam: weight-gradient slab and packed weight disagree on the padded '
                                                   'channel counts')
                    prow.append(struct.pack('<QQiiiiiiQii', wf, wd, ti, cout, cin, taps, coutp, cinp,
                                            sl[0].data_ptr() if sl is not None else 0, sl[1] if sl is not None else 0, 0))
                    continue
                for c in range((r[5] + chunk - 1) // chunk):
                    table += [ti, c]
            tab['host_d'][:len(raw)] = torch.frombuffer(bytearray(raw), dtype=torch.uint8)
            if table:
                tab['host_c'][:len(table)] = torch.tensor(table, dtype=torch.int32)
            if prow:
                praw = b''.join(prow)
                tab['host_p'][:len(praw)] = torch.frombuffer(bytearray(praw), dtype=torch.uint8)
                tab['host_i'][:len(items)] = torch.tensor(items, dtype=torch.int32)
            tab['key'], tab['n'], tab['nt'], tab['ni'] = key, len(table) // 2, len(rows), len(items) // 2
            tab['stale'] = True
        descs, chunks, nchunks = tab['descs'], tab['chunks'], tab['n']
        if capturing:
            # no memcpy nodes inside a captured step (they break the back-to-back kernel scheduling of the graph):
            # the tables of a capture are static, finish_capture() uploads them once before the first replay
            tab['dirty'] = True
        elif tab.get('stale', True):
            descs.copy_(tab['host_d'], non_blocking=True)
            chunks.copy_(tab['host_c'], non_blocking=True)
            if tab['ni']:
                tab['packs'].copy_(tab['host_p'], non_blocking=True)
                tab['items'].copy_(tab['host_i'], non_blocking=True)
            if tab.get('uploaded') is None:
                tab['uploaded'] = torch.cuda.Event()
            tab['uploaded'].record()
            tab['stream'] = torch._C._cuda_getCurrentRawStream(torch._C._cuda_getDevice())
            tab['stale'] = False
        elif tab.get('stream') != torch._C._cuda_getCurrentRawStream(torch._C._cuda_getDevice()):
            torch.cuda.current_stream().wait_event(tab['uploaded'])      # uploaded on another stream
        b1, b2 = group['betas']
        if tab['ni']:
            _lib.call('uz_adam_pack_step', descs.data_ptr(), tab['nt'], chunks.data_ptr(), nchunks, tab['packs'].data_ptr(),
                      tab['items'].data_ptr(), tab['ni'], float(group['lr']), float(b1), float(b2), float(group['eps']),
                      float(group['weight_decay']), torch._C._cuda_getCurrentRawStream(torch._C._cuda_getDevice()))
            return
        _lib.call('uz_adam_step_batched', descs.data_ptr(), tab['nt'], chunks.data_ptr(), nchunks,
                  float(group['lr']), float(b1), float(b2), float(group['eps']), float(group['weight_decay']),
                  torch._C._cuda_getCurrentRawStream(torch._C._cuda_getDevice()))
