"""Multi-GPU IAS pseudo-labelling: windows striped over the ranks, threshold state handed rank to rank.

The reference pseudo-labels on one GPU in one process (no ``dist`` call in
``workflows/pseudo_label_generator.py``).  The path shards naturally: phase A
(softmax / arg-max / histograms) and phase C (mask / counts) are independent per group; the only
cross-group dependencies are the sequential ``class_threshold`` f64[C] (:207-209) and, after phase C,
``class_mean_probs`` f64[C] (:100-105).  With R ranks (one process per GPU) the pinned global order of
image groups is cut into windows of ``window_images`` images and window w belongs to rank w mod R:

    rank r, local window j  (global window w = j*R + r), three window slots:
        main stream    phase A of window j, phase C of window j-2, phase A of window j+1, ...
        chain stream   behind A(j):  recv thr f64[C] from rank r-1   (ncclRecv over NVLink; skipped for w = 0, which starts from 0.9)
                                     phase B of window j-1           (device scan over the window's groups)
                                     send thr f64[C] to rank r+1     (ncclSend; skipped for the last window)

The 152-byte token goes round the ring once per round of R windows; as long as R x (scan + hop) is
shorter than one window of phase A + C the chain is hidden behind the bandwidth-bound work, which is
what makes image/s scale with R.  No other data-path collective exists.  At the end the per-group
(confidence sum, count)[C] of all windows are all-gathered (a few hundred KB) and every rank replays
the tiny ``class_mean_probs`` EMA in global order; the final thresholds ride in the same all-gather (taken
from the rank that scanned the last window); ``statics_class`` falls out of the gathered counts.

Results are bit-identical to a single-rank run over the same global order (tests/test_sharded_gpu.py;
on CPU with gloo and a host stand-in engine in tests/test_sharded_gloo.py).  With R = 1 the same code
is the single-GPU windowed pipeline.
"""

from __future__ import annotations

import torch
import torch.distributed as dist


_STREAMS = {}


def device_stream(device, kind):
    """One host-to-device copy stream ('copy'), one device-to-host stream ('out'), one high-priority chain stream ('chain') and one
    high-priority stream for phase C beside phase A ('select') per device for the life of the process (NCCL and
    the caching allocator keep per-stream state; a fresh stream per run would pay for it again)."""
    device = torch.device(device)
    key = (kind, device.index if device.index is not None else torch.cuda.current_device())
    st = _STREAMS.get(key)
    if st is None:
        st = _STREAMS[key] = torch.cuda.Stream(device, priority=-1 if kind in ('chain', 'select') else 0)
    return st


class TokenRing:
    """Mailboxes of the peer-memory token ring (``hiast_ring_*`` / ``hiast_ias_threshold_scan_ring``): every rank owns one
    4 KB mailbox on its GPU and maps the mailbox of the NEXT rank through CUDA IPC.  Creation is collective (one
    ``all_gather_object`` of the 64-byte handles).  ``token(w, n_windows_total)`` gives the scan kernel of global window w
    its (mailbox_in, in_seq, mailbox_out, out_seq); sequence numbers keep growing across runs (``advance``).

    ``TokenRing.get`` returns None -- and the callers fall back to ``torch.distributed`` send / recv -- on CPU, with one
    rank, when ``HIAST_RING=nccl`` is set, or when any rank could not create or map a mailbox (ranks on different nodes)."""

    _cache = {}

    def __init__(self, local_ptr, next_ptr, rank, world):
        self.local, self.next, self.rank, self.world = local_ptr, next_ptr, rank, world
        self.base = 0

    @classmethod
    def get(cls, device, rank, world, pg=None):
        import ctypes as C
        import os
        from ._lib import lib
        device = torch.device(device)
        if world < 2 or device.type != 'cuda' or os.environ.get('HIAST_RING', 'peer') == 'nccl':
            return None
        key = (id(pg), device.index if device.index is not None else torch.cuda.current_device(), rank, world)
        if key in cls._cache:
            return cls._cache[key]
        ring = None
        local, handle = C.c_void_p(), (C.c_ubyte * 64)()
        with torch.cuda.device(device):
            ok = lib().hiast_ring_create(C.byref(local), C.cast(handle, C.c_void_p)) == 0
            infos = [None] * world
            dist.all_gather_object(infos, (ok, bytes(handle), _host_id()), group=pg)
            nxt = infos[(rank + 1) % world]
            peer = C.c_void_p()
            if ok and all(i[0] for i in infos) and len({i[2] for i in infos}) == 1:
                buf = (C.c_ubyte * 64).from_buffer_copy(nxt[1])
                ok = lib().hiast_ring_open(C.cast(buf, C.c_void_p), C.byref(peer)) == 0
            else:
                ok = False
            oks = [None] * world
            dist.all_gather_object(oks, ok, group=pg)
            if all(oks):
                ring = cls(local.value, peer.value, rank, world)
            else:
                if peer.value:
                    lib().hiast_ring_close(peer)
                if local.value:
                    lib().hiast_ring_destroy(local)
        cls._cache[key] = ring
        return ring

    def token(self, w, n_windows_total):
        t_in = self.local if w > 0 else None
        t_out = self.next if w < n_windows_total - 1 else None
        return (t_in, self.base + w, t_out, self.base + w + 1)

    def advance(self, n_windows_total):
        """Every rank calls this once per finished job: the next job's sequence numbers start above this one's."""
        self.base += n_windows_total + 1


def _host_id():
    import socket
    try:
        with open('/proc/sys/kernel/random/boot_id') as f:
            return socket.gethostname() + ':' + f.read().strip()
    except OSError:
        return socket.gethostname()


def window_owner(w, world_size):
    return w % world_size


def local_windows(n_windows_total, rank, world_size):
    """Global indices of the windows rank ``rank`` owns, in processing order."""
    return list(range(rank, n_windows_total, world_size))


def window_images(w, window_size, n_images_total):
    """(first global image, image count) of global window w."""
    i0 = w * window_size
    return i0, max(0, min(window_size, n_images_total - i0))


class ShardedIAS:
    """Drives one rank's engine over device-resident windows (an ``IASEngine`` with room for two or, better, three
    windows; or any object with the same phase methods and state tensors -- the CPU tests use a host stand-in).

    Schedule with three window slots (``engine.max_images >= 3 * window_size``), per owned window j:

        main stream    A(j)                        C(j-2)   A(j+1)                     C(j-1) ...
        chain stream          [recv] B(j-1) [send]                 [recv] B(j) [send]

    The chain of window j-1 is queued behind A(j) on a high-priority side stream and A(j+1) waits for it, so it never
    competes with a running phase A for SMs and runs in the shadow of phase C of window j-2.  On R ranks the token needs
    r hops to reach rank r: the ranks fall into a stagger of one hop each once, after which every token is already there
    when it is needed.  With two slots C(j-1) follows B(j-1) directly (the round-1 schedule)."""

    def __init__(self, engine, window_size, n_images_total, rank=None, world_size=None, process_group=None):
        if window_size % engine.B:
            raise ValueError('window_size must be a multiple of the group (batch) size')
        if engine.max_images < 2 * window_size:
            raise ValueError('the engine needs room for at least two windows')
        self.engine = engine
        self.pg = process_group
        use_dist = dist.is_available() and dist.is_initialized()
        self.rank = (dist.get_rank(process_group) if use_dist else 0) if rank is None else rank
        self.world = (dist.get_world_size(process_group) if use_dist else 1) if world_size is None else world_size
        self.window_size = int(window_size)
        self.n_slots = 3 if engine.max_images >= 3 * window_size else 2
        self.n_total = int(n_images_total)
        self.n_windows_total = (self.n_total + self.window_size - 1) // self.window_size
        self.my_windows = local_windows(self.n_windows_total, self.rank, self.world)
        self.cuda = torch.is_tensor(engine.plbl) and engine.plbl.is_cuda
        self.ring = TokenRing.get(engine.device, self.rank, self.world, process_group) \
            if (self.cuda and use_dist and self.world > 1) else None
        if self.cuda:
            self.side = device_stream(engine.device, 'chain')
            self.select = device_stream(engine.device, 'select')
            self.ev_a = [torch.cuda.Event() for _ in range(self.n_slots)]
            self.ev_b = [torch.cuda.Event() for _ in range(self.n_slots)]
        self._stash_conf = self._stash_counts = None

    def _global(self, group_rank):
        return dist.get_global_rank(self.pg, group_rank) if self.pg is not None else group_rank

    def _slot(self, j):
        return (j % self.n_slots) * self.window_size

    def _n(self, j):
        return window_images(self.my_windows[j], self.window_size, self.n_total)[1]

    def run(self, window_logits, on_window=None):
        """``window_logits(w) -> logits f32 [n_w,C,H,W]`` on the device for GLOBAL window index w (called once per
        owned window).  ``on_window(w, plbl, counts, thr_groups)`` receives device views of window w's results right
        after phase C is enqueued; they stay valid until the next-but-one window is started.  Returns (class_threshold,
        class_mean_probs, statics_class) device tensors."""
        e = self.engine
        wins = self.my_windows
        gw = self.window_size // e.B
        dev = e.thr_state.device
        k = max(len(wins), 1)
        self._stash_conf = torch.zeros((k, gw, e.C), dtype=torch.int64, device=dev)
        self._stash_counts = torch.zeros((k, self.window_size, e.C), dtype=torch.int64, device=dev)
        if self.cuda and self.n_slots >= 3 and int(getattr(e, 'reserve_sms', 0) or 0) > 0:
            return self._run_concurrent(window_logits, on_window)
        lag = self.n_slots - 2                    # windows between a window's chain and its outputs
        main = torch.cuda.current_stream(e.device) if self.cuda else None
        b_done = c_done = 0

        def advance(n_closed, final):
            nonlocal b_done, c_done
            b_target = n_closed if final else n_closed - 1
            while b_done < b_target:
                self._chain(b_done, n_closed - 1)
                b_done += 1
            c_target = b_done if final else b_done - lag
            while c_done < c_target:
                self._outputs(c_done, main, on_window)
                c_done += 1

        for j, w in enumerate(wins):
            logits = window_logits(w)
            n = self._n(j)
            if logits.shape[0] != n:
                raise ValueError('window %d must hold %d images, got %d' % (w, n, logits.shape[0]))
            e.phase_a(logits, self._slot(j))
            if self.cuda:
                self.ev_a[j % self.n_slots].record(main)
            advance(j + 1, final=False)
            if self.cuda and b_done:
                # the next phase A waits for the chain just queued: the ranks fall into a stagger of one hop each ONCE and
                # the receive / scan kernels never share the SMs with a phase A (see _WindowPipeline._close)
                main.wait_event(self.ev_b[(b_done - 1) % self.n_slots])
        advance(len(wins), final=True)
        return self.finish_state()

    def _run_concurrent(self, window_logits, on_window):
        """Schedule with SMs reserved from phase A (``engine.reserve_sms`` > 0): a phase-A CTA takes a whole SM, so the launch
        leaves ``reserve_sms`` SMs empty, and the chain stream runs the scan (token hand-off inside) AND phase C of window j there
        while phase A of window j+1 streams on the others:

            main stream     A(j) .......... A(j+1) .......... A(j+2) ...        (A(j+3) waits until C(j) has left its slot)
            chain stream          scan(j)           scan(j+1)
            select stream                C(j) ..........    C(j+1) ..........

        Phase C moves 6 B/px against phase A's 81: it fits in phase A's shadow on a handful of SMs, and its traffic fills the DRAM
        cycles phase A leaves idle instead of being queued between two launches."""
        e = self.engine
        main = torch.cuda.current_stream(e.device)
        ev_c = [None] * self.n_slots
        for j, w in enumerate(self.my_windows):
            s = j % self.n_slots
            if ev_c[s] is not None:
                main.wait_event(ev_c[s])                  # window j-3 has left this slot
            logits = window_logits(w)
            n = self._n(j)
            if logits.shape[0] != n:
                raise ValueError('window %d must hold %d images, got %d' % (w, n, logits.shape[0]))
            e.phase_a(logits, self._slot(j))
            self.ev_a[s].record(main)
            self.side.wait_event(self.ev_a[s])
            with torch.cuda.stream(self.side):
                self._chain_body(j)
                self.ev_b[s].record(self.side)
            # Phase C has its own stream: the scan of window j+1 -- and with it the token the next rank is waiting for -- must
            # not queue behind phase C of window j.  On the reserved SMs phase C takes about as long as one phase A; with the scan
            # in front of it in ONE stream the cycle was a little longer than a phase A on the ranks the token reaches last, and
            # that stream reached the end of the job up to a millisecond behind.
            self.select.wait_event(self.ev_b[s])
            with torch.cuda.stream(self.select):
                self._outputs(j, None, on_window)
                ev_c[s] = torch.cuda.Event()
                ev_c[s].record(self.select)
        return self.finish_state()

    def _chain_body(self, j):
        e = self.engine
        w = self.my_windows[j]
        ring = self.world > 1
        if self.ring is not None:                     # hand-off fused into the scan kernel, over peer memory
            e.phase_b(self._slot(j), self._n(j), token=self.ring.token(w, self.n_windows_total))
            return
        if ring and w > 0:
            dist.recv(e.thr_state, src=self._global((self.rank - 1) % self.world), group=self.pg)
        e.phase_b(self._slot(j), self._n(j))
        if ring and w < self.n_windows_total - 1:
            dist.send(e.thr_state, dst=self._global((self.rank + 1) % self.world), group=self.pg)

    def _chain(self, j, latest_closed):
        e = self.engine
        w = self.my_windows[j]
        ring = self.world > 1

        def body():
            if self.ring is not None:                 # hand-off fused into the scan kernel, over peer memory
                e.phase_b(self._slot(j), self._n(j), token=self.ring.token(w, self.n_windows_total))
                return
            if ring and w > 0:
                dist.recv(e.thr_state, src=self._global((self.rank - 1) % self.world), group=self.pg)
            e.phase_b(self._slot(j), self._n(j))
            if ring and w < self.n_windows_total - 1:
                dist.send(e.thr_state, dst=self._global((self.rank + 1) % self.world), group=self.pg)

        if not self.cuda:
            body()
            return
        self.side.wait_event(self.ev_a[latest_closed % self.n_slots])
        with torch.cuda.stream(self.side):
            body()
            self.ev_b[j % self.n_slots].record(self.side)

    def _outputs(self, j, main, on_window):
        e = self.engine
        slot, n = self._slot(j), self._n(j)
        if self.cuda and main is not None:
            main.wait_event(self.ev_b[j % self.n_slots])
        e.phase_c(slot, n)
        g0, g = slot // e.B, (n + e.B - 1) // e.B
        self._stash_conf[j, :g].copy_(torch.as_tensor(e.confsum[g0:g0 + g]))
        self._stash_counts[j, :n].copy_(torch.as_tensor(e.counts[slot:slot + n]))
        if on_window is not None:
            on_window(self.my_windows[j], e.plbl[slot:slot + n], e.counts[slot:slot + n], e.thr_groups[g0:g0 + g])

    def warm_collective(self):
        """Runs the end-of-job all-gather once with THIS job's shapes on zeros (collective: every rank calls it).  NCCL
        connects the transports of an algorithm / protocol at its first use (tens of milliseconds), and which one it picks
        depends on the message size; a caller that times the job (bench.py) or cannot afford the hiccup at the end of the
        first job calls this beforehand."""
        if self.world < 2:
            return
        e = self.engine
        gw = self.window_size // e.B
        kmax = max((self.n_windows_total + self.world - 1) // self.world, 1)
        packed = torch.zeros((kmax * gw * 2 + 1, e.C), dtype=torch.int64, device=e.thr_state.device)
        dist.all_gather([torch.empty_like(packed) for _ in range(self.world)], packed, group=self.pg)

    def finish_state(self):
        """ONE all-gather carries every rank's per-group confidence sums and kept-pixel counts plus its threshold state;
        every rank then replays the mean-prob EMA over all groups in global order and takes the final thresholds from the
        rank that scanned the last window."""
        e = self.engine
        C, B = e.C, e.B
        gw = self.window_size // B
        kmax = max((self.n_windows_total + self.world - 1) // self.world, 1)
        dev = e.thr_state.device
        if self.cuda:
            torch.cuda.current_stream(e.device).wait_stream(self.side)
            torch.cuda.current_stream(e.device).wait_stream(self.select)
        if self.ring is not None:
            self.ring.advance(self.n_windows_total)
        packed = torch.zeros((kmax * gw * 2 + 1, C), dtype=torch.int64, device=dev)
        nloc = len(self.my_windows)
        if nloc and self._stash_conf is not None:
            groups = self._stash_counts[:nloc].view(nloc, gw, B, C).sum(dim=2)
            packed[:kmax * gw * 2].view(kmax, gw, 2, C)[:nloc, :, 0] = self._stash_conf[:nloc]
            packed[:kmax * gw * 2].view(kmax, gw, 2, C)[:nloc, :, 1] = groups
        packed[-1] = torch.as_tensor(e.thr_state).view(torch.int64)      # the 19 doubles ride along bit for bit
        if self.world > 1:
            gathered = [torch.empty_like(packed) for _ in range(self.world)]
            dist.all_gather(gathered, packed, group=self.pg)
        else:
            gathered = [packed]
        rows = []
        for w in range(self.n_windows_total):
            _, n = window_images(w, self.window_size, self.n_total)
            rows.append(gathered[w % self.world][:kmax * gw * 2].view(kmax, gw, 2, C)[w // self.world, :(n + B - 1) // B])
        if rows:
            allg = torch.cat(rows)
            confsum, counts = allg[:, 0].contiguous(), allg[:, 1].contiguous()
            e.mean_prob_from_groups(confsum, counts)
            statics = counts.sum(dim=0)
            last_owner = window_owner(self.n_windows_total - 1, self.world)
            e.thr_state.copy_(gathered[last_owner][-1].view(torch.float64))
        else:
            statics = torch.zeros(C, dtype=torch.int64, device=dev)
        return e.thr_state, e.mean_state, statics
