"""CTU-row bands across the GPUs of one box (BASELINE.json configs[3]).

Host-side plan + the halo exchange, written against torch.distributed so that the same code runs over NCCL on GPUs and
over gloo on CPUs (tests).  Each rank owns a contiguous band of CTU rows of the picture and holds the rows of the
reference frame it reconstructed itself; before searching it needs HALO more rows above and below its band.  Those rows
belong to the neighbouring bands -- to rank +-1 as a rule, and to rank +-2, ... as well when a band is lower than the halo
(a one-CTU-row band has 64 luma rows, the halo is 68: 720p on 8 GPUs) -- so the plan is the intersection of every rank's
halo interval with every other rank's band."""
HALO_LUMA, HALO_CHROMA = 68, 36       # 64 search + 4 (8-tap) luma; 32 + 1 + 2 (4-tap) + rounding slack chroma


def band_ctu_rows(ctu_rows, world, rank):
    """(first CTU row, number of CTU rows) of `rank`; sizes differ by at most one, earlier ranks get the longer bands"""
    base, extra = divmod(ctu_rows, world)
    n = base + (1 if rank < extra else 0)
    row0 = rank * base + min(rank, extra)
    return row0, n


def band_sample_rows(height, ctu_rows, world, rank, chroma=False):
    """[y0, y1) in samples of that plane"""
    r0, n = band_ctu_rows(ctu_rows, world, rank)
    unit = 32 if chroma else 64
    h = height // 2 if chroma else height
    return min(r0 * unit, h), min((r0 + n) * unit, h)


def _halo_intervals(height, ctu_rows, world, rank, chroma):
    """the two row intervals outside its band that `rank` reads: [y0 - halo, y0) and [y1, y1 + halo), clipped to the plane"""
    halo = HALO_CHROMA if chroma else HALO_LUMA
    h = height // 2 if chroma else height
    y0, y1 = band_sample_rows(height, ctu_rows, world, rank, chroma)
    if y1 <= y0:
        return []
    return [(max(0, y0 - halo), y0), (y1, min(h, y1 + halo))]


def halo_transfers(height, ctu_rows, world, rank, chroma=False):
    """{'send': [(peer, row0, n_rows)], 'recv': [(peer, row0, n_rows)]} for `rank`, both sorted by (peer, row0) so that the two
    sides of every transfer enumerate it in the same order.  A transfer is the overlap of the receiver's halo interval with the
    sender's own band."""
    def overlaps(needy, owner):
        o0, o1 = band_sample_rows(height, ctu_rows, world, owner, chroma)
        out = []
        for a, b in _halo_intervals(height, ctu_rows, world, needy, chroma):
            lo, hi = max(a, o0), min(b, o1)
            if hi > lo:
                out.append((lo, hi - lo))
        return out
    send, recv = [], []
    for peer in range(world):
        if peer == rank:
            continue
        recv += [(peer, r0, n) for (r0, n) in overlaps(rank, peer)]
        send += [(peer, r0, n) for (r0, n) in overlaps(peer, rank)]
    return {"send": sorted(send), "recv": sorted(recv)}


def exchange_halos(dist, planes, height, ctu_rows, world, rank):
    """planes: [Y, U, V] torch tensors of the FULL picture size living on this rank (only its own band rows are valid).
    After the call the halo rows above and below the band are valid too.  One batch of point-to-point operations."""
    ops, keep = [], []
    for c, t in enumerate(planes):
        p = halo_transfers(height, ctu_rows, world, rank, chroma=c > 0)
        for peer, r0, n in p["send"]:
            buf = t[r0:r0 + n].contiguous()
            keep.append(buf)
            ops.append(dist.P2POp(dist.isend, buf, peer))
        for peer, r0, n in p["recv"]:
            ops.append(dist.P2POp(dist.irecv, t[r0:r0 + n], peer))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    return len(ops)


class FrameHaloExchanger:
    """Halo exchange for an hb.Frame resident on this rank's GPU, entirely on the device timeline: the band's boundary rows
    are exported into tight torch tensors on the library context's stream, swapped with the other bands by one grouped NCCL
    send/recv (NVLink / NVSwitch) that torch orders after that stream with events, imported on the other side and the
    replicated border is refreshed -- the host queues all of it and never waits.  Whatever is queued on the context
    afterwards (the next frame's search) runs behind the exchange.  Buffers are allocated once."""

    def __init__(self, torch, dist, ctx, width, height, world, rank, device):
        self.torch, self.dist, self.ctx = torch, dist, ctx
        self.h, self.world, self.rank = height, world, rank
        self.ctu_rows = (height + 63) // 64
        self.items = []                       # (plane, kind, row0, n_rows, peer, tensor)
        for c in range(3):
            w = width // 2 if c else width
            p = halo_transfers(height, self.ctu_rows, world, rank, chroma=c > 0)
            for kind in ("send", "recv"):
                for peer, r0, n in p[kind]:
                    self.items.append((c, kind, r0, n, peer, torch.empty((n, w), dtype=torch.uint8, device=device)))
        self.bytes_per_exchange = sum(t.numel() for (_, kind, _, _, _, t) in self.items if kind == "send")
        # the library's stream as a torch stream: collectives issued under it are ordered against it by events on the device
        self.stream = torch.cuda.ExternalStream(ctx.stream_ptr(), device=device)

    def exchange(self, frame):
        """queue export -> send/recv -> import -> border on the context's stream; returns without waiting"""
        dist = self.dist
        with self.torch.cuda.stream(self.stream):
            for c, kind, r0, n, _, t in self.items:
                if kind == "send":
                    frame.export_rows(c, r0, n, t.data_ptr())
            ops = [dist.P2POp(dist.isend if kind == "send" else dist.irecv, t, peer) for (_, kind, _, _, peer, t) in self.items]
            if ops:
                for req in dist.batch_isend_irecv(ops):
                    req.wait()                # stream-ordered for NCCL: makes the current (= the context's) stream wait, not the host
            for c, kind, r0, n, _, t in self.items:
                if kind == "recv":
                    frame.import_rows(c, r0, n, t.data_ptr())
            frame.pad()


class PeerHaloPuller:
    """The halo exchange of a SET of pictures resident on this rank's GPU through peer memory (NVLink / NVSwitch, one process per
    GPU): at set-up every rank exports its pictures and one "finished" event per picture (CUDA IPC) and opens those of the ranks whose
    bands its halos reach into (all_gather_object, once).  Per picture and frame the exchange then is: make the stream wait for the
    owners' events, ONE copy kernel that reads the halo rows straight out of the owners' HBM, the border refresh -- two launches queued
    by one library call, no staging buffers, no collective, no host wait.  `mark_ready(j)` records this rank's event of picture j."""

    def __init__(self, hb, dist, ctx, frames, width, height, world, rank):
        self.ctx, self.frames, self.rank = ctx, frames, rank
        ctu_rows = (height + 63) // 64
        self.events = [hb.IpcEvent(ctx) for _ in frames]
        mine = [(f.ipc_export(), e.handle) for f, e in zip(frames, self.events)]
        everyone = [None] * world
        dist.all_gather_object(everyone, mine)
        need = {}                                      # peer -> [(plane, row0, n_rows)]
        for c in range(3):
            for peer, r0, n in halo_transfers(height, ctu_rows, world, rank, chroma=c > 0)["recv"]:
                need.setdefault(peer, []).append((c, r0, n))
        self.peers = sorted(need)
        self.views = [[hb.Frame.ipc_open(ctx, everyone[p][j][0]) for p in self.peers] for j in range(len(frames))]
        self.peer_events = [[hb.IpcEvent(ctx, everyone[p][j][1]) for p in self.peers] for j in range(len(frames))]
        self.spans = [(i, c, r0, n) for i, p in enumerate(self.peers) for (c, r0, n) in need[p]]
        width_of = lambda c: width // 2 if c else width
        self.bytes_per_exchange = sum(n * width_of(c) for (_, c, _, n) in self.spans)      # pulled, per picture
        self._prepared = None

    def mark_ready(self, j):
        self.events[j].record()

    def pull(self, j):
        for e in self.peer_events[j]:
            e.wait()
        if self._prepared is None:           # the ctypes argument arrays of every picture, built once (the per-frame path is two library calls)
            self._prepared = [self.frames[k].pull_rows_prepare(self.views[k], self.spans) for k in range(len(self.frames))]
        self.frames[j].pull_rows_prepared(self._prepared[j], refresh_border=True)

    def close(self):
        for row in self.views:
            for v in row:
                v.close()
        for row in self.peer_events:
            for e in row:
                e.close()
        for e in self.events:
            e.close()
