"""CTU-row bands across the GPUs of one box (BASELINE.json configs[3]).

Host-side plan + the neighbour halo exchange, written against torch.distributed so that the same code runs over NCCL on
GPUs and over gloo on CPUs (tests).  Each rank owns a contiguous band of CTU rows of the picture and holds the rows of the
reference frame it reconstructed itself; before searching it needs HALO more rows from the band above and below."""
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


def halo_plan(height, ctu_rows, world, rank, chroma=False):
    """rows to send / receive: dict with 'send_up', 'send_down', 'recv_up', 'recv_down' = (row0, n_rows) or None.
    'up' is the neighbour with the smaller rank."""
    halo = HALO_CHROMA if chroma else HALO_LUMA
    y0, y1 = band_sample_rows(height, ctu_rows, world, rank, chroma)
    plan = {"send_up": None, "send_down": None, "recv_up": None, "recv_down": None}
    if rank > 0 and y1 > y0:
        py0, py1 = band_sample_rows(height, ctu_rows, world, rank - 1, chroma)
        if py1 > py0:
            plan["send_up"] = (y0, min(halo, y1 - y0))
            plan["recv_up"] = (max(py1 - halo, py0), min(halo, py1 - py0))
    if rank < world - 1 and y1 > y0:
        ny0, ny1 = band_sample_rows(height, ctu_rows, world, rank + 1, chroma)
        if ny1 > ny0:
            plan["send_down"] = (max(y1 - halo, y0), min(halo, y1 - y0))
            plan["recv_down"] = (ny0, min(halo, ny1 - ny0))
    return plan


def exchange_halos(dist, planes, height, ctu_rows, world, rank):
    """planes: [Y, U, V] torch tensors of the FULL picture size living on this rank (only its own band rows are valid).
    After the call the halo rows above and below the band are valid too.  One batch of point-to-point operations."""
    ops, keep = [], []
    for c, t in enumerate(planes):
        p = halo_plan(height, ctu_rows, world, rank, chroma=c > 0)
        for key, peer in (("send_up", rank - 1), ("send_down", rank + 1)):
            if p[key]:
                r0, n = p[key]
                buf = t[r0:r0 + n].contiguous()
                keep.append(buf)
                ops.append(dist.P2POp(dist.isend, buf, peer))
        for key, peer in (("recv_up", rank - 1), ("recv_down", rank + 1)):
            if p[key]:
                r0, n = p[key]
                ops.append(dist.P2POp(dist.irecv, t[r0:r0 + n], peer))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    return len(ops)


class FrameHaloExchanger:
    """Halo exchange for an hb.Frame resident on this rank's GPU: the band's boundary rows are exported into tight torch
    tensors, swapped with the neighbours over NCCL (NVLink / NVSwitch), imported on the other side, and the replicated
    border is refreshed.  Buffers are allocated once."""

    def __init__(self, torch, dist, ctx, width, height, world, rank, device):
        self.torch, self.dist, self.ctx = torch, dist, ctx
        self.h, self.world, self.rank = height, world, rank
        self.ctu_rows = (height + 63) // 64
        self.items = []                       # (plane, kind, row0, n_rows, peer, tensor)
        for c in range(3):
            w = width // 2 if c else width
            p = halo_plan(height, self.ctu_rows, world, rank, chroma=c > 0)
            for key, peer in (("send_up", rank - 1), ("send_down", rank + 1), ("recv_up", rank - 1), ("recv_down", rank + 1)):
                if p[key]:
                    r0, n = p[key]
                    self.items.append((c, key[:4], r0, n, peer, torch.empty((n, w), dtype=torch.uint8, device=device)))
        self.bytes_per_exchange = sum(t.numel() for (_, kind, _, _, _, t) in self.items if kind == "send")

    def exchange(self, frame):
        dist = self.dist
        for c, kind, r0, n, _, t in self.items:
            if kind == "send":
                frame.export_rows(c, r0, n, t.data_ptr())
        self.ctx.sync()                       # exported rows are complete before NCCL reads them on torch's stream
        ops = [dist.P2POp(dist.isend if kind == "send" else dist.irecv, t, peer) for (_, kind, _, _, peer, t) in self.items]
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        self.torch.cuda.synchronize()
        for c, kind, r0, n, _, t in self.items:
            if kind == "recv":
                frame.import_rows(c, r0, n, t.data_ptr())
        frame.pad()
