"""Multi-GPU strips: one process per GPU, the film split into horizontal strips, scene + BVH
replicated.  Host-side plumbing only:

* `partition` / `storage_rows`: which rows a rank owns and which it stores (owned + halo);
* `connect_strips`: exchange CUDA-IPC handles of the neighbours' temp-reservoir buffers through torch.distributed
  and hand them to rpt_frame_connect_peers — after that the temporal kernels push their boundary rows into the
  neighbours' halo rows over NVLink and the hand-over is ordered by device-side epoch flags (csrc/peer_sync.cu);
  no collective runs per frame;
* `exchange_halo`: the same halo hand-over written with torch.distributed send/recv on a tensor — the host-logic
  reference for the index arithmetic (exercised with gloo on CPU in tests/) and an NCCL alternative to the peer path.
"""
import ctypes as C

import restirpt
from restirpt import PeerInfo, P


def partition(height, world):
    """rows [begin, end) owned by each rank: equal strips, remainder rows to the first ranks"""
    base, extra = divmod(height, world)
    out, y = [], 0
    for r in range(world):
        n = base + (1 if r < extra else 0)
        out.append((y, y + n))
        y += n
    return out


def balanced_partition(bounds, costs, min_rows, quantum=4):
    """Strip boundaries that equalise the per-strip cost, from one measurement: `costs[r]` is the device time rank r
    spent on its rows `bounds[r]` (waiting for neighbours excluded).  The cost is taken as uniform inside each measured
    strip, so the cumulative cost is piecewise linear in the row index; the new boundaries are where it crosses
    k/N of the total.  Rows are rounded to `quantum` and no strip gets fewer than `min_rows` (>= the halo, so that a
    strip's boundary rows always come from its direct neighbour).  Pure function (tested on CPU)."""
    world = len(bounds)
    height = bounds[-1][1]
    total = float(sum(costs))
    if world == 1 or total <= 0:
        return list(bounds)
    cuts, r, acc = [], 0, 0.0
    for k in range(1, world):
        target = total * k / world
        while r < world - 1 and acc + costs[r] < target:
            acc += costs[r]
            r += 1
        b, e = bounds[r]
        frac = (target - acc) / costs[r] if costs[r] > 0 else 0.0
        cuts.append(b + min(max(frac, 0.0), 1.0) * (e - b))
    rows = [int(round(c / quantum)) * quantum for c in cuts]
    # enforce the minimum height from both ends
    prev = 0
    for k in range(world - 1):
        rows[k] = max(rows[k], prev + min_rows)
        prev = rows[k]
    nxt = height
    for k in range(world - 2, -1, -1):
        rows[k] = min(rows[k], nxt - min_rows)
        nxt = rows[k]
    edges = [0] + rows + [height]
    if any(edges[i + 1] - edges[i] < min(min_rows, height // world) for i in range(world)):
        return list(bounds)   # film too small to honour the minimum: keep what we have
    return [(edges[i], edges[i + 1]) for i in range(world)]


def storage_rows(row_begin, row_end, height, halo):
    return max(row_begin - halo, 0), min(row_end + halo, height)


class StripLink:
    def __init__(self, frame, rank, world, mode):
        self.frame, self.rank, self.world, self.mode = frame, rank, world, mode

    def describe(self):
        return ("temporal kernels store boundary rows into the neighbours' halo rows through CUDA-IPC peer memory "
                "(NVLink); hand-over ordered by device-side epoch flags; no per-frame collective")

    def error(self):
        return restirpt.device_lib().rpt_frame_peer_error(self.frame)

    def close(self):
        """every rank drops its mappings of the neighbours' buffers, then a barrier: after it any strip may be destroyed"""
        import torch.distributed as dist
        restirpt.device_lib().rpt_frame_disconnect_peers(self.frame)
        dist.barrier()


def connect_strips(renderer, frame, rank, world):
    """all-gather every rank's RptPeerInfo and connect this rank's frame to the strips above and below"""
    import torch.distributed as dist
    lib = restirpt.device_lib()
    mine = PeerInfo()
    status = lib.rpt_frame_export_peer(frame, C.byref(mine))
    if status != 0:
        raise restirpt.RestirptError(f"rpt_frame_export_peer failed: {status}")
    blobs = [None] * world
    dist.all_gather_object(blobs, bytes(mine))
    infos = [PeerInfo.from_buffer_copy(b) for b in blobs]
    up = C.byref(infos[rank - 1]) if rank > 0 else None
    down = C.byref(infos[rank + 1]) if rank + 1 < world else None
    status = lib.rpt_frame_connect_peers(frame, up, down)
    if status != 0:
        raise restirpt.RestirptError(f"rpt_frame_connect_peers failed ({status}): "
                                     f"{lib.rpt_last_error(None).decode()}")
    dist.barrier()
    return StripLink(frame, rank, world, "p2p")


def exchange_halo(strip, row_begin, row_end, height, halo, rank, world, group=None):
    """strip: tensor [stored_rows, ...] holding film rows storage_rows(...) of this rank; after the call its halo
    rows hold the neighbours' boundary rows.  Plain send/recv (gloo on CPU, NCCL on CUDA)."""
    import torch.distributed as dist
    s0, _ = storage_rows(row_begin, row_end, height, halo)
    ops = []
    if rank > 0:   # strip above: send my first `halo` owned rows, receive its last `halo` owned rows
        n_up = row_begin - s0
        ops.append(dist.P2POp(dist.isend, strip[row_begin - s0: row_begin - s0 + halo].contiguous(), rank - 1, group))
        recv_up = strip[0:n_up].clone()
        ops.append(dist.P2POp(dist.irecv, recv_up, rank - 1, group))
    if rank + 1 < world:
        n_dn = min(row_end + halo, height) - row_end
        ops.append(dist.P2POp(dist.isend, strip[row_end - halo - s0: row_end - s0].contiguous(), rank + 1, group))
        recv_dn = strip[row_end - s0: row_end - s0 + n_dn].clone()
        ops.append(dist.P2POp(dist.irecv, recv_dn, rank + 1, group))
    for req in dist.batch_isend_irecv(ops) if ops else []:
        req.wait()
    if rank > 0:
        strip[0:row_begin - s0] = recv_up
    if rank + 1 < world:
        strip[row_end - s0: row_end - s0 + n_dn] = recv_dn
    return strip
