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
        pass


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
