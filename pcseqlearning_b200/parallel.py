"""Host-side plumbing for running the path on several GPUs of one box (one process per GPU, torch.distributed).

Two partitionings:
  * replicas: whole sequences round-robin over ranks, exactly like the reference's DistributedSampler over
    sequences (pcdet/datasets/__init__.py:84-91) -- no data-path collective;
  * frame windows of ONE sequence: contiguous windows aligned to lcm(10-frame proposal chunk, 8-frame tracking
    interval) = 40 frames, a +-TRACK_INTERVAL halo for tracking, and an all-gather of the per-chunk component counts
    to turn window-local component ids into the reference's sequence-global numbering
    (cluster_proposal.py:63-81: running offset over chunks).
Everything here is device agnostic (gloo on CPU tensors in the tests, NCCL on the GPU box).
"""
import math

import torch
import torch.distributed as dist

CHUNK_FRAMES = 10
TRACK_INTERVAL = 8


def sequences_of_rank(num_sequences, rank, world):
    """Round-robin sequence assignment (DistributedSampler without shuffling / padding)."""
    return list(range(rank, num_sequences, world))


def frame_windows(num_frames, world, align=None):
    """Contiguous [start, end) frame windows, one per rank, boundaries on multiples of `align`
    (default lcm(CHUNK_FRAMES, TRACK_INTERVAL) = 40).  Ranks beyond the number of aligned blocks get empty windows."""
    if align is None:
        align = CHUNK_FRAMES * TRACK_INTERVAL // math.gcd(CHUNK_FRAMES, TRACK_INTERVAL)
    blocks = (num_frames + align - 1) // align
    out = []
    for r in range(world):
        b0 = (blocks * r) // world
        b1 = (blocks * (r + 1)) // world
        out.append((min(b0 * align, num_frames), min(b1 * align, num_frames)))
    return out


def halo_window(window, num_frames, halo=TRACK_INTERVAL):
    """Frames a rank must hold to track the anchors of its window: [start - halo, end + halo) clipped."""
    s, e = window
    if e <= s:
        return (s, s)
    return (max(0, s - halo), min(num_frames, e + halo))


def anchors_of_window(window, interval=TRACK_INTERVAL):
    """Tracking anchors (frames = 0 mod interval) owned by a window."""
    s, e = window
    first = ((s + interval - 1) // interval) * interval
    return list(range(first, e, interval))


def globalize_component_ids(labels_local, frame_of_point, n_comp_local, window, num_frames, group=None):
    """Window-local component ids -> sequence-global ids.

    labels_local int64[n]: ids numbered per chunk with a running offset over the window's chunks (what
    ops.cluster_labels returns for the window's points, frames re-based to the window start or not);
    n_comp_local int64[n_chunks_local]: components per local chunk; frame_of_point: absolute frame of every point.
    All ranks call this collectively.  Returns (labels_global, n_comp_all int64[total chunks]).
    """
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    total_chunks = (num_frames + CHUNK_FRAMES - 1) // CHUNK_FRAMES
    counts = torch.zeros(total_chunks, dtype=torch.int64, device=labels_local.device)
    c0 = window[0] // CHUNK_FRAMES
    counts[c0:c0 + n_comp_local.shape[0]] = n_comp_local.to(counts)
    if world > 1:
        dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=group)  # windows are disjoint: sum == gather
    global_off = torch.cumsum(counts, 0) - counts
    local_off = torch.cumsum(n_comp_local, 0) - n_comp_local
    chunk = (frame_of_point.long() // CHUNK_FRAMES) - c0
    if labels_local.numel() == 0:
        return labels_local, counts
    labels_global = labels_local - local_off.to(labels_local)[chunk] + global_off[c0 + chunk]
    return labels_global, counts


def max_over_ranks(value, device, group=None):
    """Timing reduction of the bench contract: the slowest rank defines the step time."""
    t = torch.tensor([float(value)], device=device)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


# ---------------------------------------------------------------------------------------------------------------
# Frame-window sharding of ONE sequence (BASELINE.json configs[3]): the stages consult `SHARD` (None = single GPU)
# ---------------------------------------------------------------------------------------------------------------
def chunk_windows(num_frames, world, chunk=CHUNK_FRAMES):
    """Contiguous [start, end) frame windows on proposal-chunk boundaries, balanced by chunk count: components never
    span frames and are numbered per 10-frame chunk (cluster_proposal.py:63-81), so a chunk must stay on one rank."""
    n_chunks = (num_frames + chunk - 1) // chunk
    out = []
    for r in range(world):
        c0, c1 = (n_chunks * r) // world, (n_chunks * (r + 1)) // world
        out.append((min(c0 * chunk, num_frames), min(c1 * chunk, num_frames)))
    return out


def anchor_blocks(num_frames, world, interval=TRACK_INTERVAL):
    """Tracking anchors (frames = 0 mod interval) dealt to the ranks in contiguous, balanced blocks."""
    anchors = list(range(0, num_frames, interval))
    return [anchors[(len(anchors) * r) // world:(len(anchors) * (r + 1)) // world] for r in range(world)]


class FrameSharding:
    """Communication context of one sequence sharded by frame windows over the ranks of a process group.

    NCCL (over NVLink / NVSwitch) on the GPU box, gloo on CPU tensors in the tests.  Collectives used:
      all_reduce(MIN/MAX) of the voxel-grid bounds (32 bytes), all_gather of the ground stage's partial voxel sums,
      all_gather of the per-ratio best planes, all_reduce(SUM) of per-chunk component counts, all_to_all of the halo
      frames (+-TRACK_INTERVAL around every rank's anchor block), all_reduce(MAX) of the per-box best IoU."""

    def __init__(self, num_frames, group=None, interval=TRACK_INTERVAL):
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.F = int(num_frames)
        self.interval = int(interval)
        self.windows = chunk_windows(self.F, self.world)
        self.window = self.windows[self.rank]
        self.blocks = anchor_blocks(self.F, self.world, interval)
        self.anchors = self.blocks[self.rank]
        # frames every rank needs for tracking: [first anchor - interval, last anchor + interval]
        self.need = [((max(0, b[0] - interval), min(self.F, b[-1] + interval + 1)) if b else (0, 0)) for b in self.blocks]

    # -- small reductions ------------------------------------------------------------------------------------
    def reduce_bounds(self, bounds):
        """bounds: int32 view of order-preserving uint32 encodings, [..., 8] = (min x4, max x4); reduced in place."""
        b = bounds.view(torch.int32).reshape(-1, 8).long() & 0xffffffff
        lo, hi = b[:, :4].contiguous(), b[:, 4:].contiguous()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN, group=self.group)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX, group=self.group)
        merged = torch.cat([lo, hi], 1)
        merged = torch.where(merged >= 2 ** 31, merged - 2 ** 32, merged).int()
        bounds.view(torch.int32).reshape(-1, 8).copy_(merged)
        return bounds

    def all_reduce_max(self, t):
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        return t

    def all_reduce_sum(self, t):
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t

    def all_gather_v(self, t):
        """Variable-length all_gather along dim 0 -> (concatenated tensor in rank order, sizes list)."""
        n = torch.tensor([t.shape[0]], dtype=torch.int64, device=t.device)
        sizes = [torch.zeros_like(n) for _ in range(self.world)]
        dist.all_gather(sizes, n, group=self.group)
        sizes = [int(s.item()) for s in sizes]
        mx = max(sizes + [1])
        pad = torch.zeros((mx,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        pad[:t.shape[0]] = t
        out = [torch.empty_like(pad) for _ in range(self.world)]
        dist.all_gather(out, pad, group=self.group)
        return torch.cat([o[:s] for o, s in zip(out, sizes)], 0), sizes

    def all_gather(self, t):
        out = [torch.empty_like(t) for _ in range(self.world)]
        dist.all_gather(out, t.contiguous(), group=self.group)
        return out

    # -- halo exchange ------------------------------------------------------------------------------------
    def exchange_frames(self, tensors, frame):
        """Every rank receives the rows of the frames in its `need` window from the ranks that own them.

        tensors: dict name -> tensor [n, ...] of this rank's window; frame int64[n] (absolute frame of every row).
        Rows are shipped frame-sorted (stable), so the result holds the needed frames in ascending order with the
        owner's row order inside a frame -- exactly the rows a single GPU would select.  Returns (dict, frame)."""
        dev = frame.device
        frame = frame.reshape(-1).long()
        order = torch.argsort(frame, stable=True)
        fs = frame[order]
        edges = torch.tensor([x for lo_hi in self.need for x in lo_hi], dtype=torch.int64, device=dev)
        pos = torch.searchsorted(fs, edges).reshape(self.world, 2)
        send_rows = [order[int(pos[r, 0]):int(pos[r, 1])] for r in range(self.world)]
        send_cnt = torch.tensor([int(s.shape[0]) for s in send_rows], dtype=torch.int64, device=dev)
        cnt_all = torch.stack(self.all_gather(send_cnt))  # [src, dst]
        recv_cnt = cnt_all[:, self.rank].tolist()
        send_idx = torch.cat(send_rows) if send_rows else order[:0]
        out = {}
        for name, t in list(tensors.items()) + [("__frame__", frame)]:
            src = t[send_idx].contiguous()
            dst = torch.empty((sum(recv_cnt),) + tuple(t.shape[1:]), dtype=t.dtype, device=dev)
            if t.dtype == torch.bool:
                s8, d8 = src.to(torch.uint8), dst.to(torch.uint8)
                dist.all_to_all_single(d8, s8, recv_cnt, send_cnt.tolist(), group=self.group)
                dst = d8.bool()
            else:
                dist.all_to_all_single(dst, src, recv_cnt, send_cnt.tolist(), group=self.group)
            out[name] = dst
        return out, out.pop("__frame__")


SHARD = None  # the active FrameSharding (set by set_sharding); None = single GPU / replicas


def set_sharding(shard):
    global SHARD
    SHARD = shard
    return shard
