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
