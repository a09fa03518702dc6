"""Host-side sharding of a batch across the GPUs of one box.

AMV frames are intra-only and every ADPCM chunk carries its own decoder state, so the path
partitions into independent units: GPU g owns the contiguous unit range shard_range(n, g, G) and
runs the ordinary single-GPU entry points on it.  Nothing crosses GPUs on the data path; the only
communication is this module's gather of the per-rank (offset, size) packet tables -- metadata,
a few bytes per frame, over whatever torch.distributed backend the job uses (NCCL on the GPU box,
gloo in the CPU tests) -- and, for ONE continuous audio stream split across ranks, the 2-byte
encoder state handed from the end of one rank's range to the start of the next (adpcm.c:466).
The audio resampler in front of the encoder shards by OUTPUT range: every output sample depends on a
short window of the input only, so resample_shard gives each rank its outputs and the slice of the
stream they read (neighbouring slices overlap by one filter length; nothing is exchanged).
"""
import numpy as np


def shard_range(n, rank, world):
    """Contiguous, balanced [lo, hi) of `n` units for `rank` of `world` (first n % world ranks get one more)."""
    base, extra = divmod(int(n), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_packet_table(local_sizes, dist=None, device="cpu"):
    """All ranks contribute the sizes of the packets they produced for their own frame range;
    every rank gets back (global_sizes, global_offsets, my_base): the table of the concatenated
    stream in frame order and the byte offset at which this rank's packets start in it."""
    import torch
    local_sizes = np.ascontiguousarray(local_sizes, dtype=np.int64)
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        sizes = local_sizes
        rank_counts = [len(local_sizes)]
        rank = 0
    else:
        world, rank = dist.get_world_size(), dist.get_rank()
        cnt = torch.tensor([len(local_sizes)], dtype=torch.int64, device=device)
        counts = [torch.zeros_like(cnt) for _ in range(world)]
        dist.all_gather(counts, cnt)
        rank_counts = [int(c.item()) for c in counts]
        m = max(rank_counts) if rank_counts else 0
        pad = torch.zeros(max(m, 1), dtype=torch.int64, device=device)
        pad[: len(local_sizes)] = torch.from_numpy(local_sizes).to(device)
        allp = [torch.zeros_like(pad) for _ in range(world)]
        dist.all_gather(allp, pad)
        sizes = np.concatenate([allp[r][: rank_counts[r]].cpu().numpy() for r in range(world)])
    offsets = np.zeros(len(sizes), dtype=np.int64)
    if len(sizes) > 1:
        offsets[1:] = np.cumsum(sizes)[:-1]
    first = int(sum(rank_counts[:rank]))
    my_base = int(offsets[first]) if first < len(sizes) else int(sizes.sum())
    return sizes.astype(np.uint32), offsets.astype(np.uint64), my_base


def chain_stream_state(encode_range, dist=None, initial_state=0, device="cpu"):
    """Encode ONE continuous ADPCM stream whose chunks are sharded contiguously across ranks.
    `encode_range(step_in) -> step_out` encodes this rank's chunk range starting from encoder
    state step_in and returns the state after its last chunk.  Ranks run in order, each receiving
    the previous rank's final state (2 bytes) -- the only serial dependency the codec has."""
    import torch
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return encode_range(int(initial_state))
    world, rank = dist.get_world_size(), dist.get_rank()
    state = torch.tensor([int(initial_state)], dtype=torch.int64, device=device)
    if rank > 0:
        dist.recv(state, src=rank - 1)
    out = encode_range(int(state.item()))
    if rank < world - 1:
        dist.send(torch.tensor([int(out)], dtype=torch.int64, device=device), dst=rank + 1)
    return out


def resample_shard(n_in, in_rate, out_rate, rank, world, lib):
    """ONE stream of n_in samples per channel resampled by `world` ranks: rank r computes outputs
    [k_start, k_start + k_count) -- a contiguous, balanced share of the amv_audio_resample_count(n_in) outputs --
    from the input samples [in_base, in_base + n_window), and calls
    amv_audio_resample_from(ctx, pcm + in_base * channels, in_base, n_window, ..., k_start, out, k_count, ...).
    The window ends where the rank's last output's taps end, so the call yields exactly k_count samples.
    `lib` is the loaded libamvcuda (the two helpers used here are host arithmetic, no device involved).
    Returns (k_start, k_count, in_base, n_window)."""
    total = int(lib.amv_audio_resample_count(int(n_in), int(in_rate), int(out_rate)))
    k_lo, k_hi = shard_range(total, rank, world)
    if k_hi == k_lo:
        return k_lo, 0, 0, 0
    flen = int(lib.amv_audio_resample_bank(int(in_rate), int(out_rate), None, 0))
    first = int(lib.amv_audio_resample_first_tap(k_lo, int(in_rate), int(out_rate)))
    in_base = max(first, 0)                       # the mirrored head of the stream (first < 0) belongs to sample 0
    # smallest stream length that still yields output k_hi - 1: its taps end at first_tap(k_hi - 1) + filter length;
    # the mirrored head needs no such room (its outputs exist however short the stream is)
    last = int(lib.amv_audio_resample_first_tap(k_hi - 1, int(in_rate), int(out_rate)))
    end = int(n_in) if last < 0 else min(int(n_in), last + flen)
    if rank == world - 1:
        end = int(n_in)
    return k_lo, k_hi - k_lo, in_base, end - in_base
