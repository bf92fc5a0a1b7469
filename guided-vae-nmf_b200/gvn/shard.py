"""Utterance-level data parallelism (the only parallelism of the path).

The reference splits its file list with ``np.array_split(file_paths, nb_devices)`` and starts one
process per GPU (scripts/evaluate_M1.py:203-216); results leave the workers as files.  Here the
split is the same (contiguous shards, the first ``n % world`` ranks get one extra utterance) and
the per-utterance result rows are brought together with one all-gather (NCCL on GPUs, gloo in
the CPU tests).  Nothing else is communicated: utterances are independent problems.
"""
import torch
import torch.distributed as dist


def shard_bounds(n, world, rank):
    """[start, stop) of rank's contiguous shard, np.array_split semantics."""
    base, extra = divmod(n, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_list(items, world, rank, strided=False):
    """Rank's share of a list.  Default: the reference's contiguous np.array_split shards.  ``strided``: for a list
    that has been SORTED BY LENGTH -- dealt like cards in a snake (0 1 .. w-1, w-1 .. 1 0, ...), so that every rank
    gets the same mix of long and short utterances: a contiguous shard of a sorted list would hand all the long ones
    to one rank, and with utterances of 537-748 frames the slowest rank sets the time of the whole job."""
    if not strided:
        a, b = shard_bounds(len(items), world, rank)
        return items[a:b]
    out = []
    for i, it in enumerate(items):
        rnd, pos = divmod(i, world)
        if (pos if rnd % 2 == 0 else world - 1 - pos) == rank:
            out.append(it)
    return out


def length_sorted_shard(lengths, world, rank):
    """Indices of rank's utterances: longest first, dealt in a snake (see shard_list).  Batches cut from this list
    hold utterances of similar length."""
    order = sorted(range(len(lengths)), key=lambda i: (-int(lengths[i]), i))
    return shard_list(order, world, rank, strided=True)


def chain_tiles(n_frames, align=32, tile=128):
    """Number of 128-frame tiles the chain kernel launches for a batch of utterances with these frame counts (every
    utterance starts on a multiple of `align` frames of the global frame axis, gvn.engine.Batch)."""
    return (sum((int(n) + align - 1) // align * align for n in n_frames) + tile - 1) // tile


def wave_batches(n_frames, sms, waves=2, max_batch=256, align=32, tile=128):
    """Cuts an ordered list of utterances (frame counts `n_frames`) into consecutive batches sized for the chain
    kernel: one CTA per 128-frame tile, one CTA per SM, every tile of a launch takes the same time -- so a launch costs
    ceil(tiles / sms) waves whatever the fill of its last wave.  A batch of 64 utterances of ~650 frames is 325 tiles =
    2.2 waves on 148 SMs and pays for 3; here a batch takes as many utterances as fit `waves` full waves (at least one).
    Returns a list of (start, stop) index pairs."""
    budget = waves * sms * tile
    out, start, used = [], 0, 0
    for i, n in enumerate(n_frames):
        a = (int(n) + align - 1) // align * align
        if i > start and (used + a > budget or i - start >= max_batch):
            out.append((start, i))
            start, used = i, 0
        used += a
    if len(n_frames) > start:
        out.append((start, len(n_frames)))
    return out


def gather_rows(rows, n_total, world=None, rank=None):
    """rows: (n_local, C) tensor of this rank's result rows (first column = utterance id).
    Returns the (n_total, C) table ordered by utterance id on every rank."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return rows
    world = dist.get_world_size() if world is None else world
    rank = dist.get_rank() if rank is None else rank
    cap = max(shard_bounds(n_total, world, r)[1] - shard_bounds(n_total, world, r)[0] for r in range(world))
    pad = torch.full((cap, rows.shape[1]), -1.0, dtype=rows.dtype, device=rows.device)
    pad[:rows.shape[0]] = rows
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad)
    table = torch.cat(out, 0)
    table = table[table[:, 0] >= 0]
    return table[torch.argsort(table[:, 0])]
