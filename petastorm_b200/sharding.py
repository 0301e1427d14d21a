"""Row-group sharding across the GPUs of one box.

The reference shards by letting every process evaluate ``index % shard_count == cur_shard`` on its own listing of the
dataset (petastorm/reader.py:573-597) - no communication at all.  Here rank 0 builds the owner table once and a single
broadcast (NCCL over NVLink when the process group is NCCL, gloo in the CPU tests) hands it to every rank, so all ranks
agree even if their directory listings differ; after that there is no steady-state traffic: each GPU pulls its own
row-groups over its own PCIe link.
"""
import os
import random

import numpy as np


def owner_table(num_row_groups, shard_count, seed=None):
    """int32[num_row_groups]: rank owning each row-group under the reference's rule (the seed permutes the visiting
    order only, membership is ``index % shard_count``)."""
    owners = np.arange(num_row_groups, dtype=np.int64) % shard_count
    return owners.astype(np.int32)


def shard_order(num_row_groups, shard_count, cur_shard, seed=None):
    """Row-group indexes of one shard in the order the reference would visit them."""
    indexes = list(range(num_row_groups))
    if seed is not None:
        random.Random(seed).shuffle(indexes)
    return [i for i in indexes if i % shard_count == cur_shard]


def broadcast_row_group_assignment(num_row_groups, seed=None, device=None, group=None, listing_digest=0):
    """Collective: returns ``(cur_shard, shard_count, owners)`` where ``owners`` is the int32 owner table computed on
    rank 0 and broadcast once.  Without an initialised process group this is the single-process identity.

    The table broadcast is preceded by a 16-byte header broadcast ``[num_row_groups, listing_digest]`` of rank 0: a rank
    whose own listing differs raises *before* the table broadcast (a broadcast with mismatched sizes would hang or
    corrupt memory under NCCL)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return 0, 1, owner_table(num_row_groups, 1, seed)
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    backend = dist.get_backend(group)
    dev = torch.device('cuda', torch.cuda.current_device()) if backend == 'nccl' else torch.device('cpu')
    if device is not None:
        dev = torch.device(device)
    header = torch.tensor([int(num_row_groups), int(listing_digest)], dtype=torch.int64, device=dev)
    dist.broadcast(header, src=0, group=group)
    n0, digest0 = (int(x) for x in header.cpu().tolist())
    mismatch = n0 != int(num_row_groups) or digest0 != int(listing_digest)
    # every rank learns whether any rank disagrees, so that all of them raise instead of some waiting in the broadcast
    flag = torch.tensor([1 if mismatch else 0], dtype=torch.int32, device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MAX, group=group)
    if int(flag.item()):
        raise RuntimeError('rank {} lists {} row-groups (digest {}) but rank 0 lists {} (digest {}): the ranks do not '
                           'see the same dataset'.format(rank, num_row_groups, listing_digest, n0, digest0)
                           if mismatch else 'another rank does not see the same dataset listing as rank 0')
    if rank == 0:
        table = torch.from_numpy(owner_table(num_row_groups, world, seed)).to(dev)
    else:
        table = torch.empty(num_row_groups, dtype=torch.int32, device=dev)
    dist.broadcast(table, src=0, group=group)
    owners = table.cpu().numpy()
    return rank, world, owners


def sharded_reader_kwargs(dataset_url, seed=None, group=None):
    """``cur_shard`` / ``shard_count`` keyword arguments for make_reader / make_batch_reader, agreed on by all ranks
    through one broadcast; raises if the local listing disagrees with rank 0's."""
    from petastorm_b200.etl import dataset_metadata as dm
    from petastorm_b200.fs_utils import get_filesystem_and_path_or_paths, normalize_dataset_url_or_urls
    _, path = get_filesystem_and_path_or_paths(normalize_dataset_url_or_urls(dataset_url))
    import hashlib
    pieces = dm.load_row_groups(dm.ParquetDataset(path))
    n = len(pieces)
    h = hashlib.md5()
    for piece in pieces:
        h.update('{}:{}\n'.format(os.path.basename(piece.path), piece.row_group).encode('utf-8'))
    digest = int.from_bytes(h.digest()[:7], 'little')
    rank, world, owners = broadcast_row_group_assignment(n, seed, group=group, listing_digest=digest)
    if world == 1:
        return {}
    mine = [i for i in range(n) if owners[i] == rank]
    if mine != sorted(shard_order(n, world, rank, None)):
        raise RuntimeError('row-group assignment mismatch on rank {}'.format(rank))
    return {'cur_shard': rank, 'shard_count': world}
