"""Multi-GPU partitioning of the hot path: one process per GPU, contiguous row shards, one gather of ciphertext buffers.

Elements of encrypt / decrypt / add / mul are independent, so rank g takes rows
[g*N/G, (g+1)*N/G) of the packed limb matrices and the key material is replicated.  The only collective is the final
gather of the output rows (BASELINE config 4); there is no data-path exchange.  Works with any torch.distributed
backend: `nccl` on the GPUs (all_gather_into_tensor over NVLink), `gloo` in the CPU tests.
"""
from typing import List, Tuple

import torch
import torch.distributed as dist


def shard_bounds(count: int, world: int, rank: int) -> Tuple[int, int]:
    """Rows [lo, hi) of rank `rank`: contiguous, sizes differ by at most one, earlier ranks take the extra rows."""
    if world < 1 or not 0 <= rank < world or count < 0:
        raise ValueError("shard_bounds: bad arguments")
    base, extra = divmod(count, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_sizes(count: int, world: int) -> List[int]:
    return [hi - lo for lo, hi in (shard_bounds(count, world, r) for r in range(world))]


def gather_rows(local: torch.Tensor, count: int, group=None) -> torch.Tensor:
    """All-gather the row shards of a [count, words] matrix.  `local` is this rank's [hi - lo, words] block
    (int32 view of the uint32 limbs).  Returns the full [count, words] matrix on every rank.

    Even shards use one all_gather_into_tensor (NCCL writes peers' rows straight into the result buffer); ragged
    shards pad to the largest shard, gather, and drop the padding."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = shard_sizes(count, world)
    if local.shape[0] != sizes[rank]:
        raise ValueError("gather_rows: rank %d holds %d rows, expected %d" % (rank, local.shape[0], sizes[rank]))
    words = local.shape[1]
    if len(set(sizes)) == 1:
        out = torch.empty((count, words), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
        return out
    big = max(sizes)
    padded = torch.zeros((big, words), dtype=local.dtype, device=local.device)
    padded[: local.shape[0]] = local
    buf = torch.empty((world * big, words), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(buf, padded, group=group)
    return torch.cat([buf[r * big: r * big + sizes[r]] for r in range(world)], dim=0)


def encrypt_sharded(pk, m: torch.Tensor, r: torch.Tensor, stream: int = 0, group=None) -> torch.Tensor:
    """Encrypt the global batch m [N, n_words] with obfuscator exponents r [N, r_words] (both replicated or at least
    valid on this rank's rows, int32 views on the current CUDA device): this rank encrypts its shard with
    phe_encrypt_dev and the ciphertext shards are gathered.  `pk` is a capi.PubKey."""
    count = m.shape[0]
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    lo, hi = shard_bounds(count, world, rank)
    local = torch.empty((hi - lo, 2 * pk.n_words), dtype=torch.int32, device=m.device)
    if hi > lo:
        ms, rs = m[lo:hi].contiguous(), r[lo:hi].contiguous()
        pk.encrypt_dev(ms.data_ptr(), hi - lo, rs.data_ptr(), rs.shape[1], local.data_ptr(), stream)
    return gather_rows(local, count, group)
