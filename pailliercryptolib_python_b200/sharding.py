"""Multi-GPU partitioning of the hot path: one process per GPU, contiguous row shards, one gather of ciphertext buffers.

Elements of encrypt / decrypt / add / mul are independent, so rank g takes rows
[g*N/G, (g+1)*N/G) of the packed limb matrices and the key material is replicated.  The only collective is the final
gather of the output rows (BASELINE config 4); there is no data-path exchange.  Works with any torch.distributed
backend: `nccl` on the GPUs (all_gather_into_tensor over NVLink), `gloo` in the CPU tests.
"""
from typing import List, Optional, Tuple

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


def encrypt_sharded(pk, m: torch.Tensor, r: Optional[torch.Tensor] = None, stream: int = 0, group=None,
                    make_secure: bool = True) -> torch.Tensor:
    """Encrypt the global batch m [N, n_words] (replicated, or at least valid on this rank's rows; int32 view on the
    current CUDA device): this rank encrypts its shard with phe_encrypt_dev and the ciphertext shards are gathered.
    r = None (the default): the library draws the obfuscator exponents itself (device ChaCha20 keyed from the OS CSPRNG).
    An explicit r [N, r_words] is the deterministic hook of the parity tests and MUST come from a CSPRNG otherwise.
    `pk` is a capi.PubKey."""
    count = m.shape[0]
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    lo, hi = shard_bounds(count, world, rank)
    local = torch.empty((hi - lo, 2 * pk.n_words), dtype=torch.int32, device=m.device)
    if hi > lo:
        ms = m[lo:hi].contiguous()
        rs = r[lo:hi].contiguous() if r is not None else None
        pk.encrypt_dev(ms.data_ptr(), hi - lo, rs.data_ptr() if rs is not None else None,
                       rs.shape[1] if rs is not None else 0, local.data_ptr(), stream, make_secure=make_secure)
    return gather_rows(local, count, group)


class PeerGather:
    """The fused form of the gather (BASELINE config 4): every rank owns one full [count, words] buffer
    (capi.DeviceBuffer, a whole cudaMalloc block), exports it through CUDA IPC, and maps the other ranks' buffers; the
    encrypt kernel then stores each ciphertext row into all of them itself (phe_encrypt_dev_multi: plain st.global to
    peer-mapped addresses over NVLink, overlapped with the arithmetic), so no collective follows -- one barrier.
    NCCL ranks of one node only."""

    def __init__(self, pk, count: int, device: torch.device, group=None):
        from . import capi
        self.pk, self.count, self.group = pk, count, group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.lo, self.hi = shard_bounds(count, self.world, self.rank)
        self.words = 2 * pk.n_words
        self.own = capi.DeviceBuffer(pk, count, self.words)
        self.full = torch.as_tensor(self.own, device=device)
        handles = [None] * self.world
        dist.all_gather_object(handles, self.own.ipc_handle(), group=group)
        self.peer_ptrs = [capi.ipc_open(h) + self.lo * self.words * 4 for r_, h in enumerate(handles) if r_ != self.rank]
        torch.cuda.set_device(device)
        dist.barrier(group=group)

    def encrypt(self, m_local: torch.Tensor, r_local: Optional[torch.Tensor], stream: int = 0, make_secure: bool = True):
        """m_local / r_local: this rank's rows [hi - lo, ...].  Enqueues only; call finish() before reading `full`."""
        self.pk.encrypt_dev_multi(m_local.data_ptr(), self.hi - self.lo,
                                  r_local.data_ptr() if r_local is not None else None,
                                  r_local.shape[1] if r_local is not None else 0,
                                  self.full.data_ptr() + self.lo * self.words * 4, self.peer_ptrs, stream,
                                  make_secure=make_secure)

    def finish(self) -> torch.Tensor:
        torch.cuda.synchronize()
        dist.barrier(group=self.group)     # every rank's rows have landed in every buffer
        return self.full
