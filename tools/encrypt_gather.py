#!/usr/bin/env python
"""BASELINE config 4: batch-N DJN encrypt sharded over the ranks of one node + NCCL all-gather of the ciphertexts.

  torchrun --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/encrypt_gather.py --count 8388608
  python tools/encrypt_gather.py --count 1048576            # single GPU (no collective)

Prints one JSON line from rank 0: encrypt ops/s over the whole job (max over ranks, CUDA events), the gather's share,
and a parity spot-check of rows from every shard against the Python-int oracle formula.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import bench_key, make_workload  # noqa: E402
from pailliercryptolib_python_b200 import capi  # noqa: E402
from pailliercryptolib_python_b200.sharding import gather_rows, shard_bounds  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--count", type=int, default=1 << 20)
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--fused", action="store_true", help="the encrypt kernel stores every row into all ranks' gather buffers (CUDA IPC peer memory over NVLink) instead of an NCCL all-gather afterwards")
    args = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    capi.lib().phe_set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n, p, q, hs = bench_key()
    pk = capi.PubKey(n, 2048, djn=True, hs=hs)
    lo, hi = shard_bounds(args.count, world, rank)
    # every rank generates only its own rows of the global workload (seeded per row block)
    m_np, r_np = make_workload(args.count if args.count <= (1 << 21) else hi - lo, seed=1234 + (0 if args.count <= (1 << 21) else rank))
    if args.count <= (1 << 21):
        m_np, r_np = m_np[lo:hi], r_np[lo:hi]
    m = torch.from_numpy(m_np.view(np.int32)).to(dev)
    r = torch.from_numpy(r_np.view(np.int32)).to(dev)
    local_ct = torch.empty((hi - lo, 128), dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    fused = args.fused and world > 1
    if fused:
        # every rank owns one full gather buffer; the others map it through CUDA IPC and write their rows into it
        own = capi.DeviceBuffer(pk, args.count, 128)
        full_buf = torch.as_tensor(own, device=dev)
        full_buf.zero_()
        torch.cuda.synchronize()
        handles = [None] * world
        dist.all_gather_object(handles, own.ipc_handle())
        peer_ptrs = []
        for r_, h in enumerate(handles):
            if r_ != rank:
                peer_ptrs.append(capi.ipc_open(h) + lo * 512)
        torch.cuda.set_device(local)
        dist.barrier()

    def once():
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0.record()
        if fused:
            pk.encrypt_dev_multi(m.data_ptr(), hi - lo, r.data_ptr(), 32, full_buf.data_ptr() + lo * 512, peer_ptrs, stream)
            e1.record()
            torch.cuda.synchronize()
            dist.barrier()            # every rank's rows have landed in every buffer
            full = full_buf
        else:
            pk.encrypt_dev(m.data_ptr(), hi - lo, r.data_ptr(), 32, local_ct.data_ptr(), stream)
            e1.record()
            full = gather_rows(local_ct, args.count) if world > 1 else local_ct
        e2.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e2), e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return full, float(t[0]), float(t[1])

    once()
    best = None
    for _ in range(args.reps):
        full, total_ms, enc_ms = once()
        if best is None or total_ms < best[0]:
            best = (total_ms, enc_ms)
    # parity spot check on this rank's first and last row (oracle formula, exact ints)
    ok = True
    for row in (0, hi - lo - 1):
        mi = int.from_bytes(m_np[row].tobytes(), "little")
        ri = int.from_bytes(r_np[row].tobytes(), "little")
        want = (1 + mi * n) * pow(hs, ri, n * n) % (n * n)
        got = int.from_bytes(full[lo + row].cpu().numpy().tobytes(), "little")
        ok = ok and (got == want)
    if world > 1:   # every rank must hold the same full matrix: compare a checksum of all rows
        chk = full.to(torch.int64).sum(dim=0)
        ref = chk.clone()
        dist.broadcast(ref, 0)
        ok = ok and bool(torch.equal(chk, ref))
    flag = torch.tensor([1 if ok else 0], device=dev)
    if world > 1:
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(json.dumps({"workload": "2048-bit DJN encrypt, batch=%d sharded over %d GPU(s) + %s" % (args.count, world, "rows stored by the encrypt kernel into every rank's gather buffer (CUDA IPC peer memory)" if fused else "all-gather of ciphertexts"),
                          "fused_peer_stores": bool(fused),
                          "n_gpus": world, "encrypt_ops_s": args.count / (best[0] * 1e-3), "ms_total": best[0], "ms_encrypt": best[1],
                          "gather_share": (best[0] - best[1]) / best[0], "gathered_bytes": args.count * 512,
                          "parity_spot_check": bool(flag.item())}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
