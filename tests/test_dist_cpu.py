"""N>1 path on CPU (gloo, world_size 2): byte-range shards, one per rank, no data-path
collective; rank 0 folds the per-rank NAL tables with the product's host merge
(h264gpu_merge_*).  Each rank runs the kernel SOURCE on the SIMT emulator (no GPU here);
gloo only carries the results to rank 0 for the check, as bench.py's barrier/timing does."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

import support as S

torch = pytest.importorskip("torch")
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _shard_bounds(n, world):
    per = (n // world + 15) // 16 * 16          # 16-byte aligned cuts (bulk-copy alignment)
    return [min(r * per, n) for r in range(world)] + [n]


def _worker(rank, world, port, stream, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import libh264_b200 as L
    import support as S2
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    b = np.frombuffer(stream, np.uint8)
    n = len(b)
    bounds = _shard_bounds(n, world)
    lo, hi = bounds[rank], bounds[rank + 1]
    e = S2.ShardEdge()
    if lo >= 2:
        e.has_left, e.left[0], e.left[1] = 1, int(b[lo - 2]), int(b[lo - 1])
    if hi < n:
        e.has_right = 1
        e.right[0] = int(b[hi])
        e.right[1] = int(b[hi + 1]) if hi + 1 < n else 0xFF
    e.assume_in = 1 if lo > 0 else 0
    r = S2.emu_split_strip(b[lo:hi], strip=True, items=1, edge=e, base=lo)
    k = int(r["res"].n_nal)
    # gather variable-length results on rank 0 (padded tensors + lengths)
    cap = n // 3 + 2
    pad = lambda a, m: torch.from_numpy(np.concatenate([a.astype(np.int64), np.zeros(m - len(a), np.int64)]))
    mine = [pad(r["start"], cap), pad(r["end"].astype(np.uint64).view(np.int64), cap), pad(r["rbsp_off"], cap),
            torch.from_numpy(np.concatenate([r["rbsp"], np.zeros(n - len(r["rbsp"]), np.uint8)]).astype(np.uint8)),
            torch.from_numpy(np.frombuffer(bytes(r["res"]), np.uint8).copy()),
            torch.tensor([k, len(r["rbsp"])], dtype=torch.int64)]
    gathered = []
    for t in mine:
        out = [torch.zeros_like(t) for _ in range(world)] if rank == 0 else None
        dist.gather(t, out, dst=0)
        gathered.append(out)
    if rank == 0:
        lib = L.load_gpu_lib()
        st = np.zeros(cap, np.uint64)
        en = np.zeros(cap, np.uint64)
        ro = np.zeros(cap, np.uint64)
        parts = []
        m = L.Merge()
        lib.h264gpu_merge_init(C.byref(m))
        for rr in range(world):
            kk, nb = (int(x) for x in gathered[5][rr])
            st[m.n_nal:m.n_nal + kk] = gathered[0][rr].numpy()[:kk].astype(np.uint64)
            en[m.n_nal:m.n_nal + kk] = gathered[1][rr].numpy()[:kk].view(np.uint64)
            ro[m.n_nal:m.n_nal + kk] = gathered[2][rr].numpy()[:kk].astype(np.uint64)
            res = L.ScanResult.from_buffer_copy(gathered[4][rr].numpy().tobytes())
            skip, take = C.c_uint64(0), C.c_uint64(0)
            lib.h264gpu_merge_shard(C.byref(m), C.byref(res), S2.ptr(st), S2.ptr(en), S2.ptr(ro), cap, kk,
                                    C.byref(skip), C.byref(take))
            parts.append(gathered[3][rr].numpy()[:nb][skip.value:skip.value + take.value])
        off = C.c_uint64(0)
        lib.h264gpu_merge_finish(C.byref(m), n, S2.ptr(en), cap, C.byref(off))
        k = int(m.n_nal)
        q.put((st[:k].tobytes(), en[:k].tobytes(), ro[:k].tobytes(), np.concatenate(parts).tobytes(), off.value))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("kind", ["valid", "adversarial"])
def test_two_rank_byte_range_shards_merge_to_whole(kind):
    rng = np.random.default_rng(21)
    if kind == "valid":
        b = S.gen_annexb(rng, 40, 1, 4000)
    else:
        b = rng.choice(np.array([0, 0, 0, 1, 2, 3, 0xFF, 0x65], np.uint8), 30001)
    S.build_emu()
    o = S.oracle_split_strip(b)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + int(rng.integers(0, 2000))
    procs = [ctx.Process(target=_worker, args=(r, 2, port, b.tobytes(), q)) for r in range(2)]
    for p in procs:
        p.start()
    st, en, ro, rbsp, off = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert np.array_equal(np.frombuffer(st, np.uint64), o["start"])
    assert np.array_equal(np.frombuffer(en, np.uint64), o["end"])
    assert np.array_equal(np.frombuffer(ro, np.uint64), o["rbsp_off"][:-1])
    assert np.array_equal(np.frombuffer(rbsp, np.uint8), o["rbsp"])
    assert off == o["final_off"]


# ---- the same with the in-place kernel (gen 5, what bench.py --gpus N runs per rank) -----------

def _worker_inplace(rank, world, port, stream, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import libh264_b200 as L
    import support as S2
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    b = np.frombuffer(stream, np.uint8)
    n = len(b)
    bounds = _shard_bounds(n, world)
    lo, hi = bounds[rank], bounds[rank + 1]
    e = S2.ShardEdge()
    if lo >= 2:
        e.has_left, e.left[0], e.left[1] = 1, int(b[lo - 2]), int(b[lo - 1])
    if hi < n:
        e.has_right = 1
        e.right[0] = int(b[hi])
        e.right[1] = int(b[hi + 1]) if hi + 1 < n else 0xFF
    e.assume_in = 1 if lo > 0 else 0
    r = S2.emu_split_strip_inplace(b[lo:hi], cpt=1, edge=e, base=lo)
    k = int(r["res"].n_nal)
    cap = n // 3 + 2
    pad = lambda a, m: torch.from_numpy(np.concatenate([a.astype(np.uint64).view(np.int64), np.zeros(m - len(a), np.int64)]))
    body = np.zeros(n, np.uint8)
    body[:hi - lo] = r["body"]
    mine = [pad(r["start"], cap), pad(r["end"], cap), pad(r["rbsp_off"], cap), pad(r["rbsp_len"], cap),
            torch.from_numpy(body), torch.from_numpy(np.frombuffer(bytes(r["res"]), np.uint8).copy()),
            torch.tensor([k, lo], dtype=torch.int64)]
    gathered = []
    for t in mine:
        out = [torch.zeros_like(t) for _ in range(world)] if rank == 0 else None
        dist.gather(t, out, dst=0)
        gathered.append(out)
    if rank == 0:
        lib = L.load_gpu_lib()
        st, en, rl = np.zeros(cap, np.uint64), np.zeros(cap, np.uint64), np.zeros(cap, np.uint64)
        pieces = [[] for _ in range(cap)]
        m = L.Merge()
        lib.h264gpu_merge_init(C.byref(m))
        for rr in range(world):
            kk, base = (int(x) for x in gathered[6][rr])
            first = int(m.n_nal)
            st[first:first + kk] = gathered[0][rr].numpy()[:kk].view(np.uint64)
            en[first:first + kk] = gathered[1][rr].numpy()[:kk].view(np.uint64)
            ro = gathered[2][rr].numpy()[:kk].view(np.uint64)
            rl[first:first + kk] = gathered[3][rr].numpy()[:kk].view(np.uint64)
            shard = gathered[4][rr].numpy()
            for i in range(kk):
                pieces[first + i].append(shard[int(ro[i]):int(ro[i]) + int(rl[first + i])])
            res = L.ScanResult.from_buffer_copy(gathered[5][rr].numpy().tobytes())
            carry = C.c_uint64(0)
            lib.h264gpu_merge_shard_inplace(C.byref(m), C.byref(res), S2.ptr(en), S2.ptr(rl), cap, C.byref(carry))
            if carry.value:
                at = max(0, int(st[first - 1]) - base)
                pieces[first - 1].append(shard[at:at + carry.value])
        off = C.c_uint64(0)
        lib.h264gpu_merge_finish(C.byref(m), n, S2.ptr(en), cap, C.byref(off))
        k = int(m.n_nal)
        rbsp = np.concatenate([np.concatenate(p) if p else np.zeros(0, np.uint8) for p in pieces[:k]] +
                              [np.zeros(0, np.uint8)])
        q.put((st[:k].tobytes(), en[:k].tobytes(), rl[:k].tobytes(), rbsp.tobytes(), off.value))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("kind", ["valid", "adversarial"])
def test_two_rank_inplace_shards_merge_to_whole(kind):
    rng = np.random.default_rng(22)
    if kind == "valid":
        b = S.gen_annexb(rng, 40, 1, 4000)
    else:
        b = rng.choice(np.array([0, 0, 0, 1, 2, 3, 0xFF, 0x65], np.uint8), 30001)
    S.build_emu()
    o = S.oracle_split_strip(b)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + int(rng.integers(0, 2000))
    procs = [ctx.Process(target=_worker_inplace, args=(r, 2, port, b.tobytes(), q)) for r in range(2)]
    for p in procs:
        p.start()
    st, en, rl, rbsp, off = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert np.array_equal(np.frombuffer(st, np.uint64), o["start"])
    assert np.array_equal(np.frombuffer(en, np.uint64), o["end"])
    assert np.array_equal(np.frombuffer(rl, np.uint64), np.diff(o["rbsp_off"]).astype(np.uint64))
    assert np.array_equal(np.frombuffer(rbsp, np.uint8), o["rbsp"])
    assert off == o["final_off"]
