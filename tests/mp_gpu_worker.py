"""Multi-GPU worker, launched by tests/test_gpu_multi.py (and by hand) as
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tests/mp_gpu_worker.py
One process per GPU over NCCL.  Two checks, both against single-shard ground truth (tier T4):

  A. transport content parity: every rank ingests its own batches of group sends; wire batches travel through the
     REAL cross-process transport (default: CUDA-IPC peer memory pulled by the import/fan-out kernels; `prefetch`:
     the same with the next step's flag wait + descriptor pull running ahead on the prefetch stream; `nccl`:
     all-gather), alternating the two export buffers; every shard drains its agents and folds per-agent stream
     digests on the device; the digests of all shards together must equal oracle/cpu_ref.c's digests of ONE queue
     fed the concatenated rank-major batches (per-agent order, header fields, payload bytes).
  B. the drop-in front-end: `ShardedSwarmsDB` with its DEFAULT exchange (nothing injected) driven by the script of
     tests/test_sharded_frontend_cpu.py; streams must equal a single-process `SwarmsDB` on one GPU.

Prints `MP_GPU_OK ...` on rank 0 when everything matched; any mismatch raises (non-zero exit).
"""
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def transport_parity(rank, world, dev, transport):
    import torch
    import torch.distributed as dist
    from swarmdb_b200._native import Shard
    from swarmdb_b200.sharded import PeerExchange, ShardExchange, TorchCudaBackend, shard_map_numbered

    A, F, S, L, STEPS = 1 << 18, 64, 16384, 256, 5
    G = A // F
    perm = np.random.default_rng(2).permutation(A).astype(np.uint32)
    smap = shard_map_numbered("agent_", 7, A, world)
    shard = Shard(max_agents=A, ring_slots=256, arena_bytes=1 << 31, max_payload_bytes=L, max_groups=G,
                  member_pool_entries=A + 1024, max_batch_sends=S, max_batch_payload=S * L,
                  max_recv_records=S * F * 2 + (1 << 16), max_recv_payload=(S * F * 2 + (1 << 16)) * L,
                  device=dev.index, shard_id=rank, num_shards=world)
    shard.set_agent_shards(smap)
    for g in range(G):
        shard.create_group(g, perm[g * F:(g + 1) * F])
    if transport in ("peer", "prefetch"):
        ex = PeerExchange(shard, rank, world, S, S * L, dev)
    else:
        ex = ShardExchange(shard, rank, world, S, S * L, TorchCudaBackend(dev, shard))

    alnum = np.frombuffer(b"ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789", np.uint8)

    def batch(step, r):                                       # every rank can regenerate every rank's batch
        rng = np.random.default_rng(1000 + 17 * step + r)
        n = S if (step, r) != (3, world - 1) else 0           # one rank exports an EMPTY batch in one step
        grp = rng.integers(0, G, n).astype(np.uint32)
        snd = rng.integers(0, A, n).astype(np.uint32)         # sometimes a member: skip-sender across shards
        prio = rng.integers(0, 4, n).astype(np.uint8)
        typ = rng.integers(0, 7, n).astype(np.uint8)
        lens = rng.integers(0, L + 1, n).astype(np.uint16)
        off = np.arange(n, dtype=np.uint64) * L
        pay = alnum[rng.integers(0, 62, n * L + 32)]
        return snd, grp, prio, typ, lens, off, pay

    shard.digest_reset()
    delivered = 0
    if transport == "prefetch":
        ex.export(*batch(0, rank))
    for step in range(STEPS):
        if transport == "prefetch":
            # pipelined order: step's export is out already; import it, export the NEXT step and start its flag wait +
            # descriptor pull on the prefetch stream, then drain - the streams must equal the unpipelined sequence
            ex.import_all()
            if step + 1 < STEPS:
                ex.export(*batch(step + 1, rank))
                if step != 1:                                 # one step without a prefetch: both paths alternate
                    ex.prefetch()
        else:
            ex.step(*batch(step, rank))
        _, total, _ = shard.receive_batch(None, 7 if step % 2 == 0 else 1000, 0, copy_out=False)   # partial drains too
        shard.digest_fold()
        delivered += total
    while True:
        _, total, _ = shard.receive_batch(None, 1000, 0, copy_out=False)
        shard.digest_fold()
        if total == 0:
            break
        delivered += total
    st = shard.stats()
    assert st["ring_overflow"] == 0, st
    dg = shard.digest_read()
    foreign = dg[smap != rank]
    assert not foreign.any(), f"rank {rank}: {int((foreign != 0).sum())} agents of other shards received records here"
    t = torch.from_numpy(dg.view(np.int64).copy()).to(dev)
    dist.all_reduce(t)                                        # owners are disjoint: the sum is the union (mod 2^64)
    cnt = torch.tensor([delivered], dtype=torch.int64, device=dev)
    dist.all_reduce(cnt)
    merged = t.cpu().numpy().view(np.uint64)
    if hasattr(ex, "close"):
        ex.close()
    shard.close()
    if rank == 0:
        from oracle.cpu_ref import CpuOracle
        cpu = CpuOracle(A, G)
        for g in range(G):
            cpu.create_group(g, perm[g * F:(g + 1) * F])
        cpu.digest_enable()
        routed = 0
        for step in range(STEPS):
            for r in range(world):
                b = batch(step, r)
                if len(b[0]):
                    routed += cpu.send_group_batch(*b)[1]
            cpu.receive_counts(None, 7 if step % 2 == 0 else 1000, 0)
        while cpu.receive_counts(None, 1000, 0)[1]:
            pass
        want = cpu.digest_read()
        bad = np.nonzero(merged != want)[0]
        assert len(bad) == 0, f"{transport}: {len(bad)} agents differ from the single-queue oracle, first {bad[:8].tolist()}"
        assert int(cnt.item()) == routed, (int(cnt.item()), routed)
        cpu.close()
        return routed
    return 0


def frontend_parity(rank, world, dev, tmp):
    import torch.distributed as dist
    import swarmdb_b200 as sdb
    from swarmdb_b200 import sharded
    from tests.test_sharded_frontend_cpu import AGENTS, _play, _script, _setup, _view

    cfg = sdb.GpuConfig(device=dev.index, max_agents=64, max_groups=8, flush_threshold=64, max_payload_bytes=1024,
                        ring_slots=1024, arena_bytes=1 << 24, deterministic_ids=True)
    db = sharded.ShardedSwarmsDB(rank, world, save_dir=f"{tmp}/r{rank}", auto_save=False, gpu_config=cfg)   # default exchange
    assert isinstance(db.exchange, sharded.PeerExchange)
    _setup(db)
    script = _script(np.random.default_rng(31), world, 7)
    mine = [a for a in AGENTS if db.owner(a) == rank]
    got, sent = {}, []
    for rnd, per_rank in enumerate(script):
        sent.append(_play(db, per_rank[rank]))
        db.flush()
        for a in mine:
            got.setdefault(a, []).extend((m.id, _view(m)) for m in db.receive_messages(a, 3 if rnd % 2 == 0 else 1000))
    for a in mine:
        got.setdefault(a, []).extend((m.id, _view(m)) for m in db.receive_messages(a, 100000))
    allgot, allsent = [None] * world, [None] * world
    dist.all_gather_object(allgot, got)
    dist.all_gather_object(allsent, sent)
    db.close()                                                # collective; must not raise (ADVICE sharded.py:540)
    n = 0
    if rank == 0:
        single = sdb.SwarmsDB(save_dir=f"{tmp}/single", auto_save=False, gpu_config=cfg)
        _setup(single)
        for per_rank in script:
            for calls in per_rank:
                _play(single, calls)
        merged = {}
        for d in allgot:
            merged.update(d)
        for a in AGENTS:
            want = [_view(m) for m in single.receive_messages(a, 100000)]
            have = [v for _, v in merged.get(a, [])]
            assert want == have, (a, len(want), len(have))
            n += len(want)
        ids_sent = [i for per in allsent for rnd in per for i in rnd]
        ids_seen = {i for d in allgot for lst in d.values() for i, _ in lst}
        assert len(ids_sent) == len(set(ids_sent)) and ids_seen <= set(ids_sent) and n > 100
        single.close()
    return n


def main():
    import tempfile

    import torch
    import torch.distributed as dist
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    tmp = tempfile.mkdtemp(prefix="sdb_mp_") if rank == 0 else None
    box = [tmp]
    dist.broadcast_object_list(box, src=0)
    res = {}
    for transport in os.environ.get("SDB_MP_TRANSPORTS", "peer,prefetch,nccl").split(","):
        res[transport] = transport_parity(rank, world, dev, transport)
        dist.barrier()
    res["frontend_msgs"] = frontend_parity(rank, world, dev, box[0])
    dist.barrier()
    if rank == 0:
        print(f"MP_GPU_OK world={world} {res}", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
