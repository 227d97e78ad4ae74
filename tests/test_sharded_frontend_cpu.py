"""world_size-2 gloo test of the multi-process drop-in front-end (sharded.make_sharded_swarmsdb):
replicated registry calls, rank-local sends, collective flush with composite sequence bases,
owner-local receive.  The device is the oracle-backed stand-in of tests/fake_shard.py (tests only);
the expected streams come from a single-process SwarmsDB fed the same sends in (round, rank, call) order."""
import os
import socket

import numpy as np

from swarmdb_b200 import sharded
from tests.fake_shard import OracleShard
from tests.test_sharded_cpu import CpuBackend

AGENTS = [f"agent_{i:03d}" for i in range(40)]
GROUPS = {"g0": AGENTS[0:12], "g1": AGENTS[8:30], "g2": AGENTS[25:40]}


def _script(rng, world, rounds):
    """The same pseudo-random send script on every rank: script[round][rank] = list of calls."""
    out = []
    for _ in range(rounds):
        per_rank = []
        for _r in range(world):
            calls = []
            for _c in range(int(rng.integers(0, 9))):
                k = int(rng.integers(0, 4))
                s = AGENTS[int(rng.integers(0, len(AGENTS)))]
                body = "".join(chr(int(c)) for c in rng.integers(97, 123, int(rng.integers(0, 90))))
                prio = int(rng.integers(0, 4))
                if k == 0:
                    calls.append(("p2p", s, AGENTS[int(rng.integers(0, len(AGENTS)))], body, prio))
                elif k == 1:
                    calls.append(("group", s, ["g0", "g1", "g2"][int(rng.integers(0, 3))], {"text": body, "n": prio}, prio))
                elif k == 2:
                    calls.append(("bcast", s, None, body, prio))
                else:
                    calls.append(("meta", s, AGENTS[int(rng.integers(0, len(AGENTS)))], [body, prio], prio))
            per_rank.append(calls)
        out.append(per_rank)
    return out


def _play(db, calls):
    from swarmdb_b200.core import MessagePriority, MessageType
    ids = []
    for kind, s, t, body, prio in calls:
        if kind == "p2p":
            ids.append(db.send_message(s, body, t, MessageType.COMMAND, MessagePriority(prio)))
        elif kind == "group":
            ids.extend(db.send_to_group(s, t, body, MessageType.SYSTEM, MessagePriority(prio)))
        elif kind == "bcast":
            ids.append(db.broadcast_message(s, body, MessageType.STATUS, MessagePriority(prio), exclude_agents=[AGENTS[3]]))
        else:
            ids.append(db.send_message(s, body, t, MessageType.FUNCTION_RESULT, MessagePriority(prio), metadata={"k": prio, "who": s}))
    return ids


def _view(m):
    return (m.sender_id, m.receiver_id, m.content if isinstance(m.content, str) else repr(m.content), m.type.value,
            int(m.priority), tuple(sorted((k, str(v)) for k, v in m.metadata.items())), tuple(sorted(m.visible_to)))


def _setup(db):
    db.register_agents(AGENTS)
    for g, members in GROUPS.items():
        db.add_agent_group(g, list(members))


def _worker(rank, world, port, q, tmp):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from swarmdb_b200.core import GpuConfig
        cfg = GpuConfig(max_agents=64, max_groups=8, flush_threshold=64, max_payload_bytes=1024, deterministic_ids=True)
        db = sharded.make_sharded_swarmsdb(
            rank, world, shard=OracleShard(64, 8, rank, world),
            exchange_factory=lambda sh, r, w, ms, mp: sharded.ShardExchange(sh, r, w, ms, mp, CpuBackend()),
            save_dir=f"{tmp}/r{rank}", auto_save=False, gpu_config=cfg)
        _setup(db)
        script = _script(np.random.default_rng(11), world, 5)
        sent, got = [], {}
        mine = [a for a in AGENTS if db.owner(a) == rank]
        for rnd, per_rank in enumerate(script):
            sent.append(_play(db, per_rank[rank]))
            db.flush()
            for a in mine:
                lim = 3 if rnd % 2 == 0 else 1000                  # partial drains in between
                got.setdefault(a, []).extend((m.id, _view(m)) for m in db.receive_messages(a, lim))
        # peek and the device snapshot are rank-local reads: they show what the next receive returns and consume nothing
        snap = db.pending_snapshot()
        peeks_ok = True
        for a in mine:
            peek = [(m.id, _view(m)) for m in db.peek_messages(a, 100000)]
            rest = [(m.id, _view(m)) for m in db.receive_messages(a, 100000)]
            peeks_ok &= [i for i, _ in peek] == [i for i, _ in rest] == [m.id for m in snap.get(a, [])]
            got[a].extend(rest)
        # strictness: unknown agents cannot be introduced by a rank-local call
        try:
            db.send_message(AGENTS[0], "x", "nobody")
            strict = False
        except KeyError:
            strict = True
        try:
            other = next(a for a in AGENTS if db.owner(a) != rank)
            db.receive_messages(other)
            strict = False
        except ValueError:
            pass
        allgot, allsent, allpeeks = [None] * world, [None] * world, [None] * world
        dist.all_gather_object(allpeeks, peeks_ok)
        dist.all_gather_object(allgot, got)
        dist.all_gather_object(allsent, sent)
        db.close()
        if rank == 0:
            single = __import__("swarmdb_b200").SwarmsDB(save_dir=f"{tmp}/single", auto_save=False, gpu_config=cfg,
                                                         _shard=OracleShard(64, 8, 0, 1))
            _setup(single)
            for per_rank in script:
                for calls in per_rank:
                    _play(single, calls)
            merged = {}
            for d in allgot:
                merged.update(d)
            bad = []
            for a in AGENTS:
                want = [_view(m) for m in single.receive_messages(a, 100000)]
                have = [v for _, v in merged.get(a, [])]
                if want != have:
                    bad.append((a, len(want), len(have)))
            # every id a receiver saw was returned by exactly one send call on some rank
            ids_sent = [i for per in allsent for rnd in per for i in rnd]
            ids_seen = {i for d in allgot for lst in d.values() for i, _ in lst}
            q.put({"bad": bad, "strict": strict, "peeks": all(allpeeks), "unique": len(ids_sent) == len(set(ids_sent)),
                   "seen_subset": ids_seen <= set(ids_sent), "n_seen": len(ids_seen),
                   "n_msgs": sum(len(v) for v in merged.values())})
            single.close()
    finally:
        dist.destroy_process_group()


def test_sharded_frontend_matches_single_process(tmp_path):
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q, str(tmp_path))) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
        assert p.exitcode == 0
    res = q.get(timeout=5)
    assert res["bad"] == [] and res["strict"] and res["peeks"] and res["unique"] and res["seen_subset"], res
    assert res["n_msgs"] > 100 and res["n_seen"] > 50, res


def test_composite_seq_layout():
    assert sharded.composite_seq(1, 0) == 1 << 40
    assert sharded.composite_seq(3, 5, 7) == (3 << 40) | (5 << 32) | 7
    assert sharded.composite_seq(2, 0) > sharded.composite_seq(1, 255, (1 << 32) - 1)


def _balancer_worker(rank, world, port, q, tmp):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from swarmdb_b200.core import GpuConfig
        cfg = GpuConfig(max_agents=64, max_groups=8, flush_threshold=64, max_payload_bytes=1024, deterministic_ids=True)
        db = sharded.make_sharded_swarmsdb(
            rank, world, shard=OracleShard(64, 8, rank, world),
            exchange_factory=lambda sh, r, w, ms, mp: sharded.ShardExchange(sh, r, w, ms, mp, CpuBackend()),
            save_dir=f"{tmp}/r{rank}", auto_save=False, gpu_config=cfg)
        _setup(db)
        db.register_llm_backends(["b0", "b1", "b2", "b3"], [1, 2, 3, 1])
        picks = db.select_llm_backends(10 + 5 * rank)                   # rank-local picks change the rank's own table
        local_after_picks = db.llm_backend_loads()
        db.flush()                                                       # collective: the 4 load deltas are summed across ranks
        synced = db.llm_backend_loads()
        # queue-fed loads: agents are assigned identically everywhere, only their owner knows their backlog
        for i, a in enumerate(AGENTS[:16]):
            db.assign_llm_backend(a, f"b{i % 4}")
        for i in range(12):
            db.send_message(AGENTS[20 + rank], f"x{i}", AGENTS[i % 16])
        db.flush()
        from_queues = db.refresh_llm_backend_loads()
        allv = [None] * world
        dist.all_gather_object(allv, (len(picks), local_after_picks, synced, from_queues))
        db.close()
        if rank == 0:
            q.put(allv)
    finally:
        dist.destroy_process_group()


def test_one_balancer_for_all_shards(tmp_path):
    """backend loads are global - after a collective flush every rank holds base + the sum of all ranks'
    changes; queue-fed loads are summed over the owners of the assigned agents."""
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_balancer_worker, args=(r, 2, port, q, str(tmp_path))) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
        assert p.exitcode == 0
    res = q.get(timeout=5)
    (n0, l0, s0, f0), (n1, l1, s1, f1) = res
    assert sum(l0.values()) == n0 and sum(l1.values()) == n1 and n0 != n1          # ranks diverged before the flush ...
    assert s0 == s1 and sum(s0.values()) == n0 + n1                                # ... and agree afterwards
    assert {b: l0[b] + l1[b] for b in l0} == s0
    assert f0 == f1 and sum(f0.values()) == 24                                     # 12 messages per rank, all to assigned agents
