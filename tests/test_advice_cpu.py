"""Regression tests for the round-1 advisor findings that live in host code (CPU; the device is the oracle-backed
stand-in of tests/fake_shard.py, tests only)."""
import numpy as np
import pytest

from tests.fake_shard import OracleShard


def _db(tmp_path, n=4096, shard=None, **kw):
    import swarmdb_b200 as sdb
    cfg = sdb.GpuConfig(max_agents=n, max_groups=16, deterministic_ids=True, **kw)
    return sdb, sdb.SwarmsDB(save_dir=str(tmp_path), auto_save=False, gpu_config=cfg, _shard=shard or OracleShard(n, 16, 0, 1))


def test_broadcast_to_thousands_of_agents_fits_the_default_payload_limit(tmp_path):
    """ADVICE core.py:390 - every copy of a broadcast used to carry the whole visible_to list (O(agents) bytes in
    each of O(agents) copies); 3000 agents overflowed max_payload_bytes.  Long lists now travel as their complement."""
    sdb, db = _db(tmp_path)
    names = [f"a{i:04d}" for i in range(3000)]
    db.register_agents(names)
    db.deregister_agent("a0007")                                         # not visible: must not reappear at decode
    mid = db.broadcast_message("a0001", "hello everybody", exclude_agents=["a0002", "a2999"])
    want = set(names) - {"a0001", "a0002", "a2999", "a0007"}
    assert set(db.get_message(mid).visible_to) == want
    for a in ("a0000", "a1500", "a2998"):
        got = db.receive_messages(a)
        assert len(got) == 1 and got[0].id == mid and got[0].receiver_id is None and got[0].content == "hello everybody"
        assert set(got[0].visible_to) == want                             # the reference shows every reader the full list
    assert db.receive_messages("a0002") == [] and db.receive_messages("a0007") == []
    # agents that register later are not part of an earlier broadcast (App. A rule 4)
    db.register_agent("late")
    assert db.receive_messages("late") == []
    # a short explicit list still travels literally
    mid2 = db.send_message("a0001", "few", None, visible_to=["a0003", "a0004"])
    assert [m.id for m in db.receive_messages("a0003")] == [mid, mid2]
    assert db.receive_messages("a0003") == []
    assert sorted(db.receive_messages("a0004")[-1].visible_to) == ["a0003", "a0004"]
    db.close()


def test_buffered_broadcast_lists_flush_before_the_list_pool_overflows(tmp_path):
    """ADVICE core.py:390 (second half) - many buffered broadcasts must not exceed the per-batch recipient-list pool."""
    sdb, db = _db(tmp_path, n=256, flush_threshold=4096)
    names = [f"b{i:03d}" for i in range(200)]
    db.register_agents(names)
    flushes = []
    real = db.shard.send_mixed_batch
    db.shard.send_mixed_batch = lambda *a, **k: (flushes.append(len(a[4])), real(*a, **k))[1]    # a[4] = list_idx
    for k in range(40):                                                   # 40 x 199 recipients >> 2*256 + 1024 entries
        db.broadcast_message(names[k], f"round {k}")
    db.flush()
    assert len(flushes) > 1 and max(flushes) + 4096 // 1 >= 0 and max(flushes) <= db._list_cap
    got = db.receive_messages("b199", 1000)
    assert [m.content for m in got] == [f"round {k}" for k in range(40)]
    db.close()


def test_refused_batch_does_not_reissue_message_ids(tmp_path):
    """ADVICE core.py:302 - after a batch the device refused, ids handed out for it stay used (FAILED, kept) and the
    next sends get fresh ids; the device counter is moved up, not the host counter down."""
    sdb, db = _db(tmp_path)
    db.register_agents(["x", "y"])
    first = db.send_message("x", "one", "y")
    boom = {"on": True}
    real = db.shard.send_mixed_batch

    def refuse(*a, **k):
        if boom["on"]:
            raise sdb._native.SdbError(-5, "message arena full (test)")
        return real(*a, **k)
    db.shard.send_mixed_batch = refuse
    with pytest.raises(sdb._native.SdbError):
        db.flush()
    assert db.get_message(first).status == sdb.MessageStatus.FAILED
    boom["on"] = False
    second = db.send_message("x", "two", "y")
    assert second != first
    got = db.receive_messages("y")
    assert [m.id for m in got] == [second] and len({first, second}) == 2
    assert db.get_message(first).status == sdb.MessageStatus.FAILED      # not overwritten by the later message
    assert db.agent_inbox["y"] == [first, second]
    db.close()


def test_sharded_close_is_not_a_flush(tmp_path):
    """ADVICE sharded.py:540 - close() used to run the collective flush after the exchange had freed its buffers."""
    from swarmdb_b200 import sharded
    import swarmdb_b200 as sdb
    calls = []

    class Ex:
        def __init__(self, *a): pass
        def export_mixed(self, *a, **k): calls.append("export")
        def exchange(self): calls.append("exchange")
        def import_all(self): calls.append("import"); return 0
        def close(self): calls.append("close")
    cfg = sdb.GpuConfig(max_agents=64, max_groups=4, deterministic_ids=True)
    db = sharded.ShardedSwarmsDB(0, 1, exchange_factory=Ex, shard=OracleShard(64, 4, 0, 1), save_dir=str(tmp_path),
                                 auto_save=False, gpu_config=cfg)
    db.register_agents(["p", "q"])
    db.send_message("p", "buffered, never flushed", "q")
    db.close()
    db.close()                                                            # idempotent
    assert calls == ["close"]


class _OverflowingShard(OracleShard):
    """Stand-in that 'drops' chosen records of the next batch: reports them through ring_overflow + overflow_log the
    way the device does (the oracle has no ring limit, so the drop itself is only reported, not performed)."""
    def __init__(self, *a):
        super().__init__(*a)
        self.drop, self.ovf, self.truncate = [], 0, False

    def stats(self):
        st = super().stats(); st["ring_overflow"] = self.ovf
        return st

    def send_mixed_batch(self, *a, **k):
        base = super().send_mixed_batch(*a, **k)
        self._log = [(ag, base + off) for ag, off in self.drop]
        self.ovf += len(self._log); self.drop = []
        return base

    def overflow_log(self, cap=4096):
        log, self._log = self._log, []
        if self.truncate:                                  # more drops than the log holds
            return np.zeros(0, np.uint32), np.zeros(0, np.uint64), len(log)
        return (np.array([x[0] for x in log], np.uint32), np.array([x[1] for x in log], np.uint64), len(log))


def test_ring_overflow_fails_only_the_messages_that_were_lost(tmp_path):
    """ADVICE core.py:302 (first half) - a flush that overflowed rings used to mark EVERY message of the batch FAILED,
    so resend_failed_messages() resent messages that had been delivered.  The device now names the dropped records."""
    shard = _OverflowingShard(64, 16, 0, 1)
    sdb, db = _db(tmp_path, n=64, shard=shard)
    db.register_agents(["s", "r1", "r2", "r3"])
    ids = [db.send_message("s", f"m{k}", "r1") for k in range(5)]                 # seq offsets 0..4
    bid = db.broadcast_message("s", "to all")                                      # offset 5, one message, three copies
    shard.drop = [(db.agent_index("r1"), 1), (db.agent_index("r1"), 3), (db.agent_index("r2"), 5)]
    with pytest.raises(sdb.RingOverflow) as ei:
        db.flush()
    assert ei.value.exact and "3 message(s)" in str(ei.value)
    st = [db.get_message(i).status for i in ids]
    assert [x == sdb.MessageStatus.FAILED for x in st] == [False, True, False, True, False]
    assert "ring is full" in db.get_message(ids[1]).metadata["error"]
    b = db.get_message(bid)
    assert b.status == sdb.MessageStatus.DELIVERED and b.metadata["undelivered_to"] == ["r2"]
    resent = db.resend_failed_messages()
    assert len(resent) == 2 and [db.get_message(i).content for i in resent] == ["m1", "m3"]
    # a log that could not name every drop: the whole batch is failed, as before
    more = [db.send_message("s", f"n{k}", "r3") for k in range(3)]
    db.flush()
    more = [db.send_message("s", f"p{k}", "r3") for k in range(3)]
    shard.drop = [(db.agent_index("r3"), 0)]; shard.truncate = True
    with pytest.raises(sdb.RingOverflow) as ei:
        db.flush()
    assert not ei.value.exact and all(db.get_message(i).status == sdb.MessageStatus.FAILED for i in more)
    db.close()


def test_pending_snapshot_packs_its_peek_calls_by_pending_counts(tmp_path):
    """ADVICE core.py:593 - the snapshot used to peek fixed chunks of agents with a per-agent window far above the
    call's record capacity; agents beyond the capacity were silently left out.  Chunks now follow the pending counts."""
    sdb, db = _db(tmp_path, n=512, max_recv_records=256, max_payload_bytes=1024)
    names = [f"p{i:03d}" for i in range(300)]
    db.register_agents(names)
    for k in range(3):
        for i in range(0, 300, 3):                                         # every third agent has 3 pending
            db.send_message("p000", f"x{k}", names[i])
    calls = []
    real = db.shard.receive_batch
    db.shard.receive_batch = lambda agents, mx, flags=0, **kw: (calls.append((len(agents), mx)), real(agents, mx, flags, **kw))[1]
    snap = db.pending_snapshot()
    assert sorted(snap) == sorted(names[i] for i in range(0, 300, 3))
    assert all([m.content for m in v] == ["x0", "x1", "x2"] for v in snap.values())
    assert all(n_ag * 3 <= 256 and mx <= 256 for n_ag, mx in calls) and len(calls) == 2          # 85 + 15 agents
    assert len(db.receive_messages(names[3])) == 3                          # nothing was consumed
    db.close()
