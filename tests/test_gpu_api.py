"""GPU parity at the Python surface: `swarmdb_b200.SwarmsDB` replays the golden scenarios that
the UNMODIFIED reference class produced (tests/golden/*.json) and must return identical
messages - ids by rank, sender, receiver, content, type, priority, status, metadata,
token_count, visible_to - for every receive call, plus the same inbox side record."""
import json

import pytest

from oracle import scenarios

pytestmark = pytest.mark.gpu


def _db(tmp_path, **kw):
    import swarmdb_b200 as sdb
    cfg = sdb.GpuConfig(max_agents=4096, ring_slots=2048, arena_bytes=1 << 26, deterministic_ids=True, **kw)
    return sdb, sdb.SwarmsDB(save_dir=str(tmp_path), auto_save=False, gpu_config=cfg)


@pytest.mark.parametrize("name", sorted(scenarios.SCENARIOS))
def test_surface_matches_reference_golden(golden_dir, tmp_path, name):
    doc = json.loads((golden_dir / f"{name}.json").read_text())
    ops = scenarios.SCENARIOS[name]()
    sdb, db = _db(tmp_path)
    try:
        id_rank = {}
        got = scenarios.run_ops(db, ops, sdb, id_rank=id_rank)
        final = scenarios.final_state(db)
        history = json.loads(json.dumps(scenarios.history_state(db, id_rank)))      # N2: the reference's history file
    finally:
        db.close()
    if name in scenarios.HASHED:
        assert scenarios.digest(got) == doc["expected_digest"]
        assert scenarios.digest(history) == doc["history_digest"]
    else:
        got = json.loads(json.dumps(got))
        for i, (g, e) in enumerate(zip(got, doc["expected"])):
            assert g == e, (i, ops[i])
        assert history == doc["history"]
    assert json.loads(json.dumps(final)) == doc["final"]


def test_import_path_of_the_reference_rest_layer():
    from swarmdb import KafkaConfig, Message, MessagePriority, MessageStatus, MessageType, SwarmsDB  # noqa: F401


def test_flush_threshold_and_group_mutation(tmp_path):
    sdb, db = _db(tmp_path, flush_threshold=8)
    try:
        members = ["a", "b"]
        db.add_agent_group("g", members)
        db.send_to_group("s", "g", "one")
        members.append("c")                       # the reference stores the caller's list (M:1223)
        db.send_to_group("s", "g", "two")
        for i in range(40):
            db.send_message("s", f"m{i}", "a")
        assert [m.content for m in db.receive_messages("c")] == ["two"]
        a = [m.content for m in db.receive_messages("a", 1000)]
        assert a == ["one", "two"] + [f"m{i}" for i in range(40)]
        assert db.receive_messages("a") == []
    finally:
        db.close()


def test_priority_dequeue_extension(tmp_path):
    sdb, db = _db(tmp_path, priority_dequeue=True)
    try:
        for i, p in enumerate([0, 3, 1, 3, 2, 0]):
            db.send_message("s", f"m{i}", "r", priority=sdb.MessagePriority(p))
        assert [m.content for m in db.receive_messages("r", 4)] == ["m1", "m3", "m4", "m2"]
        assert [m.content for m in db.receive_messages("r", 4)] == ["m0", "m5"]
    finally:
        db.close()


def test_balancer_surface(tmp_path):
    sdb, db = _db(tmp_path)
    try:
        db.register_llm_backends(["gpt", "claude", "llama"], [1, 2, 1])
        assert db.select_llm_backend("x") is None           # balancing off, nothing assigned (M:1323-1325)
        db.assign_llm_backend("x", "gpt")
        assert db.get_llm_backend("x") == "gpt" and db.select_llm_backend("x") == "gpt"
        db.set_llm_load_balancing(True)
        picks = [db.select_llm_backend(f"agent{i}") for i in range(8)]
        assert picks.count("claude") == 4 and db.get_llm_backend("agent0") == picks[0]
        assert sum(db.llm_backend_loads().values()) == 8
        db.release_llm_backend("claude", 2)
        assert db.llm_backend_loads()["claude"] == 2
    finally:
        db.close()


def test_random_interleavings_match_the_pinned_oracle(tmp_path):
    """Property test on the GPU surface: seeded random op sequences vs the Python oracle (itself pinned
    to the reference's goldens).  Seeds instead of hypothesis shrinking: device construction is the cost."""
    from oracle import pyref
    for seed in range(20, 28):
        ops = scenarios.scenario_random(seed, n_agents=10 + seed % 7, n_ops=120)
        want = json.loads(json.dumps(scenarios.run_ops(pyref.OracleSwarmsDB(id_factory=pyref.counter_ids()), ops, pyref)))
        sdb, db = _db(tmp_path)
        try:
            got = json.loads(json.dumps(scenarios.run_ops(db, ops, sdb)))
        finally:
            db.close()
        for i, (g, w) in enumerate(zip(got, want)):
            assert g == w, (seed, i, ops[i])


def test_peek_and_device_history_snapshot(tmp_path):
    """N2: non-destructive device dump -> history file in the reference's schema (M:878-884)."""
    sdb, db = _db(tmp_path)
    try:
        db.add_agent_group("g", ["a", "b", "c"])
        db.send_to_group("s", "g", "hello", metadata={"k": 1})
        db.send_message("s", {"x": [1, 2]}, "a", priority=sdb.MessagePriority.HIGH)
        peek = db.peek_messages("a")
        assert [m.content for m in peek] == ["hello", {"x": [1, 2]}] and peek[0].status == sdb.MessageStatus.DELIVERED
        assert [m.content for m in db.peek_messages("a", 1)] == ["hello"]          # nothing was consumed
        # bulk, index-level sends leave no host-side Message objects: only the device knows them
        import numpy as np
        ia, ib = db.agent_index("a"), db.agent_index("b")
        buf = np.frombuffer(b"bulk-one" + bytes(24) + b"bulk-two" + bytes(24), np.uint8)
        db.send_batch([ib, ib], [ia, ia], [1, 1], [0, 0], [8, 8], [0, 32], buf)
        snap = db.pending_snapshot()
        assert [m.content for m in snap["a"]] == ["hello", {"x": [1, 2]}, "bulk-one", "bulk-two"]
        assert [m.content for m in snap["b"]] == ["hello"] and "s" not in snap
        db.save_message_history("h.json", include_device=True)
        doc = json.loads((tmp_path / "h.json").read_text())
        assert set(doc) == {"messages", "agent_inbox", "registered_agents", "timestamp", "message_count"}
        assert len(doc["agent_inbox"]["a"]) == 4 and len(doc["messages"]) == 6
        one = next(iter(doc["messages"].values()))
        assert list(one) == ["id", "sender_id", "receiver_id", "content", "type", "priority", "timestamp", "status",
                             "metadata", "token_count", "visible_to"]          # field order of M:54-82
        # the snapshot consumed nothing
        assert [m.content for m in db.receive_messages("a")] == ["hello", {"x": [1, 2]}, "bulk-one", "bulk-two"]
        # and a fresh instance can load the file (M:894-934)
        sdb2, db2 = _db(tmp_path / "second")
        try:
            db2.load_message_history(tmp_path / "h.json")
            assert db2.message_count == doc["message_count"] and "a" in db2.registered_agents
            assert len(db2.messages) == 6
        finally:
            db2.close()
    finally:
        db.close()


def test_plain_c_caller(tmp_path):
    """examples/c_abi_demo.c (strict C99) creates a shard, fans one message out to a group and drains a member."""
    import subprocess
    from tests.test_abi_cpu import _build_c_demo
    p = subprocess.run([str(_build_c_demo(tmp_path))], capture_output=True, text=True)
    assert p.returncode == 0 and 'agent 2 got 1 message(s)' in p.stdout and '"hello"' in p.stdout, p.stdout + p.stderr


@pytest.mark.gpu
def test_python_surface_fails_only_the_messages_a_full_ring_dropped(tmp_path):
    """core.flush + sdb_overflow_log on the real device: 10 messages into a 4-slot ring - four are delivered and stay
    DELIVERED, the other six are FAILED and are the only ones resend_failed_messages() sends again."""
    import swarmdb_b200 as sdb
    cfg = sdb.GpuConfig(max_agents=64, max_groups=4, ring_slots=4, deterministic_ids=True, flush_threshold=1000)
    db = sdb.SwarmsDB(save_dir=str(tmp_path), auto_save=False, gpu_config=cfg)
    ids = [db.send_message("s", f"m{k}", "r") for k in range(10)]
    with pytest.raises(sdb.RingOverflow) as ei:
        db.flush()
    assert ei.value.exact
    failed = [db.get_message(i).status == sdb.MessageStatus.FAILED for i in ids]
    assert sum(failed) == 6                               # (which four win the slots is decided by the device: buffered
    kept = [f"m{k}" for k in range(10) if not failed[k]]  #  sends travel as a mixed batch, whose slots are claimed atomically)
    assert [m.content for m in db.receive_messages("r")] == kept
    resent = db.resend_failed_messages()
    assert len(resent) == 6 and sorted(db.get_message(i).content for i in resent) == sorted(f"m{k}" for k in range(10) if failed[k])
    with pytest.raises(sdb.RingOverflow) as ei:
        db.flush()                                         # six into four slots again: two more are lost, by name
    assert ei.value.exact
    again = [db.get_message(i) for i in resent]
    assert sum(m.status == sdb.MessageStatus.FAILED for m in again) == 2
    assert [m.content for m in db.receive_messages("r")] == [m.content for m in again if m.status != sdb.MessageStatus.FAILED]
    db.close()
