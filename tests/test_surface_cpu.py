"""Host logic of the drop-in surface (swarmdb_b200/core.py) against the goldens the UNMODIFIED reference
produced - on CPU.  The device is replaced by the oracle-backed stand-in of tests/fake_shard.py (tests only:
the product itself has no CPU path), so everything core.py does - id assignment, buffering, wire encoding of
content / metadata / visible_to, group sync, inbox side record, decode, status transitions, and the history
file (schema M:878-884) - is checked without a GPU.  tests/test_gpu_api.py repeats it on the device."""
import json

import pytest

from oracle import scenarios
from tests.fake_shard import OracleShard


def _db(tmp_path, **kw):
    import swarmdb_b200 as sdb
    cfg = sdb.GpuConfig(max_agents=4096, max_groups=256, deterministic_ids=True, **kw)
    return sdb, sdb.SwarmsDB(save_dir=str(tmp_path), auto_save=False, gpu_config=cfg, _shard=OracleShard(4096, 256, 0, 1))


@pytest.mark.parametrize("name", sorted(scenarios.SCENARIOS))
def test_surface_host_logic_matches_reference_golden(golden_dir, tmp_path, name):
    doc = json.loads((golden_dir / f"{name}.json").read_text())
    ops = scenarios.SCENARIOS[name]()
    sdb, db = _db(tmp_path)
    try:
        id_rank = {}
        got = scenarios.run_ops(db, ops, sdb, id_rank=id_rank)
        final = scenarios.final_state(db)
        history = json.loads(json.dumps(scenarios.history_state(db, id_rank)))
    finally:
        db.close()
    if name in scenarios.HASHED:
        assert scenarios.digest(got) == doc["expected_digest"]
        assert scenarios.digest(history) == doc["history_digest"]
    else:
        got = json.loads(json.dumps(got))
        for i, (g, e) in enumerate(zip(got, doc["expected"])):
            assert g == e, (i, ops[i])
        for k in doc["history"]:
            assert history[k] == doc["history"][k], k
    assert json.loads(json.dumps(final)) == doc["final"]


def test_history_file_round_trip(tmp_path):
    """save -> load into a fresh instance restores messages, inboxes, registry and the count (M:895-934)."""
    sdb, db = _db(tmp_path / "a")
    db.add_agent_group("g", ["x", "y", "z"])
    db.send_to_group("boss", "g", {"task": 1}, priority=sdb.MessagePriority.HIGH, metadata={"k": "v"})
    db.send_message("x", "ping", "y")
    db.broadcast_message("boss", ["all", 1])
    assert [m.content for m in db.receive_messages("y")] == [{"task": 1}, "ping", ["all", 1]]
    db.save_message_history("h.json")
    before = {k: v.to_dict() for k, v in db.messages.items()}
    inbox = {a: list(v) for a, v in db.agent_inbox.items()}
    sdb2, db2 = _db(tmp_path / "b")
    db2.load_message_history(tmp_path / "a" / "h.json")
    assert {k: v.to_dict() for k, v in db2.messages.items()} == before
    assert db2.agent_inbox == inbox and db2.registered_agents == db.registered_agents
    assert db2.message_count == db.message_count
    assert db2.get_message(next(iter(before))).status in (sdb.MessageStatus.READ, sdb.MessageStatus.DELIVERED)
    db.close(); db2.close()


@pytest.mark.parametrize("seed", range(40, 52))
def test_random_interleavings_match_the_pinned_oracle_on_cpu(tmp_path, seed):
    """Seeded random op mixes (register / p2p / visible_to / groups / broadcasts / partial drains) through the surface
    versus oracle/pyref.py, which is itself pinned to the reference's goldens."""
    from oracle import pyref
    ops = scenarios.scenario_random(seed, n_agents=8 + seed % 9, n_ops=150)
    want = json.loads(json.dumps(scenarios.run_ops(pyref.OracleSwarmsDB(id_factory=pyref.counter_ids()), ops, pyref)))
    sdb, db = _db(tmp_path)
    try:
        got = json.loads(json.dumps(scenarios.run_ops(db, ops, sdb)))
        final = scenarios.final_state(db)
    finally:
        db.close()
    for i, (g, w) in enumerate(zip(got, want)):
        assert g == w, (seed, i, ops[i])
    assert final["message_count"] > 0


def test_priority_dequeue_and_flush_threshold_on_cpu(tmp_path):
    sdb, db = _db(tmp_path, priority_dequeue=True, flush_threshold=4)
    for i, p in enumerate([0, 3, 1, 3, 2, 0]):
        db.send_message("s", f"m{i}", "r", priority=sdb.MessagePriority(p))      # crosses the flush threshold
    assert [m.content for m in db.peek_messages("r", 3)] == ["m1", "m3", "m4"]
    assert [m.content for m in db.receive_messages("r", 4)] == ["m1", "m3", "m4", "m2"]
    assert [m.content for m in db.receive_messages("r", 4)] == ["m0", "m5"]
    db.close()
