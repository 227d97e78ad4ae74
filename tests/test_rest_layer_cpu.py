"""the reference's REST layer (api.py, UNMODIFIED) bound to this package.  Runs only where the
reference is present (the build container); the device is the oracle-backed stand-in of tests/fake_shard.py, so
what is checked is that every `messaging_system.*` call the routes make exists here with the reference's meaning.
Route coroutines are called directly (no HTTP client library in this image)."""
import asyncio
import importlib.util
import os
import sys
from pathlib import Path

import pytest

REF_API = Path(os.environ.get("SWARMDB_REFERENCE_ROOT", "/root/reference")) / "api.py"
pytestmark = pytest.mark.skipif(not REF_API.is_file(), reason="reference api.py not present")


@pytest.fixture()
def api(tmp_path, monkeypatch):
    pytest.importorskip("fastapi"); pytest.importorskip("jwt")
    import swarmdb                      # the alias package the reference imports (api.py:29-36)
    import swarmdb_b200 as sdb
    from tests.fake_shard import OracleShard

    class CpuBackedDb(sdb.SwarmsDB):
        def __init__(self, *a, **k):
            k.setdefault("gpu_config", sdb.GpuConfig(max_agents=256, max_groups=16, deterministic_ids=True))
            super().__init__(*a, _shard=OracleShard(256, 16, 0, 1), **k)

    monkeypatch.setattr(swarmdb, "SwarmsDB", CpuBackedDb)
    monkeypatch.setenv("MESSAGE_HISTORY_DIR", str(tmp_path / "history"))
    spec = importlib.util.spec_from_file_location("reference_api_under_test", str(REF_API))
    mod = importlib.util.module_from_spec(spec)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        spec.loader.exec_module(mod)
    yield mod
    mod.shutdown_event()
    sys.modules.pop("reference_api_under_test", None)


def run(coro):
    return asyncio.run(coro)


def test_rest_routes_run_on_this_package(api, tmp_path):
    db = api.messaging_system
    assert type(db).__mro__[1].__module__ == "swarmdb_b200.core"
    tok = run(api.login_for_access_token(api.UserCredentials(username="a1", password="x")))
    from fastapi.security import HTTPAuthorizationCredentials
    assert api.get_current_agent(HTTPAuthorizationCredentials(scheme="Bearer", credentials=tok["access_token"])) == "a1"

    for a in ("a1", "a2", "a3"):
        r = run(api.register_agent(api.AgentRegistrationRequest(agent_id=a, description="d", capabilities=["c"]), agent_id=a))
        assert r == {"status": "success", "agent_id": a}
    assert db.agent_metadata["a2"]["capabilities"] == ["c"]

    sent = run(api.send_message(api.MessageRequest(content={"q": 1}, receiver_id="a2", priority=api.MessagePriorityEnum.HIGH,
                                                   metadata={"k": "v"}), current_agent="a1"))
    assert sent.sender_id == "a1" and sent.receiver_id == "a2" and sent.status.value == "delivered"
    got = run(api.receive_messages(max_messages=10, timeout=0.1, current_agent="a2"))
    assert [m.id for m in got] == [sent.id] and got[0].content == {"q": 1} and got[0].status.value == "read"
    assert got[0].metadata == {"k": "v"} and got[0].priority.value == 2

    assert run(api.create_agent_group(api.AgentGroupRequest(group_name="team", agent_ids=["a1", "a2", "a3"]), current_agent="a1"))["status"] == "success"
    ids = run(api.send_group_message(api.GroupMessageRequest(group_name="team", content="stand-up"), current_agent="a1"))
    assert len(ids) == 2
    assert [m.content for m in run(api.receive_messages(current_agent="a3"))] == ["stand-up"]
    bid = run(api.broadcast_message(api.BroadcastRequest(content="all hands", exclude_agents=["a3"]), current_agent="a1"))
    assert [m.content for m in run(api.receive_messages(current_agent="a2"))] == ["stand-up", "all hands"]
    assert run(api.receive_messages(current_agent="a3")) == []

    one = run(api.get_message(sent.id, current_agent="a2"))
    assert one.id == sent.id
    assert run(api.update_message_status(sent.id, api.MessageStatusEnum.PROCESSED, current_agent="a2"))["status"] == "success"
    assert db.get_message(sent.id).status.value == "processed"

    health = run(api.health_check())
    assert health.status == "ok" and health.kafka_connected is True          # admin_client.list_topics liveness probe
    stats = run(api.system_stats(current_agent="admin"))
    assert stats["total_messages"] == db.message_count and stats["active_agents"] == 3
    assert run(api.trigger_save(current_agent="admin"))["status"] == "success"
    assert list((tmp_path / "history").glob("message_history_*.json"))
    assert run(api.resend_failed_messages(current_agent="admin"))["status"] == "success"
    assert run(api.flush_old_messages(older_than=None, current_agent="admin"))["status"] == "success"
    assert run(api.auto_scale_partitions(current_agent="admin"))["status"] == "success"
    assert run(api.deregister_agent("a3", current_agent="a3")) == {"status": "success", "agent_id": "a3"}
    assert bid["message_id"] in db.messages          # (the route returns a dict although it declares List[str], A:507/530)
