"""the package's own REST front-end (swarmdb_b200/server.py) over HTTP, on CPU - the device is the
oracle-backed stand-in of tests/fake_shard.py (tests only).  tests/test_gpu_server.py repeats it on a real shard;
tests/test_rest_layer_cpu.py keeps checking that the reference's UNMODIFIED api.py runs on this package."""
import pytest

from tests.fake_shard import OracleShard
from tests.rest_scenario import drive


def test_rest_front_end_over_http(tmp_path):
    pytest.importorskip("fastapi"); pytest.importorskip("jwt"); pytest.importorskip("httpx")
    import swarmdb_b200 as sdb
    from swarmdb_b200.server import Settings, create_app
    db = sdb.SwarmsDB(save_dir=str(tmp_path), auto_save=False, _shard=OracleShard(256, 16, 0, 1),
                      gpu_config=sdb.GpuConfig(max_agents=256, max_groups=16, deterministic_ids=True))
    app = create_app(db, Settings(history_dir=str(tmp_path)))
    assert drive(app, db)
    paths = {(m, r.path) for r in app.routes if hasattr(r, "methods") for m in r.methods if m in ("GET", "POST", "PUT", "DELETE")}
    reference_routes = {("POST", "/auth/token"), ("POST", "/agents/register"), ("DELETE", "/agents/{agent_id}"), ("POST", "/messages"),
                        ("POST", "/messages/broadcast"), ("GET", "/messages/{message_id}"), ("GET", "/messages"),
                        ("GET", "/agents/{agent_id}/messages"), ("POST", "/agents/receive"), ("PUT", "/messages/{message_id}/status"),
                        ("POST", "/groups"), ("POST", "/groups/message"), ("GET", "/health"), ("GET", "/stats"), ("POST", "/admin/save"),
                        ("POST", "/admin/flush"), ("POST", "/admin/resend_failed"), ("POST", "/admin/scale_partitions")}
    assert reference_routes <= paths                                       # every route of api.py:365-935 is bound
    db.close()


def test_rate_limit_answers_429(tmp_path):
    pytest.importorskip("httpx")
    import swarmdb_b200 as sdb
    from fastapi.testclient import TestClient
    from swarmdb_b200.server import Settings, create_app
    db = sdb.SwarmsDB(save_dir=str(tmp_path), auto_save=False, _shard=OracleShard(64, 4, 0, 1),
                      gpu_config=sdb.GpuConfig(max_agents=64, max_groups=4, deterministic_ids=True))
    c = TestClient(create_app(db, Settings(history_dir=str(tmp_path), rate_limit_per_minute=5)))
    codes = [c.get("/health").status_code for _ in range(8)]
    assert codes[:5] == [200] * 5 and codes[5:] == [429] * 3
    db.close()
