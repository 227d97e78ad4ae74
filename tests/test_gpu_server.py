"""SURVEY 8f N4 on the device: swarmdb_b200/server.py over HTTP, backed by a real GPU shard (ctypes -> C ABI -> kernels)."""
import pytest

pytestmark = pytest.mark.gpu


def test_rest_front_end_on_a_gpu_shard(tmp_path):
    pytest.importorskip("fastapi"); pytest.importorskip("jwt"); pytest.importorskip("httpx")
    import swarmdb_b200 as sdb
    from swarmdb_b200._native import Shard
    from swarmdb_b200.server import Settings, create_app
    from tests.rest_scenario import drive
    db = sdb.SwarmsDB(save_dir=str(tmp_path), auto_save=False,
                      gpu_config=sdb.GpuConfig(max_agents=256, max_groups=16, ring_slots=256, arena_bytes=1 << 24, deterministic_ids=True))
    assert isinstance(db.shard, Shard)
    assert drive(create_app(db, Settings(history_dir=str(tmp_path))), db)
    assert db.shard.stats()["kernel_launches"] > 0
    db.close()
