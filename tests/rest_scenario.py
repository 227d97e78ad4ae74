"""One REST scenario driven over HTTP (starlette TestClient) against `swarmdb_b200.server.create_app(db)`; used on CPU
with the oracle-backed stand-in shard (tests/test_server_cpu.py) and on the device (tests/test_gpu_server.py)."""


def drive(app, db):
    from fastapi.testclient import TestClient
    c = TestClient(app)

    def token(user):
        r = c.post("/auth/token", json={"username": user, "password": "pw"})
        assert r.status_code == 200 and r.json()["token_type"] == "bearer"
        return {"Authorization": f"Bearer {r.json()['access_token']}"}
    a1, a2, a3, adm = token("a1"), token("a2"), token("a3"), token("admin")
    assert c.post("/auth/token", json={"username": "", "password": ""}).status_code == 401
    assert c.post("/agents/receive", headers={"Authorization": "Bearer garbage"}).status_code == 401
    assert c.get("/health").json()["kafka_connected"] is True

    for who, hdr in (("a1", a1), ("a2", a2), ("a3", a3)):
        r = c.post("/agents/register", json={"agent_id": who, "description": "d", "capabilities": ["c"]}, headers=hdr)
        assert r.status_code == 201 and r.json() == {"status": "success", "agent_id": who}
    assert c.post("/agents/register", json={"agent_id": "a2"}, headers=a1).status_code == 403
    assert app.state.agent_metadata["a2"]["capabilities"] == ["c"]

    sent = c.post("/messages", json={"content": {"q": 1}, "receiver_id": "a2", "priority": 2, "metadata": {"k": "v"}}, headers=a1)
    assert sent.status_code == 200
    sent = sent.json()
    assert sent["sender_id"] == "a1" and sent["receiver_id"] == "a2" and sent["status"] == "delivered" and sent["priority"] == 2
    load = c.get("/agents/a2/load", headers=a2).json()                      # answered by the device queue (N3)
    assert load["unread_count"] == 1 and load["inbox_size"] == 1 and load["unread_by_priority"] == [0, 0, 1, 0]
    assert c.get("/agents/a2/load", headers=a1).status_code == 403
    got = c.post("/agents/receive", params={"max_messages": 10, "timeout": 0.1}, headers=a2).json()
    assert [m["id"] for m in got] == [sent["id"]] and got[0]["content"] == {"q": 1} and got[0]["status"] == "read"
    assert got[0]["metadata"] == {"k": "v"}
    assert c.get("/agents/a2/load", headers=a2).json()["unread_count"] == 0

    assert c.post("/groups", json={"group_name": "team", "agent_ids": ["a1", "a2", "a3"]}, headers=a1).status_code == 201
    # the two routes that always answer 500 in the reference (response_model=List[str] vs a dict, A:507/530, A:760/781)
    g = c.post("/groups/message", json={"group_name": "team", "content": "stand-up"}, headers=a1)
    assert g.status_code == 200 and g.json()["status"] == "success" and len(g.json()["message_ids"]) == 2
    b = c.post("/messages/broadcast", json={"content": "all hands", "exclude_agents": ["a3"]}, headers=a1)
    assert b.status_code == 200 and b.json()["status"] == "success" and b.json()["message_id"] in db.messages
    assert [m["content"] for m in c.post("/agents/receive", headers=a3).json()] == ["stand-up"]
    assert [m["content"] for m in c.post("/agents/receive", headers=a2).json()] == ["stand-up", "all hands"]
    assert c.post("/agents/receive", headers=a3).json() == []

    # routes whose `status` parameter shadows fastapi.status in the reference: their 403 / 404 branches work here
    assert c.get(f"/messages/{sent['id']}", headers=a2).json()["id"] == sent["id"]
    assert c.get("/messages/nope", headers=a2).status_code == 404
    q = c.get("/messages", params={"sender_id": "a1", "status": "read", "limit": 10}, headers=a1)
    assert q.status_code == 200 and sent["id"] in [m["id"] for m in q.json()]
    assert c.get("/messages", params={"sender_id": "a1"}, headers=a3).status_code == 403       # not 500
    mine = c.get("/agents/a2/messages", params={"status": "read"}, headers=a2)
    assert mine.status_code == 200 and all(m["status"] == "read" for m in mine.json()) and len(mine.json()) >= 2
    assert c.get("/agents/a2/messages", headers=a3).status_code == 403
    assert c.put(f"/messages/{sent['id']}/status", params={"status": "processed"}, headers=a3).status_code == 403
    assert c.put("/messages/nope/status", params={"status": "processed"}, headers=a2).status_code == 404
    assert c.put(f"/messages/{sent['id']}/status", params={"status": "processed"}, headers=a2).json()["status"] == "success"
    assert db.get_message(sent["id"]).status.value == "processed"

    assert c.get("/stats", headers=a1).status_code == 403
    st = c.get("/stats", headers=adm).json()
    assert st["total_messages"] == db.message_count and st["active_agents"] == 3
    qs = c.get("/queue", headers=adm).json()
    assert qs["pending"] == 0 and qs["received"] == 4
    assert c.post("/admin/save", headers=adm).json()["status"] == "success"
    assert c.post("/admin/resend_failed", headers=adm).json()["resent_count"] == 0
    assert c.post("/admin/flush", headers=adm).json()["status"] == "success"
    assert c.post("/admin/scale_partitions", headers=adm).json()["status"] == "success"
    assert c.post("/admin/save", headers=a1).status_code == 403
    assert c.delete("/agents/a3", headers=a3).json() == {"status": "success", "agent_id": "a3"}
    return True
