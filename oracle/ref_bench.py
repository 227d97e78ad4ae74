"""Timing legs through the UNMODIFIED reference class (ORACLE / TEST INFRASTRUCTURE ONLY; used by bench.py's
`cpu_baseline_reference` leg).  BASELINE.md section 3 "reference-python": the reference's own `SwarmsDB`
(/root/reference/swarmdb/" main.py" in place, or the copy staged by oracle/build_ref.py) over the in-memory
partition-log stub of `confluent_kafka` (no broker, no network, no linger.ms: an upper bound on what the real Kafka
path could do), `KafkaConfig(num_partitions=1)`, `auto_save=False`, per-message file logging off, ONE core (the class
is single-threaded pure Python, M:393-601).

  c1          BASELINE config 1 exactly: agents "agent_a"/"agent_b", 1,000 x send_message("agent_a", content,
              "agent_b") with 128 [A-Za-z0-9] characters from default_rng(1), then agent_b drains (M:521-601).
  reduced c2  the c2 shape scaled to what the reference's O(agents x topic records) receive (M:553-601: every
              consumer parses every record) finishes in tens of seconds: 256 agents, 4 groups x 64, group sends of
              256-byte content (64-way fan-out, M:1267-1277), then EVERY agent drains.
  p50 dequeue median of receive_messages(agent, max_messages=1) with one message pending.
"""
from __future__ import annotations

import tempfile
import time

import numpy as np

ALNUM = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789"


def _content(rng, n):
    return "".join(ALNUM[i] for i in rng.integers(0, 62, n))


def run_c1(load_reference, make_reference_db):
    mod = load_reference(deterministic=False)
    rng = np.random.default_rng(1)
    bodies = [_content(rng, 128) for _ in range(1000)]
    with tempfile.TemporaryDirectory() as d:
        db = make_reference_db(mod, d, num_partitions=1)
        db.register_agent("agent_a"); db.register_agent("agent_b")
        t0 = time.perf_counter()
        for b in bodies:
            db.send_message("agent_a", b, "agent_b")
        t1 = time.perf_counter()
        got = []
        while len(got) < 1000:                      # 10 full batches of the default size (M:524): the drain loop never idles
            part = db.receive_messages("agent_b", max_messages=100, timeout=60.0)
            if not part:
                break
            got.extend(part)
        t2 = time.perf_counter()
        assert [m.content for m in got] == bodies, "reference c1 delivered something else"
        db.close()
    return {"config": "c1: 2 agents, 1000 point-to-point 128-byte messages, 1 partition", "messages": 1000,
            "send_msgs_per_s": 1000 / (t1 - t0), "receive_msgs_per_s": 1000 / (t2 - t1),
            "value": 1000 / (t2 - t0), "unit": "messages/s"}


def run_reduced_c2(load_reference, make_reference_db, budget_s: float = 25.0):
    mod = load_reference(deterministic=False)
    A, F, G = 256, 64, 4
    rng = np.random.default_rng(2)
    perm = rng.permutation(A)
    names = [f"agent_{i:07d}" for i in range(A)]
    with tempfile.TemporaryDirectory() as d:
        db = make_reference_db(mod, d, num_partitions=1)
        for n in names:
            db.register_agent(n)
        for g in range(G):
            db.add_agent_group(f"g{g}", [names[i] for i in perm[g * F:(g + 1) * F]])
        member_of = {int(i): g for g in range(G) for i in perm[g * F:(g + 1) * F]}
        body_rng = np.random.default_rng(3)
        # size the sample from a short probe so the whole leg fits the budget: send cost ~ per routed message,
        # receive cost ~ agents x routed messages (every consumer reads the whole topic, M:553-601)
        sends = 0
        t0 = time.perf_counter()
        routed = 0
        per_send = None
        max_sends = 64
        while sends < max_sends:
            g = int(rng.integers(0, G))
            s = int(rng.integers(0, A))
            while member_of[s] == g:
                s = int(rng.integers(0, A))
            ids = db.send_to_group(names[s], f"g{g}", _content(body_rng, 256))
            routed += len(ids); sends += 1
            if sends == 4:
                per_send = (time.perf_counter() - t0) / 4
                # receive estimate: A agents x routed x ~25 us; keep send + receive within the budget
                max_sends = int(max(8, min(64, budget_s / (per_send + A * F * 25e-6))))
        t1 = time.perf_counter()
        # the reference's drain loop idles until `timeout` when fewer than max_messages are pending (M:553-556); asking for
        # exactly what is pending (the local inbox record, M:456-463, knows) times the work and not the idle wait
        delivered = 0
        for n in names:
            want = len(db.agent_inbox.get(n, []))
            while want > 0:
                part = db.receive_messages(n, max_messages=min(100, want), timeout=60.0)
                delivered += len(part); want -= len(part)
                if not part:
                    break
        t2 = time.perf_counter()
        assert delivered == routed, (delivered, routed)
        # p50 dequeue: one pending message, max_messages=1
        lat = []
        for k in range(60):
            db.send_message(names[0], "x" * 128, names[1])
            t = time.perf_counter()
            got = db.receive_messages(names[1], max_messages=1, timeout=60.0)
            lat.append((time.perf_counter() - t) * 1e6)
            assert len(got) == 1
        db.close()
    return {"config": f"reduced c2: {A} agents, {G} groups x {F}, {sends} group sends x 256 bytes ({routed} routed "
                      f"messages), every agent drains", "messages": routed, "send_msgs_per_s": routed / (t1 - t0),
            "receive_msgs_per_s": routed / (t2 - t1), "value": routed / (t2 - t0), "unit": "messages/s",
            "p50_dequeue_us_topic_of_%d_records" % (routed + 60): float(np.median(lat))}


def run_all(budget_s: float = 25.0):
    from oracle import ref_loader
    if not ref_loader.reference_available():
        return None
    src = str(ref_loader.reference_path())
    c1 = run_c1(ref_loader.load_reference, ref_loader.make_reference_db)
    c2 = run_reduced_c2(ref_loader.load_reference, ref_loader.make_reference_db, budget_s)
    return {"kind": "reference", "cores": 1, "source": src,
            "how": "unmodified reference SwarmsDB class over the in-memory confluent_kafka stub (oracle/kafka_stub), "
                   "to_dict shim, auto_save off, per-message file logging off",
            "c1": c1, "reduced_c2": c2, "value": c2["value"], "unit": "messages/s",
            "sample": c2["config"]}


if __name__ == "__main__":
    import json
    print(json.dumps(run_all(), indent=1))
