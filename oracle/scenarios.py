"""Scenario scripts shared by the golden-fixture generator and the parity tests
(TEST INFRASTRUCTURE ONLY - nothing in the shipped product imports this).

A scenario is a JSON-able list of operations against the `SwarmsDB` Python surface
(SURVEY.md section 8b(i)).  `run_ops` executes it against ANY object exposing that surface - the
unmodified reference class (via `oracle/ref_loader.py`), the pure-Python restatement
(`oracle/pyref.py`) or the GPU-backed `swarmdb_b200.SwarmsDB` - and returns results in a
canonical, implementation-independent form:

  * message ids are replaced by their rank of first appearance (the reference's ids are
    uuid4, M:72; with the deterministic-uuid shim they are a counter, and so are ours);
  * `timestamp` is dropped (M:78 binds the real `time.time` as a pydantic default_factory
    at class-definition time, so it cannot be made deterministic without editing the class;
    it is carried, never compared - SURVEY.md section 8a R2);
  * `visible_to` is compared as a sorted list (set-iteration order, Appendix A rule 4).
"""
from __future__ import annotations

import hashlib
import json
import os
from typing import Any, Dict, List, Optional

import numpy as np

_ALNUM = np.frombuffer(b"ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789", dtype=np.uint8)

TYPES = ["chat", "command", "function_call", "function_result", "system", "error", "status"]


def alnum(rng: np.random.Generator, n: int) -> str:
    """`n` characters drawn from [A-Za-z0-9] (SURVEY.md section 8d content generator)."""
    return _ALNUM[rng.integers(0, len(_ALNUM), size=n)].tobytes().decode("ascii")


# --------------------------------------------------------------------------- scenario builders
def scenario_example_main() -> List[list]:
    """The reference's only behavioural fixture: the `__main__` example, M:1398-1453."""
    return [
        ["register", "agent1"], ["register", "agent2"], ["register", "agent3"],
        ["send", "agent1", "Hello, Agent 2!", "agent2", "chat", 1, None, None],
        ["send", "agent2", "Hi, Agent 1! How are you?", "agent1", "chat", 1, None, None],
        ["broadcast", "agent3", "Important announcement for everyone!", "chat", 1, None, None],
        ["recv", "agent1", 100],
        ["group", "team_alpha", ["agent1", "agent2"]],
        ["send_group", "agent3", "team_alpha", "Message for Team Alpha!", "chat", 1, None],
        ["recv", "agent1", 100], ["recv", "agent2", 100], ["recv", "agent3", 100],
    ]


def scenario_appendix_a() -> List[list]:
    """One hand-written case per rule of SURVEY.md Appendix A."""
    ops: List[list] = []
    # rule 1: auto-registration of sender and receiver by send; of the caller by recv
    ops += [["send", "s0", "auto-reg", "r0", "chat", 1, None, None], ["recv", "r0", 100],
            ["recv", "never_seen", 100]]
    # rule 3: p2p visibility - receiver listed / not listed in visible_to
    ops += [["send", "s0", "vis-yes", "r0", "chat", 1, None, ["r0", "x"]],
            ["send", "s0", "vis-no", "r0", "chat", 1, None, ["x"]],
            ["send", "s0", "vis-empty", "r0", "chat", 1, None, []],
            ["recv", "r0", 100], ["recv", "x", 100]]
    # rule 4: default broadcast reaches every registered agent INCLUDING the sender
    ops += [["send", "s0", "bcast-default", None, "system", 2, None, None],
            ["recv", "s0", 100], ["recv", "r0", 100], ["recv", "x", 100]]
    # rule 4: broadcast_message excludes the sender and the exclude list
    ops += [["broadcast", "s0", "bcast-excl", "status", 0, {"k": "v"}, ["x"]],
            ["recv", "s0", 100], ["recv", "r0", 100], ["recv", "x", 100], ["recv", "never_seen", 100]]
    # rule 4/12: late joiner never sees an earlier broadcast
    ops += [["register", "late"], ["recv", "late", 100]]
    # broadcast restricted with explicit visible_to through send_message
    ops += [["send", "s0", "bcast-vis", None, "chat", 1, None, ["late", "r0"]],
            ["recv", "late", 100], ["recv", "r0", 100], ["recv", "x", 100]]
    # rule 5/6: group store overwrite, duplicates, skip-sender, unknown group, order = member order
    ops += [["group", "g", ["m1", "m2", "m3"]],
            ["send_group", "m2", "g", "grp-1", "command", 3, {"a": 1}],
            ["group", "g", ["m3", "m1", "m1", "m2"]],
            ["send_group", "outsider", "g", "grp-2", "chat", 1, None],
            ["send_group", "m1", "nope", "grp-unknown", "chat", 1, None],
            ["recv", "m1", 100], ["recv", "m2", 100], ["recv", "m3", 100], ["recv", "outsider", 100]]
    # rule 8: dequeue continues where the previous call stopped; max_messages honoured
    for i in range(7):
        ops.append(["send", "p", f"seq-{i}", "q", "chat", i % 4, None, None])
    ops += [["recv", "q", 3], ["recv", "q", 3], ["recv", "q", 3], ["recv", "q", 3]]
    # rule 9: priority never reorders (LOW, CRITICAL, NORMAL come back in send order)
    ops += [["send", "p", "low", "z", "chat", 0, None, None],
            ["send", "p", "critical", "z", "chat", 3, None, None],
            ["send", "p", "normal", "z", "chat", 1, None, None], ["recv", "z", 100]]
    # rule 11: deregister keeps offsets; re-registration through send resumes, nothing re-delivered
    ops += [["send", "p", "before-dereg", "d", "chat", 1, None, None], ["recv", "d", 100],
            ["deregister", "d"], ["deregister", "ghost"],
            ["send", "p", "after-dereg", "d", "chat", 1, None, None], ["recv", "d", 100]]
    # a broadcast while an agent is deregistered is not visible to it
    ops += [["deregister", "q"], ["send", "p", "bcast-while-q-away", None, "chat", 1, None, None],
            ["recv", "q", 100], ["recv", "z", 100]]
    # content kinds: dict, list, unicode, empty string, long string
    ops += [["send", "p", {"k": [1, 2, {"n": None}], "t": True}, "c", "function_call", 2, {"m": [1]}, None],
            ["send", "p", [1, "two", 3.5, None], "c", "function_result", 1, None, None],
            ["send", "p", "héllo 世界 \U0001f600", "c", "chat", 1, None, None],
            ["send", "p", "", "c", "chat", 1, None, None],
            ["send", "p", "L" * 1000, "c", "chat", 1, None, None],
            ["recv", "c", 100]]
    return ops


def scenario_random(seed: int, n_agents: int = 24, n_ops: int = 400) -> List[list]:
    """Seeded random interleaving of every operation kind."""
    rng = np.random.default_rng(seed)
    agents = [f"a{i:02d}" for i in range(n_agents)]
    groups: List[str] = []
    ops: List[list] = []
    for a in agents[: n_agents // 2]:
        ops.append(["register", a])

    def pick() -> str:
        return agents[int(rng.integers(0, n_agents))]

    for _ in range(n_ops):
        r = float(rng.random())
        typ = TYPES[int(rng.integers(0, len(TYPES)))]
        prio = int(rng.integers(0, 4))
        content: Any = alnum(rng, int(rng.integers(1, 200)))
        if rng.random() < 0.1:
            content = {"text": content[:20], "n": int(rng.integers(0, 1000))}
        md: Optional[Dict[str, Any]] = {"tag": int(rng.integers(0, 9))} if rng.random() < 0.2 else None
        if r < 0.40:
            vis = None
            if rng.random() < 0.08:
                vis = sorted({pick() for _ in range(3)})
            ops.append(["send", pick(), content, pick(), typ, prio, md, vis])
        elif r < 0.55 and groups:
            ops.append(["send_group", pick(), groups[int(rng.integers(0, len(groups)))], content, typ, prio, md])
        elif r < 0.62:
            name = f"g{int(rng.integers(0, 6))}"
            members = [pick() for _ in range(int(rng.integers(1, 9)))]
            ops.append(["group", name, members])
            if name not in groups:
                groups.append(name)
        elif r < 0.68:
            excl = [pick()] if rng.random() < 0.5 else None
            ops.append(["broadcast", pick(), content, typ, prio, md, excl])
        elif r < 0.70:
            ops.append(["send", pick(), content, None, typ, prio, md, None])
        elif r < 0.73:
            ops.append(["deregister", pick()])
        elif r < 0.76:
            ops.append(["register", pick()])
        else:
            ops.append(["recv", pick(), int(rng.choice([1, 2, 5, 100]))])
    for a in agents:
        ops.append(["recv", a, 1000])
    return ops


def scenario_c1(n_msgs: int = 1000, content_len: int = 128) -> List[list]:
    """BASELINE config 1 (SURVEY.md section 8d 'c1'): 2 agents, 1k p2p 128-byte messages, one drain."""
    rng = np.random.default_rng(1)
    ops: List[list] = [["register", "agent_a"], ["register", "agent_b"]]
    for _ in range(n_msgs):
        ops.append(["send", "agent_a", alnum(rng, content_len), "agent_b", "chat", 1, None, None])
    ops.append(["recv", "agent_b", n_msgs + 10])
    ops.append(["recv", "agent_a", 10])
    return ops


def scenario_group_fanout(seed: int = 7, n_agents: int = 128, group_size: int = 16,
                          n_sends: int = 200, max_len: int = 256) -> List[list]:
    """Reduced BASELINE config 2: disjoint groups, variable-length alnum payloads, 4 priorities."""
    rng = np.random.default_rng(seed)
    agents = [f"agent_{i:07d}" for i in range(n_agents)]
    perm = rng.permutation(n_agents)
    n_groups = n_agents // group_size
    ops: List[list] = []
    for g in range(n_groups):
        ops.append(["group", f"grp{g}", [agents[int(i)] for i in perm[g * group_size:(g + 1) * group_size]]])
    for _ in range(n_sends):
        g = int(rng.integers(0, n_groups))
        sender = agents[int(rng.integers(0, n_agents))]      # may be a member: skip-sender rule exercised
        ops.append(["send_group", sender, f"grp{g}", alnum(rng, int(rng.integers(1, max_len + 1))),
                    TYPES[int(rng.integers(0, 7))], int(rng.integers(0, 4)), None])
    for a in agents:
        ops.append(["recv", a, 7])
    for a in agents:
        ops.append(["recv", a, 1000])
    return ops


def scenario_bookkeeping(seed: int = 5, n_agents: int = 10, n_ops: int = 260) -> List[list]:
    """The host-side bookkeeping methods of the surface (M:603-850, M:973-1206) interleaved with traffic:
    get_message, get_agent_messages, mark_message_as_processed, query_messages, search_messages,
    get_conversation, get_unread_message_count, get_agent_load, get_stats, delete_message,
    flush_old_messages, resend_failed_messages.  Message arguments are RANKS (first-seen order of ids);
    timestamp bounds are given as the rank of the message whose timestamp is the bound, and are always
    point-to-point or broadcast messages: the copies of one group send are stamped microseconds apart by
    the reference (independent send_message calls, M:1267-1277) and share one stamp in a batched fan-out."""
    rng = np.random.default_rng(seed)
    agents = [f"b{i}" for i in range(n_agents)]
    team = agents[2:7]
    ops: List[list] = [["register", a] for a in agents[:6]]
    ops.append(["group", "team", team])
    solo: List[int] = []              # ranks of p2p / broadcast messages
    n_ranks = 0

    def pick() -> str:
        return agents[int(rng.integers(0, n_agents))]

    def rank() -> int:
        return int(rng.integers(0, max(n_ranks, 1)))

    def bound() -> Optional[int]:
        return solo[int(rng.integers(0, len(solo)))] if solo else None

    words = ["alpha", "Beta", "gamma", "DELTA", "eps"]
    for _ in range(n_ops):
        r = float(rng.random())
        typ = TYPES[int(rng.integers(0, len(TYPES)))]
        text = " ".join(words[int(k)] for k in rng.integers(0, len(words), 3))
        content: Any = text if rng.random() < 0.8 else {"note": text, "k": [1, 2]}
        if r < 0.30:
            ops.append(["send", pick(), content, pick(), typ, int(rng.integers(0, 4)), None, None])
            solo.append(n_ranks); n_ranks += 1
        elif r < 0.36:
            sender = pick()
            ops.append(["send_group", sender, "team", content, typ, 1, {"t": 1}])
            n_ranks += len(team) - (sender in team)
        elif r < 0.40:
            ops.append(["broadcast", pick(), content, typ, 2, None, None])
            solo.append(n_ranks); n_ranks += 1
        elif r < 0.50:
            ops.append(["recv", pick(), int(rng.choice([1, 3, 100]))])
        elif r < 0.55:
            ops.append(["get_message", rank()])
        elif r < 0.62:
            st = [None, "delivered", "read", "processed"][int(rng.integers(0, 4))]
            ops.append(["agent_messages", pick(), st, int(rng.choice([2, 5, 100])), int(rng.integers(0, 3))])
        elif r < 0.67:
            ops.append(["mark_processed", rank() if rng.random() < 0.8 else -1])
        elif r < 0.76:
            q = {"sender": pick() if rng.random() < 0.5 else None, "receiver": pick() if rng.random() < 0.5 else None,
                 "type": typ if rng.random() < 0.3 else None,
                 "status": [None, "delivered", "read", "processed", "failed"][int(rng.integers(0, 5))],
                 "after": bound() if rng.random() < 0.3 else None, "before": bound() if rng.random() < 0.3 else None,
                 "limit": int(rng.choice([1, 4, 100]))}
            ops.append(["query", q])
        elif r < 0.81:
            ops.append(["search", words[int(rng.integers(0, len(words)))].lower(), bool(rng.random() < 0.5), int(rng.choice([3, 100]))])
        elif r < 0.85:
            ops.append(["conversation", pick(), pick(), int(rng.choice([2, 7, 100]))])
        elif r < 0.89:
            ops.append(["unread", pick()])
        elif r < 0.92:
            ops.append(["agent_load", pick()])
        elif r < 0.95:
            ops.append(["stats"])
        elif r < 0.97:
            ops.append(["delete", rank() if rng.random() < 0.8 else -1])
        else:
            ops.append(["force_status", rank(), "failed"])
    # resends create ids whose number depends on run-time state: keep them after every rank-addressed op
    ops += [["flush_old", bound()], ["stats"], ["force_status", solo[0], "failed"], ["resend_failed"], ["stats"], ["flush_old", None]]
    for a in agents:
        ops.append(["recv", a, 1000])
    ops.append(["stats"])
    return ops


SCENARIOS = {
    "example_main": scenario_example_main,
    "appendix_a": scenario_appendix_a,
    "random_1": lambda: scenario_random(1),
    "random_2": lambda: scenario_random(2),
    "random_3": lambda: scenario_random(3, n_agents=40, n_ops=600),
    "c1_p2p_1k": scenario_c1,
    "group_fanout_small": scenario_group_fanout,
    "bookkeeping": scenario_bookkeeping,
}
# scenarios that exercise host-side bookkeeping methods only the full surface has (not the path oracles)
SURFACE_ONLY = {"bookkeeping"}
PATH_SCENARIOS = sorted(set(SCENARIOS) - SURFACE_ONLY)
# scenarios whose full expected output is too bulky to commit: only a digest is stored
HASHED = {"c1_p2p_1k", "group_fanout_small"}


# --------------------------------------------------------------------------- runner
def _enum_val(x: Any) -> Any:
    return getattr(x, "value", x)


def canon_message(m: Any, id_rank: Dict[str, int]) -> list:
    rank = id_rank.setdefault(m.id, len(id_rank))
    return [rank, m.sender_id, m.receiver_id, m.content, _enum_val(m.type), int(_enum_val(m.priority)),
            _enum_val(m.status), m.metadata, m.token_count, sorted(m.visible_to)]


def run_ops(db: Any, ops: List[list], enums: Any, recv_timeout: float = 1.0e6,
            id_rank: Dict[str, int] = None) -> List[Any]:
    """Execute `ops` on `db`; `enums` provides MessageType / MessagePriority for that db.
    Pass an `id_rank` dict to keep the id -> first-seen-rank map (history_state needs it)."""
    id_rank = {} if id_rank is None else id_rank
    out: List[Any] = []
    for op in ops:
        kind = op[0]
        if kind == "register":
            db.register_agent(op[1]); out.append(None)
        elif kind == "deregister":
            db.deregister_agent(op[1]); out.append(None)
        elif kind == "send":
            _, sender, content, receiver, typ, prio, md, vis = op
            mid = db.send_message(sender, _copy(content), receiver, enums.MessageType(typ),
                                  enums.MessagePriority(prio), _copy(md), _copy(vis))
            out.append(id_rank.setdefault(mid, len(id_rank)))
        elif kind == "broadcast":
            _, sender, content, typ, prio, md, excl = op
            mid = db.broadcast_message(sender, _copy(content), enums.MessageType(typ),
                                       enums.MessagePriority(prio), _copy(md), _copy(excl))
            out.append(id_rank.setdefault(mid, len(id_rank)))
        elif kind == "group":
            db.add_agent_group(op[1], list(op[2])); out.append(None)
        elif kind == "send_group":
            _, sender, group, content, typ, prio, md = op
            ids = db.send_to_group(sender, group, _copy(content), enums.MessageType(typ),
                                   enums.MessagePriority(prio), _copy(md))
            out.append([id_rank.setdefault(i, len(id_rank)) for i in ids])
        elif kind == "recv":
            msgs = db.receive_messages(op[1], max_messages=op[2], timeout=recv_timeout)
            out.append([canon_message(m, id_rank) for m in msgs])
        else:
            out.append(_bookkeeping_op(db, op, enums, id_rank))
    return out


def _bookkeeping_op(db: Any, op: list, enums: Any, id_rank: Dict[str, int]) -> Any:
    """Host-side bookkeeping methods (M:603-850, M:973-1206); message arguments are ranks."""
    kind = op[0]
    by_rank = {r: i for i, r in id_rank.items()}

    def mid(r: Any) -> str:
        return by_rank.get(r, "no-such-message-id")

    def ranks(msgs: List[Any]) -> list:
        return [[id_rank.setdefault(m.id, len(id_rank)), _enum_val(m.status)] for m in msgs]

    def stamp(r: Any) -> Any:
        m = db.get_message(mid(r)) if r is not None else None
        return None if m is None else m.timestamp

    if kind == "get_message":
        m = db.get_message(mid(op[1]))
        return None if m is None else canon_message(m, id_rank)
    if kind == "agent_messages":
        st = None if op[2] is None else enums.MessageStatus(op[2])
        return ranks(db.get_agent_messages(op[1], status=st, limit=op[3], skip=op[4]))
    if kind == "mark_processed":
        return db.mark_message_as_processed(mid(op[1]))
    if kind == "query":
        q = op[1]
        return ranks(db.query_messages(
            sender_id=q["sender"], receiver_id=q["receiver"],
            message_type=None if q["type"] is None else enums.MessageType(q["type"]),
            status=None if q["status"] is None else enums.MessageStatus(q["status"]),
            after_timestamp=stamp(q["after"]), before_timestamp=stamp(q["before"]), limit=q["limit"]))
    if kind == "search":
        return ranks(db.search_messages(op[1], case_sensitive=op[2], limit=op[3]))
    if kind == "conversation":
        return ranks(db.get_conversation(op[1], op[2], limit=op[3]))
    if kind == "unread":
        return db.get_unread_message_count(op[1])
    if kind == "agent_load":
        d = dict(db.get_agent_load(op[1]))
        d["processing_rate"] = round(d["processing_rate"] * 60)          # messages in the last minute (all of them here)
        return d
    if kind == "stats":
        st = db.get_stats()
        return {k: st[k] for k in ("total_messages", "active_agents", "messages_by_type", "messages_by_status")} | {
            "messages_by_agent": {a: st["messages_by_agent"][a] for a in sorted(st["messages_by_agent"])}}
    if kind == "delete":
        return db.delete_message(mid(op[1]))
    if kind == "flush_old":
        return db.flush_old_messages(stamp(op[1]) if op[1] is not None else None)
    if kind == "force_status":            # white-box: what a failed produce leaves behind (M:501-519)
        m = db.get_message(mid(op[1]))
        if m is not None:
            m.status = enums.MessageStatus(op[2])
        return m is not None
    if kind == "resend_failed":
        return [id_rank.setdefault(i, len(id_rank)) for i in db.resend_failed_messages()]
    raise ValueError(kind)  # pragma: no cover


def _copy(x: Any) -> Any:
    return json.loads(json.dumps(x)) if isinstance(x, (dict, list)) else x


def final_state(db: Any) -> Dict[str, Any]:
    """Inbox lengths + registered set - the reference's local side record (M:453-463)."""
    return {
        "registered": sorted(db.registered_agents),
        "inbox_len": {a: len(v) for a, v in sorted(db.agent_inbox.items())},
        "message_count": db.message_count,
    }


def history_state(db: Any, id_rank: Dict[str, int]) -> Dict[str, Any]:
    """The history file `save_message_history` writes (schema M:878-884), canonicalised: message ids
    replaced by their first-seen rank, timestamps dropped, sets sorted.  `keys` pins the field order
    of a stored message (M:54-111)."""
    name = "golden_history_probe.json"
    db.save_message_history(name)
    path = os.path.join(str(db.save_dir), name)
    with open(path) as f:
        h = json.load(f)
    os.remove(path)
    rank = lambda mid: id_rank.setdefault(mid, len(id_rank))          # noqa: E731
    msgs = {}
    keys = None
    for mid, d in h["messages"].items():
        assert d["id"] == mid
        keys = keys or list(d.keys())
        msgs[str(rank(mid))] = [d["sender_id"], d["receiver_id"], d["content"], d["type"], d["priority"], d["status"],
                                d["metadata"], d["token_count"], sorted(d["visible_to"])]
    return {"top_keys": sorted(h.keys()), "keys": keys, "messages": msgs,
            "agent_inbox": {a: [rank(i) for i in v] for a, v in sorted(h["agent_inbox"].items())},
            "registered_agents": sorted(h["registered_agents"]), "message_count": h["message_count"]}


def digest(obj: Any) -> str:
    return hashlib.sha256(json.dumps(obj, sort_keys=True, ensure_ascii=True).encode()).hexdigest()
