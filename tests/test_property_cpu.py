"""Property test: random interleavings of register / send / group / broadcast / receive
must give the same delivered streams in the two independent restatements - the surface-level
Python oracle (pinned to the reference's goldens) and the ABI-level C oracle."""
import json

from hypothesis import HealthCheck, given, settings, strategies as st

from oracle import pyref, scenarios
from tests.test_oracle_c import _rank_ids, _reduce_expected, _run_c

AGENTS = [f"a{i}" for i in range(6)]
GROUPS = ["g0", "g1"]

agent = st.sampled_from(AGENTS)
content = st.one_of(st.text(alphabet="abcXYZ019 é", max_size=40), st.dictionaries(st.sampled_from(["k", "n"]), st.integers(0, 9), max_size=2))
typ = st.sampled_from(scenarios.TYPES)
prio = st.integers(0, 3)

op = st.one_of(
    st.tuples(st.just("register"), agent).map(list),
    st.tuples(st.just("deregister"), agent).map(list),
    st.tuples(st.just("send"), agent, content, agent, typ, prio, st.none(), st.one_of(st.none(), st.lists(agent, max_size=3))).map(list),
    st.tuples(st.just("send"), agent, content, st.none(), typ, prio, st.none(), st.one_of(st.none(), st.lists(agent, max_size=3))).map(list),
    st.tuples(st.just("broadcast"), agent, content, typ, prio, st.none(), st.one_of(st.none(), st.lists(agent, max_size=2))).map(list),
    st.tuples(st.just("group"), st.sampled_from(GROUPS), st.lists(agent, max_size=5)).map(list),
    st.tuples(st.just("send_group"), agent, st.sampled_from(GROUPS + ["nope"]), content, typ, prio, st.none()).map(list),
    st.tuples(st.just("recv"), agent, st.sampled_from([1, 2, 100])).map(list),
)


@settings(max_examples=60, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(st.lists(op, max_size=40))
def test_python_and_c_restatements_agree(ops):
    ops = ops + [["recv", a, 1000] for a in AGENTS]
    db = pyref.OracleSwarmsDB(id_factory=pyref.counter_ids())
    expected = json.loads(json.dumps(scenarios.run_ops(db, ops, pyref)))
    want = _rank_ids(_reduce_expected(expected, ops))
    got = _rank_ids(_run_c(ops))
    for i, (g, w) in enumerate(zip(got, want)):
        assert g == w, (i, ops[i])
