"""Generate the committed golden fixtures by running the UNMODIFIED reference class
(/root/reference/swarmdb/" main.py") over the in-memory Kafka stub.

Run in the build container only (the reference does not travel to the GPU box):

    python tests/golden/make_golden.py

Writes tests/golden/<scenario>.json = {"ops": [...], "expected": [...], "final": {...}} or,
for bulky scenarios, {"ops_digest", "expected_digest", "n_delivered", "final"} (the ops are
regenerated from their seed by oracle/scenarios.py).  Shims applied: see oracle/ref_loader.py.
"""
import json
import sys
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

from oracle import ref_loader, scenarios  # noqa: E402


def main() -> None:
    out_dir = Path(__file__).resolve().parent
    for name, build in scenarios.SCENARIOS.items():
        ops = build()
        mod = ref_loader.load_reference(deterministic=True)
        with tempfile.TemporaryDirectory() as d:
            db = ref_loader.make_reference_db(mod, d, num_partitions=1)
            id_rank = {}
            expected = scenarios.run_ops(db, ops, mod, recv_timeout=1.0e6, id_rank=id_rank)
            final = scenarios.final_state(db)
            history = scenarios.history_state(db, id_rank)      # N2: the on-disk schema either side of the path
        n_delivered = sum(len(r) for op, r in zip(ops, expected) if op[0] == "recv")
        if name in scenarios.HASHED:
            doc = {"scenario": name, "ops_digest": scenarios.digest(ops),
                   "expected_digest": scenarios.digest(expected), "n_delivered": n_delivered, "final": final,
                   "history_digest": scenarios.digest(history)}
        else:
            doc = {"scenario": name, "ops": ops, "expected": expected, "n_delivered": n_delivered,
                   "final": final, "history": history}
        (out_dir / f"{name}.json").write_text(json.dumps(doc, indent=None, sort_keys=True) + "\n")
        print(f"{name}: {len(ops)} ops, {n_delivered} delivered")


if __name__ == "__main__":
    main()
