"""cProfile of the c1 scenario through the Python surface (where does host time go?)."""
import cProfile, pstats, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import swarmdb_b200 as sdb
from bench import ALNUM
rng = np.random.default_rng(1)
contents = [ALNUM[rng.integers(0, 62, 128)].tobytes().decode() for _ in range(1000)]
db = sdb.SwarmsDB(save_dir="/tmp/sdb_c1p", auto_save=False, gpu_config=sdb.GpuConfig(max_agents=1024, ring_slots=2048, arena_bytes=1 << 26))
for c in contents[:50]:
    db.send_message("agent_a", c, "agent_b")
db.receive_messages("agent_b", 100)
pr = cProfile.Profile(); pr.enable()
for c in contents:
    db.send_message("agent_a", c, "agent_b")
got = db.receive_messages("agent_b", 2000)
pr.disable()
assert len(got) == 1000
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
db.close()
