"""Cost of a device-wide barrier on this GPU: the barrier of k_relax2 against cluster-hierarchical arrival (kmc_barrier_bench)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kmcislandgrowth_b200 import runtime, workloads
gv = workloads.constants_for(18, 10, 10)
ctx = runtime.Context(runtime.params_from(gv), 0)
out = {}
for variant, cluster in ((0, 1), (2, 1), (3, 1), (4, 1), (1, 2), (1, 4)):
    try:
        ctx.barrier_bench(variant, cluster, 2000)
        us, stale = ctx.barrier_bench(variant, cluster, 50000)
        out["variant%d_cluster%d" % (variant, cluster)] = {"us_per_barrier": round(us, 3), "stale_reads": stale}
    except Exception as e:
        out["variant%d_cluster%d" % (variant, cluster)] = {"error": str(e)[:160]}
print(json.dumps(out))
