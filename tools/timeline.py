"""Development helper: per-phase timeline of one create_proof (CUDA events), to spot idle gaps."""
import os
import sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402
pkg = graft.load_package()
ctx = pkg.Context(0)
k, cols = int(os.environ.get("K", 17)), int(os.environ.get("COLS", 3))
w = pkg.workload.build(ctx, k, cols, seed=1)
for i in range(3):
    pkg.create_proof(w.pk, w.advice_dev, w.instances, pkg.seed_from_u64(i))
ctx.set_overlap(os.environ.get("OVERLAP", "1") == "1")
ctx.profile_enable(True)
ctx.profile_report()
pkg.create_proof(w.pk, w.advice_dev, w.instances, pkg.seed_from_u64(9))
tl = ctx.profile_timeline()
end = 0.0
for name, start, dur in tl:
    gap = start - end
    print("%9.3f %8.3f  %s%s" % (start, dur, name, ("   <-- gap %.3f" % gap) if gap > 0.03 and not name.startswith("prove.") else ""))
    if not name.startswith("prove."):
        end = max(end, start + dur)
