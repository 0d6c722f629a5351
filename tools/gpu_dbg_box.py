"""Debug helper: full-sequence parity for unusual window sizes, with exceptions caught per case."""
import sys, os, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import __graft_entry__ as ge
import parity_common as pc
pkg = ge.load_package()
rb = pc.ref_binding()
L = pkg._lib
scene = pkg.scene.make_scene("tiny")
for order in ("ref_first", "mine_first"):
    for box, nb, cc in [(5, 1, 1), (25, 1, 1), (12, 1, 1), (11, 3, 0), (3, 1, 1), (7, 1, 1)]:
        try:
            params, mine, refs = pc.make_engines(pkg, scene, iterations=2, box=box, n_best=nb, cost_comb=cc, variants=("snapshot",))
            snap = refs["snapshot"]
            if order == "ref_first":
                snap.depthmap(7, iters=2); mine.depthmap(7)
            else:
                mine.depthmap(7); snap.depthmap(7, iters=2)
            a = pc.frac_bit_exact(mine.download(L.F_NORM4), snap.download(rb.F_NORM4))
            c = pc.frac_bit_exact(mine.download(L.F_COST), snap.download(rb.F_COST))
            print(order, box, nb, cc, "planes", a, "cost", c, flush=True)
            mine.close(); snap.close()
        except Exception as e:
            print(order, box, nb, cc, "EXC", repr(e), flush=True)
