"""Bracket statistics of the robust-scale selections (library built with `make prof`, SVO_B200_LIB pointing at it): how wide
the hot brackets are, how often they hit and how many keys they hold -- the data behind the tier policy in cluster_select.cuh."""
import importlib, sys, numpy as np
sys.path.insert(0, "/root/repo")
pkg = importlib.import_module("semi-direct-visual-odometry_b200")
mode = int(sys.argv[1]) if len(sys.argv) > 1 else 2
tot = np.zeros(20, np.int64)
evals = 0
for idx in range(12):
    pair = pkg.synth.make_pair(idx, 500)
    with pkg.Context(pair["w"], pair["h"], pair["K"], levels=4, max_frames=2, max_jobs=1, max_features=512) as ctx:
        ctx.upload(0, np.stack([pair["ref"], pair["cur"]]))
        j = pkg.capi.make_jobs(1)
        j[0]["ref_slot"], j[0]["kf_slot"], j[0]["cur_slot"] = 0, 0, 1
        j[0]["n_ref"], j[0]["n_kf"] = pair["n_ref"], 0
        j[0]["T_ref"], j[0]["T_kf"], j[0]["T_cur"] = pair["T_ref"], pair["T_kf"], pair["T_cur_init"]
        res, st = ctx.sparse_align(j, pair["feats"], mode=mode, max_iter=30)
        tot += ctx.debug_cycles().reshape(-1)[44:64]
        evals += int(res[0]["evaluations"])
print("mode", mode, "pairs 12, evaluations", evals, "selections", 2 * evals)
print("hot attempts by shift:", tot[:8].tolist(), "sum", int(tot[:8].sum()))
print("hot hits by shift:    ", tot[8:16].tolist(), "sum", int(tot[8:16].sum()))
print("first evaluation of a level, bracket carried from the level above: median %d attempts / %d hits, MAD %d attempts / %d hits"
      % tuple(tot[16:20].tolist()))
