"""Inputs of the Gibbs sampler derived from a golden case exactly as rsem-run-gibbs derives them
(Gibbs.cpp:101-204): .ofg matrix, omit list, totc, eel from the .model's gld, mw."""
import numpy as np

import rsem_files as rf


def gibbs_inputs(case, pseudo_c=1.0):
    M, N0, rp, sid, val = rf.read_ofg(case.out("ofg20"))
    m = rf.read_model(case.out("model20"), case.M)
    init = np.zeros(M + 1, np.int32)
    init[case.omit] = -1
    n1 = len(rp) - 1
    totc = (M + 1 - len(case.omit)) * pseudo_c + N0 + n1
    return dict(row_ptr=rp, sid=sid, conprb=val, N0=N0, init_counts=init, alpha=np.full(M + 1, pseudo_c), totc=totc,
                eel=case.eel(m["gld"]), mw=m["mw"], gld=m["gld"])
