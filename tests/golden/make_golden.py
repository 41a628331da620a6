"""Generate tests/golden/*.json with the CPU oracle (the reference itself cannot run here: no
MATLAB / Octave in the image -- parity unpinned; these vectors pin the oracle/GPU pair against
regressions).  Run from the repo root:  python tests/golden/make_golden.py"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from oracle.ds import DSConfig, ds_setup, ds_realization
from oracle import rng
from tests.helpers import err_from_oracle

S = ds_setup(DSConfig())
seed, first, n_iter, reps = 20181018, 100, 4, [0, 1, 2]
err = [err_from_oracle(ds_realization(S, rng.draws_for(S, seed, first + r)), n_iter).tolist() for r in reps]
out = dict(seed=seed, first_rep=first, n_iter=n_iter, reps=reps, err=err,
           layout="err[rep][snr][it][scheme aux,cod,ofdm][csi est,perfect][edge all,noedge]",
           config="DoublySelectiveChannelEstimation.m default parameters, oracle setup")
json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "ds_default_seeded_errors.json"), "w"))
d0 = rng.draws_for(S, seed, first)
S["chan"].NewRealization(d0["doppler_u"], d0["phase_u"])
h = S["chan"].ImpulseResponse
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "ds_default_h_rep100.npz"),
                    h=h, doppler_u=d0["doppler_u"], phase_u=d0["phase_u"])
print("written", np.array(err).shape, h.shape)
