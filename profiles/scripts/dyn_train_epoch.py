"""Wall time of EnsembleDynamics.train epochs at the reference's scale (1M transitions, E=7, hidden 200x4, batch 256),
with the host reshuffle overlapped with the device epoch and without.  Usage: dyn_train_epoch.py [out.json]"""
import json
import os
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch

from offlinerlkit_b200.dynamics import EnsembleDynamics
from offlinerlkit_b200.modules import EnsembleDynamicsModel
from offlinerlkit_b200.synthetic import make_dataset
from offlinerlkit_b200.utils.logger import Logger
from offlinerlkit_b200.utils.scaler import StandardScaler
from offlinerlkit_b200.utils.termination_fns import termination_fn_halfcheetah

O, A, N, EPOCHS = 17, 6, 1_000_000, 3
data = make_dataset(N, O, A, seed=0)
data["rewards"] = data["rewards"].reshape(-1, 1)
out = {}
# one discarded short run first: library load, graph capture and allocator warm-up are not part of the comparison
_m = EnsembleDynamicsModel(O, A, [200, 200, 200, 200], num_ensemble=7, num_elites=5,
                           weight_decays=[2.5e-5, 5e-5, 7.5e-5, 7.5e-5, 1e-4], device="cuda:0")
_d = EnsembleDynamics(_m, torch.optim.Adam(_m.parameters(), lr=1e-3), StandardScaler(), termination_fn_halfcheetah)
_lg = Logger(tempfile.mkdtemp())
_lg.quiet = True
_d.train({k: v[:50_000] for k, v in data.items()}, _lg, max_epochs=1, max_epochs_since_update=100)
for flag in ("0", "1"):
    os.environ["ORLK_DYN_SHUFFLE_OVERLAP"] = flag
    torch.manual_seed(0)
    np.random.seed(0)
    model = EnsembleDynamicsModel(O, A, [200, 200, 200, 200], num_ensemble=7, num_elites=5,
                                  weight_decays=[2.5e-5, 5e-5, 7.5e-5, 7.5e-5, 1e-4], device="cuda:0")
    dyn = EnsembleDynamics(model, torch.optim.Adam(model.parameters(), lr=1e-3), StandardScaler(), termination_fn_halfcheetah)
    logger = Logger(tempfile.mkdtemp())
    logger.quiet = True
    t0 = time.perf_counter()
    dyn.train(data, logger, max_epochs=EPOCHS, max_epochs_since_update=100)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    out["overlap" if flag == "1" else "sequential"] = {"train_s": round(dt, 2), "epochs": EPOCHS, "rows": N,
                                                        "batches_per_epoch": -(-(N - 1000) // 256)}
    print(flag, out, flush=True)
if len(sys.argv) > 1:
    json.dump(out, open(sys.argv[1], "w"), indent=1)
