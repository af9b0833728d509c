import os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from offlinerlkit_b200.modules import EnsembleDynamicsModel
from offlinerlkit_b200.dynamics import EnsembleDynamics
from offlinerlkit_b200.buffer import ReplayBuffer
from offlinerlkit_b200.utils.logger import Logger
from offlinerlkit_b200.utils.scaler import StandardScaler
from offlinerlkit_b200.utils.termination_fns import get_termination_fn
from offlinerlkit_b200.synthetic import make_dataset
DEV = "cuda:0"
O, A = 5, 3
torch.manual_seed(0); np.random.seed(0)
data = make_dataset(2000, O, A, seed=3)
model = EnsembleDynamicsModel(O, A, [24, 24], num_ensemble=3, num_elites=2, weight_decays=[2.5e-5, 5e-5, 7.5e-5], device=DEV)
dyn = EnsembleDynamics(model, torch.optim.Adam(model.parameters(), lr=1e-3), StandardScaler(), get_termination_fn("halfcheetah-medium-v2"), penalty_coef=0.5)
real = ReplayBuffer(2000, (O,), np.float32, A, np.float32, device=DEV)
real.load_dataset(data)
logger = Logger(tempfile.mkdtemp()); logger.quiet = True
dyn.engine.use_graph = len(sys.argv) < 2
dyn.train(real.sample_all(), logger, max_epochs=2, max_epochs_since_update=5)
torch.cuda.synchronize()
print("ok", model.elites.tolist())
