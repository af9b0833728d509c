"""GPU: SAC (MOPO's learner), IQL and TD3+BC ``learn`` through the CUDA engine vs golden vectors from the real reference."""
import pytest

from tests.helpers import Golden

pytestmark = pytest.mark.gpu
TOL = 1e-4


@pytest.mark.parametrize("name", ["sac_small", "sac_hc"])
def test_sac_matches_reference(name):
    from tests.gpu_common import run_golden_steps
    run_golden_steps(Golden(name), tol=TOL, verbose=True)


@pytest.mark.parametrize("name", ["iql_small", "iql_walker", "iql_walker_b1024"])
def test_iql_matches_reference(name):
    from tests.gpu_common import run_golden_steps
    run_golden_steps(Golden(name), tol=TOL, verbose=True)


@pytest.mark.parametrize("name", ["td3bc_small", "td3bc_walker", "td3bc_walker_b1024"])
def test_td3bc_matches_reference(name):
    """td3bc_small runs 4 steps: the actor / polyak phase only runs on steps 0 and 2 (td3bc.py:107-116)."""
    from tests.gpu_common import run_golden_steps
    run_golden_steps(Golden(name), tol=TOL, verbose=True)


@pytest.mark.parametrize("name", ["edac_small", "edac_hc", "edac_small_maxq", "edac_hopper_e50"])
def test_edac_matches_reference(name):
    """Ensemble critics + the input-gradient diversity loss (hand-derived double backward) vs the reference's autograd."""
    from tests.gpu_common import run_golden_steps
    run_golden_steps(Golden(name), tol=TOL, verbose=True)
