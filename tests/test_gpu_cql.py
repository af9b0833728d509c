"""GPU: CQLPolicy.learn through the CUDA engine vs golden vectors from the real reference (fp32, rtol 1e-4)."""
import pytest

from tests.helpers import Golden

pytestmark = pytest.mark.gpu
TOL = 1e-4      # north_star: losses / parameters within 1e-4 relative in fp32


@pytest.mark.parametrize("name", ["cql_small", "cql_small_lagrange", "cql_hopper", "cql_hc", "cql_hc_lagrange"])
def test_cql_matches_reference(name):
    from tests.gpu_common import run_golden_steps
    run_golden_steps(Golden(name), tol=TOL, verbose=True)


def test_cql_eager_equals_graph():
    """The captured CUDA graph and the eager launch sequence give identical results."""
    import torch
    from tests.gpu_common import run_golden_steps
    g = Golden("cql_small_lagrange")
    p1 = run_golden_steps(g, tol=TOL, use_graph=True)
    p2 = run_golden_steps(g, tol=TOL, use_graph=False)
    for (k, a), (_, b) in zip(p1.state_dict().items(), p2.state_dict().items()):
        assert torch.equal(a, b), k
