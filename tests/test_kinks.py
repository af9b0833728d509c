"""CPU: the ReLU-kink-aware gradient comparison used by the GPU parity tests (tests/kinks.py) recovers deliberately
flipped mask bits and reports strict agreement when nothing was flipped."""
import copy

import pytest

from tests.helpers import Golden
from tests.gpu_common import make_oracle
from tests.kinks import KinkRecorder, assert_grads_close_up_to_kinks


@pytest.mark.parametrize("name", ["cql_small", "edac_small", "iql_small"])
def test_flipped_relu_bits_are_recovered(name):
    g = Golden(name)
    ora, data = make_oracle(g.meta), g.dataset()
    has_noise = any(k.startswith("noise0|") for k in g.z.files)
    run = (lambda o: o.step(g.batch(0, data), g.noise(0))) if has_noise else (lambda o: o.step(g.batch(0, data)))
    tau = 2e-3
    while True:                                     # the smallest window holding a few pre-activations
        o = copy.deepcopy(ora)
        with KinkRecorder(tau) as rec:
            run(o)
        if len(rec.found) <= 20:
            break
        tau /= 2
    assert len(rec.found) >= 2
    picks = rec.found[:2]
    flipped = copy.deepcopy(ora)
    with KinkRecorder(0.0, [(c, i, not (z > 0)) for c, i, z in picks]):
        run(flipped)
    # (a bit whose upstream gradient is zero changes nothing and need not be "recovered")
    assert assert_grads_close_up_to_kinks(flipped.grads, ora, run, 1e-5, name, taus=(tau,)) in (1, 2)
    assert assert_grads_close_up_to_kinks(o.grads, ora, run, 1e-5, name) == 0
    with pytest.raises(AssertionError):             # a difference that is NOT a kink decision is reported
        wrong = {k: v * 1.001 for k, v in o.grads.items()}
        assert_grads_close_up_to_kinks(wrong, ora, run, 1e-5, name, taus=(tau,))
