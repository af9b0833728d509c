"""ReLU-kink-aware gradient comparison (test infrastructure).

A ReLU network's gradient is defined only up to the decisions ``z > 0`` at pre-activations that are zero within rounding.
At BASELINE.json's sizes a CQL step evaluates 12 M pre-activations; a handful sit within 2e-7 of zero, the engine's
fp32-grade arithmetic (error ~2e-6 abs) decides some of them the other way than torch does, and because the critic
gradient is a heavily cancelling sum over 7936 rows a SINGLE flipped bit moves it by ~5e-4 in relative L2 -- although,
under the engine's own mask bits, the engine's gradients agree with a float64 evaluation to <= 1e-5
(profiles/mask_flip_r02.txt).  So the parity check is: the engine's gradient must equal the oracle's within ``tol``
AFTER the best choice of the oracle's near-zero mask bits -- g_engine ~= g_oracle + sum_j c_j d_j with d_j the exact
change of the oracle's gradient when bit j is flipped and every c_j in {0, 1} -- and nothing else may differ."""
import copy
from typing import Callable, Dict, List, Tuple

import numpy as np
import torch

from oracle import nets


class _KinkReLU(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z, rec, call_id):
        mask = z > 0
        if rec.tau > 0:
            zd = z.detach()
            near = (zd.abs() <= rec.tau).reshape(-1).nonzero().reshape(-1)
            width = zd.shape[-1]
            rows = zd.reshape(-1, width)
            for i in near.tolist():
                rec.found.append((call_id, i, float(zd.reshape(-1)[i])))
                rec.rows.append(torch.relu(rows[i // width]).clone())        # the post-activation row it belongs to
        forced = [(i, v) for (c, i, v) in rec.force if c == call_id]
        if forced:
            mask = mask.contiguous().clone()
            for i, v in forced:
                mask.view(-1)[i] = v
        ctx.save_for_backward(mask)
        return z * mask

    @staticmethod
    def backward(ctx, grad):
        (mask,) = ctx.saved_tensors
        return grad * mask, None, None


class KinkRecorder:
    """Installed as the oracle's ReLU for one ``step``: numbers the ReLU calls in execution order, records elements with
    ``|z| <= tau`` and overrides the mask bits listed in ``force`` [(call, flat index, bool)]."""

    def __init__(self, tau: float = 0.0, force=()):
        self.tau, self.force, self.found, self.rows, self.calls = tau, list(force), [], [], 0

    def __call__(self, z):
        cid = self.calls
        self.calls += 1
        return _KinkReLU.apply(z, self, cid)

    def __enter__(self):
        nets.set_relu_impl(self)
        return self

    def __exit__(self, *exc):
        nets.set_relu_impl(None)


def _flat(grads: Dict[str, torch.Tensor], names) -> np.ndarray:
    return np.concatenate([grads[k].detach().double().cpu().reshape(-1).numpy() for k in names])


def _per_tensor_ok(a: Dict[str, np.ndarray], b: Dict[str, np.ndarray], tol: float) -> List[Tuple[str, float, float]]:
    """Per tensor: relative L2 <= tol and max |diff| <= 2 tol max|g|.  One-element tensors (the scalar head's bias
    gradient, sum_m dq[m] over up to 7936 rows that cancels to ~0.4 % of sum |dq|) get 4 tol: fp32 summation ORDER alone
    moves such a sum by ~1e-4 of itself, in torch as much as in the engine."""
    bad = []
    for k in b:
        d = a[k] - b[k]
        l2 = np.sqrt((d * d).sum()) / max(np.sqrt((b[k] * b[k]).sum()), 1e-30)
        mx = np.abs(d).max() / max(np.abs(b[k]).max(), 1e-30)
        t = 4 * tol if b[k].size == 1 else tol
        if l2 > t or mx > 2 * t:
            bad.append((k, l2, mx))
    return bad


def engine_activation_rows(eng) -> Dict[int, torch.Tensor]:
    """width -> [rows, width] stack of every post-ReLU activation row the engine holds after a step: the ``H`` buffers of
    all its passes (MlpRun objects: attributes of the engine or kept alive by its launch plans), all members."""
    out: Dict[int, List[torch.Tensor]] = {}
    seen_t, seen_o = set(), set()

    def visit(v, depth):
        if depth > 4 or id(v) in seen_o or torch.is_tensor(v) or isinstance(v, (str, bytes, int, float, type(None))):
            return
        seen_o.add(id(v))
        if isinstance(v, dict):
            for x in v.values():
                visit(x, depth + 1)
            return
        if isinstance(v, (list, tuple)):
            for x in v:
                visit(x, depth + 1)
            return
        d = getattr(v, "__dict__", None)
        if not isinstance(d, dict):
            return
        H = d.get("H")
        if isinstance(H, (list, tuple)) and H and all(torch.is_tensor(h) for h in H):
            for h in H:
                if h.is_floating_point() and h.data_ptr() not in seen_t:
                    seen_t.add(h.data_ptr())
                    out.setdefault(h.shape[-1], []).append(h.reshape(-1, h.shape[-1]))
            return
        if type(v).__name__ == "Plan":
            visit(d.get("keep"), depth + 1)

    for v in list(eng.__dict__.values()):
        visit(v, 0)
    return {w: torch.cat(v, 0) for w, v in out.items()}


def engine_bits(eng, found, rows, match_tol: float = 3e-5):
    """For every near-zero pre-activation of the oracle, the ENGINE's decision: the oracle's post-activation row is looked
    up among the engine's activation rows (nearest in the max norm, which must be within ``match_tol``: the engine's
    forward error is ~2e-6) and the engine's bit is ``H_engine[row, col] > 0``.  Returns [(call, index, bit)] for the
    bits on which the two disagree; elements whose row the engine does not hold are skipped."""
    acts = engine_activation_rows(eng)
    flips = []
    for (c, i, z), row in zip(found, rows):
        w = row.numel()
        if w not in acts:
            continue
        E = acts[w]
        r = row.to(E.device, E.dtype)
        d = (E - r).abs().amax(dim=1)
        j = int(d.argmin())
        if float(d[j]) > match_tol:
            continue
        bit = bool(E[j, i % w] > 0)
        if bit != (z > 0):
            flips.append((c, i, bit))
    return flips


def assert_grads_close_up_to_kinks(got: Dict[str, torch.Tensor], ora_before, run_step: Callable, tol: float, what: str = "",
                                   taus=(5e-7, 4e-6), max_kinks: int = 48, eng=None) -> int:
    """``ora_before``: the oracle in its pre-step state (it is deep-copied, never advanced here); ``run_step(ora)`` runs the
    step on a copy and returns nothing (the gradients are read from ``ora.grads``).  Returns the number of mask bits that
    had to be flipped (0 = strict agreement).  Per tensor: rel-L2 <= tol and max |diff| <= 2 tol max|g|."""
    names = list(got.keys())
    gotn = {k: got[k].detach().double().cpu().reshape(-1).numpy() for k in names}

    def oracle_grads(tau, force):
        ora = copy.deepcopy(ora_before)
        with KinkRecorder(tau, force) as rec:
            run_step(ora)
        return {k: ora.grads[k].detach().double().cpu().reshape(-1).numpy() for k in names}, rec.found

    base, _ = oracle_grads(0.0, ())
    bad = _per_tensor_ok(gotn, base, tol)
    if not bad:
        return 0
    if eng is not None:
        # the engine's own decisions at the oracle's near-zero pre-activations, read from its activation buffers
        ora = copy.deepcopy(ora_before)
        with KinkRecorder(2e-5, ()) as rec:
            run_step(ora)
        flips = engine_bits(eng, rec.found, rec.rows)
        if flips:
            adj, _ = oracle_grads(0.0, flips)
            bad = _per_tensor_ok(gotn, adj, tol)
            if not bad:
                zs = {(c, i): z for c, i, z in rec.found}
                print(f"   [{what}] gradients agree under the engine's decisions at {len(flips)} of {len(rec.found)} ReLU "
                      f"pre-activations within 2e-5 of zero: " + ", ".join(f"z={zs[(c, i)]:.1e}" for c, i, _ in flips[:8]),
                      flush=True)
                return len(flips)
        raise AssertionError(f"{what}: engine gradients differ from the oracle beyond the {len(flips)} differing ReLU "
                             "decisions: " + "; ".join(f"{k} rel-L2 {l2:.2e} max {mx:.2e}" for k, l2, mx in bad[:6]))
    for tau in taus:
        _, found = oracle_grads(tau, ())
        assert len(found) <= max_kinks, f"{what}: {len(found)} pre-activations within {tau} of zero"
        if not found:
            continue
        flat = lambda d: np.concatenate([d[k] for k in names])
        resid = flat(gotn) - flat(base)
        D = []
        for (c, i, z) in found:
            gj, _ = oracle_grads(0.0, [(c, i, not (z > 0))])
            D.append(flat(gj) - flat(base))
        D = np.stack(D, 1)
        coef, *_ = np.linalg.lstsq(D, resid, rcond=None)
        pick = [j for j in range(len(found)) if coef[j] > 0.5]
        if any(abs(coef[j] - round(coef[j])) > 0.1 or round(coef[j]) not in (0, 1) for j in range(len(found))):
            continue
        force = [(found[j][0], found[j][1], not (found[j][2] > 0)) for j in pick]
        adj, _ = oracle_grads(0.0, force)          # exact: all chosen bits flipped together
        bad2 = _per_tensor_ok(gotn, adj, tol)
        if not bad2:
            print(f"   [{what}] gradients agree after flipping {len(pick)} of {len(found)} ReLU bits with |z| <= {tau:g}: "
                  + ", ".join(f"call {found[j][0]} z={found[j][2]:.1e}" for j in pick), flush=True)
            return len(pick)
        bad = bad2
    raise AssertionError(f"{what}: engine gradients differ from the oracle beyond ReLU-kink decisions: "
                         + "; ".join(f"{k} rel-L2 {l2:.2e} max {mx:.2e}" for k, l2, mx in bad[:6]))
