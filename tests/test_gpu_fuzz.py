"""-m gpu: randomised parity sweep (profiles/fuzz_parity.py): random (D, B, T, K) over every kernel family --
2-CTA cluster chain, single-CTA chain/filler, unified D = 64 + tensor-core passes, row-split 4-CTA clusters,
checkpointed backward, the wave pipelines -- against the float64 oracle.  Per-clip loss 1e-4, gradients 1e-3
(worst row); the sweep's own margins are ~50x (loss <= 6e-6, gradients <= 4e-5 over 180 cases at build time)."""
import pytest

from profiles.fuzz_parity import run, run_samplers

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed", [11, 12])
def test_random_shapes_all_families(cuda, lib, seed):
    failures = run(n_cases=14, seed=seed, tmax_big=40, verbose=False)
    assert not failures, failures


def test_random_samplers_and_rho(cuda, lib):
    """Psi sampler (all three kernel families, D 1..128), Rho sampler and Rho loss + gradient on random shapes."""
    failures = run_samplers(n_cases=18, seed=21, verbose=False)
    assert not failures, failures
