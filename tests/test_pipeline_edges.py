"""Host logic of the pipelined numpy boundary (GPUVecEnv -> np_env_step_host): the chunk edges."""
import pytest

from neuralplane_b200.envs.env_wrappers import DEFAULT_PIPELINE, MAX_PIPELINE_CHUNKS, pipeline_edges


@pytest.mark.parametrize("n", [1, 2, 255, 256, 257, 20_001, 1_000_000, 999_999])
@pytest.mark.parametrize("pattern", [1, 2, 4, 16, DEFAULT_PIPELINE, (1, 2, 3, 4, 5, 5, 5, 5, 5), (1, 1000)])
def test_edges_cover_the_population_in_order(n, pattern):
    chunks = pipeline_edges(n, pattern)
    if chunks is None:      # single launch
        return
    assert 2 <= len(chunks) <= MAX_PIPELINE_CHUNKS
    assert chunks[0][0] == 0 and chunks[-1][1] == n
    for (a0, a1), (b0, b1) in zip(chunks, chunks[1:]):
        assert a1 == b0
    for i0, i1 in chunks:
        assert i1 > i0 and i0 % 256 == 0       # whole pairs, 16-byte aligned rows (np_env_step_host / the TMA slab kernel)


def test_small_populations_are_one_launch():
    assert pipeline_edges(100, 4) is None and pipeline_edges(10**6, 1) is None


def test_bad_patterns_are_rejected():
    with pytest.raises(ValueError):
        pipeline_edges(10**6, 17)
    with pytest.raises(ValueError):
        pipeline_edges(10**6, (1, 0, 2))
