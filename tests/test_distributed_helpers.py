"""CPU: host-side helpers of stochqn_b200.distributed (no device work)."""
import numpy as np
import pytest

from stochqn_b200 import _lib
from stochqn_b200.distributed import RowShardedCombiner, shard_bounds


def test_shard_bounds_cover_the_vector_once():
    for n in (1, 7, 100003, 1 << 20):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == n
            for (o0, c0), (o1, _) in zip(spans[:-1], spans[1:]):
                assert o0 + c0 == o1
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1


def test_row_sharded_combiner_argument_checks():
    abi = _lib.load(np.float64)
    with pytest.raises(ValueError):
        RowShardedCombiner(abi, None, 10, 3)
    one = RowShardedCombiner(abi, None, 12, 1)
    assert not one.p2p and one.blk == 12
    assert one.gather(0x1000) == 0x1000                 # one rank: the block is the vector
    with pytest.raises(RuntimeError):
        one.send_buffer()
    with pytest.raises(RuntimeError):
        one.reduce_scatter(0x1000)
