"""GPU replay of the call traces that were recorded AFTER the round's last GPU run (tests/trace_util.ORACLE_ONLY_TAGS:
merge_new=False, the portrait 20 x 15 grid, the 1024 x 1024 stress pair).

The CPU oracle is pinned on these records (tests/test_oracle_trace.py) and the CUDA path is bit-exact against the oracle on
the same shapes (tests/test_gpu_regroup.py, tests/test_gpu_subdivide.py), but these particular replays have not yet run on a
B200.  They are therefore non-strict xfail: a pass shows up as XPASS, a failure as XFAIL, neither breaks the suite -- and
the file sorts last, so nothing runs behind it.  Once seen green the tags move to trace_util.TAGS and this file goes away.
"""
from __future__ import annotations

import pytest

import trace_util as T
from test_gpu_trace import test_replay_reference_call as _replay

pytestmark = pytest.mark.gpu
EXTRA = T.records(T.ORACLE_ONLY_TAGS)


@pytest.mark.xfail(strict=False, reason="recorded after the round's last GPU run: oracle-verified, GPU replay not yet observed")
@pytest.mark.parametrize("tag,seq,name", EXTRA, ids=T.ids(EXTRA))
def test_replay_unverified_trace(tag, seq, name):
    import torch

    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    _replay(tag, seq, name)
