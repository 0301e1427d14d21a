"""petastorm_b200: a B200-native (sm_100a CUDA) Parquet -> GPU tensor input pipeline that keeps the read-side API of
uber/petastorm (``make_reader`` / ``make_batch_reader`` / ``petastorm.pytorch.DataLoader`` / ``Unischema`` / codecs /
``TransformSpec`` / predicates / ``NGram``).  The decode work runs in hand-written CUDA kernels behind the C-ABI of
``libpst_b200.so`` (``include/pst_b200.h``); there is no CPU fallback."""

__version__ = '0.1.0'

import os as _os

# CUDA multiplexes streams onto CUDA_DEVICE_MAX_CONNECTIONS hardware queues (default 8).  The readers keep 5 decode
# streams busy with ~300 MB copies and ~15 ms kernels; when the consumer's stream lands on the same hardware queue as
# one of them its work is falsely serialised behind a whole row-group (measured: +4 ms per step).  Must be set before
# the CUDA context exists, hence at import; an explicit user setting wins.
_os.environ.setdefault('CUDA_DEVICE_MAX_CONNECTIONS', '32')

from petastorm_b200.reader import make_reader, make_batch_reader  # noqa: F401,E402
from petastorm_b200.transform import TransformSpec  # noqa: F401,E402
