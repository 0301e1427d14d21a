"""petastorm_b200: a B200-native (sm_100a CUDA) Parquet -> GPU tensor input pipeline that keeps the read-side API of
uber/petastorm (``make_reader`` / ``make_batch_reader`` / ``petastorm.pytorch.DataLoader`` / ``Unischema`` / codecs /
``TransformSpec`` / predicates / ``NGram``).  The decode work runs in hand-written CUDA kernels behind the C-ABI of
``libpst_b200.so`` (``include/pst_b200.h``); there is no CPU fallback."""

__version__ = '0.1.0'

from petastorm_b200.reader import make_reader, make_batch_reader  # noqa: F401,E402
from petastorm_b200.transform import TransformSpec  # noqa: F401,E402
