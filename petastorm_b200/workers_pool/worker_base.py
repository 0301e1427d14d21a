"""Import location kept for code written against ``petastorm.workers_pool.worker_base``."""
from petastorm_b200.workers_pool.gpu_pool import WorkerBase  # noqa: F401
