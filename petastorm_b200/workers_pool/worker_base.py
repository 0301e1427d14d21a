"""Worker protocol (petastorm/workers_pool/worker_base.py:18-35)."""


class WorkerBase(object):
    def __init__(self, worker_id, publish_func, args):
        self.worker_id = worker_id
        self.publish_func = publish_func
        self.args = args

    def process(self, *args, **kargs):
        raise NotImplementedError

    def shutdown(self):
        pass
