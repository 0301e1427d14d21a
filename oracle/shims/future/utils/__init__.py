"""Harness-side stand-in for the `future` package (absent from this image).

TEST INFRASTRUCTURE ONLY -- lets the *unmodified* reference under /root/reference import
(`petastorm/utils.py:22` does `from future.utils import raise_with_traceback`)."""
import sys


def raise_with_traceback(exc, traceback=Ellipsis):
    if traceback is Ellipsis:
        _, _, traceback = sys.exc_info()
    raise exc.with_traceback(traceback)
